// gather.cu -- K3/K4: fixed-radius (2h) neighbour gather fused with the SPH reductions.
//
// Replaces typed_populate's gather loop (pynbody/kdtree/kdmain.cpp:764-790), smBallGather
// (smooth.h:270-324) and the reductions smDensity / smMeanQty1D / smMeanQtyND / smDispQty1D /
// smDispQtyND / smDivQty / smCurlQty (smooth.h:626-918).
//
// Same warp-per-bucket mapping as the kNN kernel (knn.cu): the warp sweeps the tree once from the
// root for its 32 query particles, each lane testing every node against its own ball of radius
// 2 h_i (periodic cell test kd.h:68-154, acceptance d2 <= fl(4 h h), smooth.h:310 with
// kdmain.cpp:779).  Candidate buckets are staged in shared memory together with the per-particle
// fields the reduction needs (mass, rho, quantity); accepted candidates are collected in a register
// bit-mask and reduced with all lanes working together.  Nothing is written to global memory except
// the final per-particle result: no neighbour lists are materialised.
#include "traverse.cuh"

namespace pnb {

constexpr int GATHER_WARPS = 4;
constexpr int RESMOOTH_SAFE = 500;   // smooth.h:15

enum { OP_RHO = 2, OP_MEAN1 = 3, OP_MEAN3 = 4, OP_DISP1 = 5, OP_DISP3 = 6, OP_DIV = 7, OP_CURL = 8 };

template <typename T, typename Q>
struct GatherArgs {
    TreeView<T> tv;
    int k;
    int kernel_id;
    double wendland_self;
    const T *smooth_t;      // [n]
    const T *rho_t;         // [n] (ops >= 3)
    const Q *qty_t;         // [ncols][n] (ops >= 3)
    void *out_t;            // T[n] for OP_RHO, Q[ncols_out][n] otherwise
    int *overflow;
};

template <int OP> struct OpTraits {
    static constexpr int qcols = (OP == OP_MEAN1 || OP == OP_DISP1) ? 1 : (OP == OP_RHO ? 0 : 3);
    static constexpr int ocols = (OP == OP_MEAN3 || OP == OP_CURL) ? 3 : 1;
    static constexpr bool two_pass = (OP == OP_DISP1 || OP == OP_DISP3);
    static constexpr bool grad = (OP == OP_DIV || OP == OP_CURL);
};

// shared-memory tile of one candidate bucket: x,y,z,mass,rho (T) then up to 3 quantity columns (Q)
template <typename T, typename Q>
struct Tile {
    T x[32], y[32], z[32], m[32], rho[32];
    Q q[3][32];
};

template <typename T, typename Q, int OP, bool PERIODIC, int PASS>
__device__ __forceinline__ void gather_sweep(const GatherArgs<T, Q> &a, Tile<T, Q> &tile, int lane, bool active,
                                             T qx, T qy, T qz, T ball2, T ih2, double norm, const double *qi,
                                             const double *mean, double *acc, int &count)
{
    using OT = OpTraits<OP>;
    const TreeView<T> &tv = a.tv;
    const T L = tv.period;
    int64_t cp = 1;
    for (;;) {
        Box6<T> b = load_box(tv.box, cp);
        T sx, sy, sz;
        T gap2 = box_gap2<T, PERIODIC>(b, qx, qy, qz, L, sx, sy, sz);
        bool pass = active && gap2 <= ball2 * Ex<T>::slack();
        if (__any_sync(FULL, pass)) {
            if (cp < tv.nsplit) { cp = 2 * cp; continue; }
            const int64_t bucket = cp - tv.nsplit;
            const int start = tv.bucket_start[bucket];
            const int cnt = tv.bucket_start[bucket + 1] - start;
            __syncwarp();
            if (lane < cnt) {
                const int64_t j = start + lane;
                tile.x[lane] = tv.x[j]; tile.y[lane] = tv.y[j]; tile.z[lane] = tv.z[j];
                tile.m[lane] = tv.mass[j];
                if (OP != OP_RHO) {
                    tile.rho[lane] = a.rho_t[j];
#pragma unroll
                    for (int c = 0; c < OT::qcols; ++c) tile.q[c][lane] = a.qty_t[(int64_t)c * tv.n + j];
                }
            }
            __syncwarp();
            unsigned mask = 0;
            // some lane's shift is not the nearest image for every particle of this cell: image per particle
            const bool wide = PERIODIC && __any_sync(FULL, pass && shift_is_ambiguous<T, PERIODIC>(b, sx, sy, sz, L));
            if (!wide) {
#pragma unroll 4
                for (int j = 0; j < cnt; ++j) {
                    T d2 = dist2<T>(sx, sy, sz, tile.x[j], tile.y[j], tile.z[j]);
                    mask |= (unsigned)(d2 <= ball2) << j;      // smooth.h:310 (non-strict)
                }
            } else {
                for (int j = 0; j < cnt; ++j) {
                    T d2 = dist2_nearest<T>(qx, qy, qz, tile.x[j], tile.y[j], tile.z[j], L);
                    mask |= (unsigned)(d2 <= ball2) << j;
                }
            }
            if (!pass) mask = 0;
            if (PASS == 0) count += __popc(mask);
            while (__any_sync(FULL, mask != 0)) {
                if (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const T d2 = wide ? dist2_nearest<T>(qx, qy, qz, tile.x[j], tile.y[j], tile.z[j], L)
                                      : dist2<T>(sx, sy, sz, tile.x[j], tile.y[j], tile.z[j]);
                    const double mj = (double)tile.m[j];
                    if (OP == OP_RHO) {
                        // rho_i += (W * norm) * m_j, accumulated in T by the caller's cast (smooth.h:640-645)
                        T w = (T)kern_w(a.kernel_id, a.wendland_self, (double)(d2 * ih2));
                        w *= (T)norm;
                        acc[0] = (double)((T)acc[0] + w * tile.m[j]);
                    } else if (!OT::grad) {
                        T w = (T)kern_w(a.kernel_id, a.wendland_self, (double)(d2 * ih2));
                        w *= (T)norm;
                        const double wm = (double)w * mj / (double)tile.rho[j];
                        if (!OT::two_pass || PASS == 0) {
#pragma unroll
                            for (int c = 0; c < OT::qcols; ++c) acc[c] = (double)((Q)acc[c] + (Q)(wm * (double)tile.q[c][j]));
                        } else {
#pragma unroll
                            for (int c = 0; c < OT::qcols; ++c) {
                                Q tq = (Q)mean[c] - tile.q[c][j];
                                acc[0] = (double)((Q)acc[0] + (Q)(wm * (double)tq * (double)tq));
                            }
                        }
                    } else {
                        // raw UNWRAPPED separation (smooth.h:736-738, :791-793), wrapped r2 in the kernel
                        const T dx = qx - tile.x[j], dy = qy - tile.y[j], dz = qz - tile.z[j];
                        T g = (T)kern_g(a.kernel_id, (double)(d2 * ih2), (double)d2);
                        g *= (T)norm;
                        const T dq0 = (T)((double)tile.q[0][j] - qi[0]);
                        const T dq1 = (T)((double)tile.q[1][j] - qi[1]);
                        const T dq2 = (T)((double)tile.q[2][j] - qi[2]);
                        const double gm = (double)g * mj / (double)tile.rho[j];
                        if (OP == OP_DIV) {
                            const T dv = dx * dq0 + dy * dq1 + dz * dq2;
                            acc[0] = (double)((Q)acc[0] + (Q)((double)g * (double)dv * mj / (double)tile.rho[j]));
                        } else {
                            const T c0 = dy * dq2 - dz * dq1, c1 = dz * dq0 - dx * dq2, c2 = dx * dq1 - dy * dq0;
                            acc[0] = (double)((Q)acc[0] + (Q)(gm * (double)c0));
                            acc[1] = (double)((Q)acc[1] + (Q)(gm * (double)c1));
                            acc[2] = (double)((Q)acc[2] + (Q)(gm * (double)c2));
                        }
                    }
                }
            }
        }
        cp = next_node(cp);
        if (cp == 1) break;
    }
}

template <typename T, typename Q, int OP, bool PERIODIC>
__global__ void __launch_bounds__(GATHER_WARPS * 32) k_gather(GatherArgs<T, Q> a)
{
    using OT = OpTraits<OP>;
    __shared__ Tile<T, Q> tiles[GATHER_WARPS];
    const TreeView<T> &tv = a.tv;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t bucket = (int64_t)blockIdx.x * GATHER_WARPS + wib;
    if (bucket >= tv.nsplit) return;
    Tile<T, Q> &tile = tiles[wib];

    const int start = tv.bucket_start[bucket];
    const int cnt = tv.bucket_start[bucket + 1] - start;
    const bool active = lane < cnt && (!tv.query_mask || tv.query_mask[start + lane]);
    if (!__any_sync(FULL, active)) return;
    const int64_t p = start + lane;
    T qx = 0, qy = 0, qz = 0, h = 1;
    double qi[3] = {0, 0, 0};
    if (active) {
        qx = tv.x[p]; qy = tv.y[p]; qz = tv.z[p];
        h = a.smooth_t[p];
        if (OT::grad)
            for (int c = 0; c < 3; ++c) qi[c] = (double)(T)a.qty_t[(int64_t)c * tv.n + p];   // (T) cast: smooth.h:727-733
    }
    const T ball2 = Ex<T>::mul(Ex<T>::mul((T)4, h), h);              // kdmain.cpp:779
    const T ih = (T)(1.0 / (double)h);
    const T ih2 = ih * ih;
    // 1/(pi h^3) for values, 1/(pi h^4) for gradients (smooth.h:634-636, :724-726)
    const double norm = OT::grad ? (double)(T)(0.31830988618379067154 * (double)ih2 * (double)ih2)
                                 : (double)(T)(0.31830988618379067154 * (double)ih * (double)ih2);
    double acc[3] = {0, 0, 0}, mean[3] = {0, 0, 0};
    int count = 0;
    gather_sweep<T, Q, OP, PERIODIC, 0>(a, tile, lane, active, qx, qy, qz, ball2, ih2, norm, qi, mean, acc, count);
    if (OT::two_pass) {
        for (int c = 0; c < 3; ++c) { mean[c] = acc[c]; acc[c] = 0; }
        gather_sweep<T, Q, OP, PERIODIC, 1>(a, tile, lane, active, qx, qy, qz, ball2, ih2, norm, qi, mean, acc, count);
    }
    if (!active) return;
    if (count > a.k + RESMOOTH_SAFE) atomicOr(a.overflow, 1);      // smooth.h:253-267
    if (OP == OP_RHO) {
        reinterpret_cast<T *>(a.out_t)[p] = (T)acc[0];
    } else if (OT::two_pass) {
        reinterpret_cast<Q *>(a.out_t)[p] = (Q)sqrt(acc[0]);         // smooth.h:866, :917
    } else {
#pragma unroll
        for (int c = 0; c < OT::ocols; ++c) reinterpret_cast<Q *>(a.out_t)[(int64_t)c * tv.n + p] = (Q)acc[c];
    }
}

template <typename T, typename Q, int OP>
static void launch_gather(Tree &t, GatherArgs<T, Q> &a)
{
    const unsigned grid = (unsigned)((t.nsplit + GATHER_WARPS - 1) / GATHER_WARPS);
    KernelTimer timer(1, t.stream);
    if (a.tv.periodic) PNB_LAUNCH((k_gather<T, Q, OP, true>), grid, GATHER_WARPS * 32, 0, t.stream, a);
    else PNB_LAUNCH((k_gather<T, Q, OP, false>), grid, GATHER_WARPS * 32, 0, t.stream, a);
    timer.stop();
}

template <typename T, typename Q>
static bool run_gather_typed(Tree &t, int propid, int k, double period, const KernelParams &kp,
                             const void *smooth_tree, const void *rho_tree, const void *qty_tree, void *out_tree)
{
    DevBuf flag(sizeof(int));
    PNB_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), t.stream));
    GatherArgs<T, Q> a;
    a.tv = make_view<T>(t, period);
    a.k = k;
    a.kernel_id = kp.kernel_id;
    a.wendland_self = kp.wendland_self;
    a.smooth_t = (const T *)smooth_tree;
    a.rho_t = (const T *)rho_tree;
    a.qty_t = (const Q *)qty_tree;
    a.out_t = out_tree;
    a.overflow = flag.as<int>();
    switch (propid) {
    case OP_RHO: launch_gather<T, Q, OP_RHO>(t, a); break;
    case OP_MEAN1: launch_gather<T, Q, OP_MEAN1>(t, a); break;
    case OP_MEAN3: launch_gather<T, Q, OP_MEAN3>(t, a); break;
    case OP_DISP1: launch_gather<T, Q, OP_DISP1>(t, a); break;
    case OP_DISP3: launch_gather<T, Q, OP_DISP3>(t, a); break;
    case OP_DIV: launch_gather<T, Q, OP_DIV>(t, a); break;
    case OP_CURL: launch_gather<T, Q, OP_CURL>(t, a); break;
    default: throw Error(PNB_ERR_VALUE, "unknown smoothing property id");
    }
    int h = 0;
    PNB_CUDA(cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, t.stream));
    PNB_CUDA(cudaStreamSynchronize(t.stream));
    return h != 0;
}

bool run_gather(Tree &t, int propid, int k, double period, const KernelParams &kp, const void *smooth_tree,
                const void *rho_tree, const void *qty_tree, int qdtype, void *out_tree)
{
    if (!t.has_mass) throw Error(PNB_ERR_VALUE, "mass array has not been set");
    if (t.dtype == PNB_F64) {
        if (qdtype == PNB_F64 || propid == OP_RHO)
            return run_gather_typed<double, double>(t, propid, k, period, kp, smooth_tree, rho_tree, qty_tree, out_tree);
        return run_gather_typed<double, float>(t, propid, k, period, kp, smooth_tree, rho_tree, qty_tree, out_tree);
    }
    if (qdtype == PNB_F32 || propid == OP_RHO)
        return run_gather_typed<float, float>(t, propid, k, period, kp, smooth_tree, rho_tree, qty_tree, out_tree);
    return run_gather_typed<float, double>(t, propid, k, period, kp, smooth_tree, rho_tree, qty_tree, out_tree);
}

}  // namespace pnb
