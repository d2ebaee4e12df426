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
// fields the reduction needs (mass, mass/rho, quantity); the accepted (query, candidate) pairs of a
// bucket are listed in shared memory and evaluated one pair per lane (see gather_sweep).  Nothing is
// written to global memory except the final per-particle result: no neighbour lists are materialised.
#include "traverse.cuh"

namespace pnb {

constexpr int GATHER_WARPS = 4;
constexpr int RESMOOTH_SAFE = 500;   // smooth.h:15

enum { OP_RHO = 2, OP_MEAN1 = 3, OP_MEAN3 = 4, OP_DISP1 = 5, OP_DISP3 = 6, OP_DIV = 7, OP_CURL = 8,
       OP_COUNT = 9 /* only the neighbour count nCnt of the gather (kdmain.cpp:779) */ };

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
    static constexpr int qcols = (OP == OP_MEAN1 || OP == OP_DISP1) ? 1 : ((OP == OP_RHO || OP == OP_COUNT) ? 0 : 3);
    static constexpr int ocols = (OP == OP_MEAN3 || OP == OP_CURL) ? 3 : 1;
    static constexpr bool two_pass = (OP == OP_DISP1 || OP == OP_DISP3);
    static constexpr bool grad = (OP == OP_DIV || OP == OP_CURL);
};

// shared-memory working set of one warp
template <typename T, typename Q>
struct Tile {
    T x[32], y[32], z[32], m[32];               // candidate bucket: coordinates, mass
    double mr[32];                              // ... mass / density (one division per candidate, not per pair)
    Q q[3][32];                                 // ... and up to 3 quantity columns
    T sx[32], sy[32], sz[32];                   // per query lane: its position shifted for this bucket
    T ball2[32];                                // per query lane: (2h)^2
    double c[8][32];                            // per query lane: ih2, norm, qx, qy, qz, e0, e1, e2 (e = own quantity / mean)
    double acc[3][32];                          // per query lane: running sums
    uint16_t pairs[1024];                       // accepted (query lane << 5 | candidate) pairs of the current bucket
};

// One sweep of the tree for the warp's 32 queries.  Per candidate bucket:
//   1. lane j holds candidate j; the queries in reach are taken in turn (position and ball radius broadcast by
//      shuffle), the 32 verdicts of one query come back as a ballot and the accepted pairs are appended to a list
//      in shared memory;
//   2. the list is worked off 32 pairs at a time -- every lane evaluates one (query, candidate) interaction, so
//      all lanes are busy whatever the geometry (taking each lane's own candidates in turn kept ~8 of 32 busy);
//      the list is ordered by query, so the terms of one query sit in adjacent lanes and are summed by a
//      segmented shuffle reduction whose head lane adds the sum to the query's accumulator.
template <typename T, typename Q, int OP, bool PERIODIC, int PASS>
__device__ __forceinline__ void gather_sweep(const GatherArgs<T, Q> &a, Tile<T, Q> &tile, int lane, bool active,
                                             T qx, T qy, T qz, T ball2, int &count)
{
    using OT = OpTraits<OP>;
    constexpr int NC = (OT::two_pass && PASS == 1) ? 1 : (OP == OP_RHO || OP == OP_DIV || OP == OP_COUNT ? 1 : OT::qcols);
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
            T cx = 0, cy = 0, cz = 0;
            if (lane < cnt) {
                const int64_t j = start + lane;
                cx = tv.x[j]; cy = tv.y[j]; cz = tv.z[j];
                tile.x[lane] = cx; tile.y[lane] = cy; tile.z[lane] = cz;
                if (OP != OP_COUNT) tile.m[lane] = tv.mass[j];
                if (OP != OP_RHO && OP != OP_COUNT) {
                    tile.mr[lane] = (double)tile.m[lane] / (double)a.rho_t[j];
#pragma unroll
                    for (int c = 0; c < OT::qcols; ++c) tile.q[c][lane] = a.qty_t[(int64_t)c * tv.n + j];
                }
            }
            tile.sx[lane] = sx; tile.sy[lane] = sy; tile.sz[lane] = sz;
            // some lane's shift is not the nearest image for every particle of this cell: image per particle
            const bool wide = PERIODIC && __any_sync(FULL, pass && shift_is_ambiguous<T, PERIODIC>(b, sx, sy, sz, L));
            __syncwarp();
            // 1. accepted pairs, query by query
            int npairs = 0;
            unsigned todo = __ballot_sync(FULL, pass);
            while (todo) {
                const int i = __ffs(todo) - 1;
                todo &= todo - 1;
                // (the query's data comes from shared memory: one broadcast load per value instead of one or two
                // shuffles)
                const T ub = tile.ball2[i];
                T d2;
                if (!wide) d2 = dist2<T>(tile.sx[i], tile.sy[i], tile.sz[i], cx, cy, cz);
                else d2 = dist2_nearest<T>((T)tile.c[2][i], (T)tile.c[3][i], (T)tile.c[4][i], cx, cy, cz, L);
                const bool in = lane < cnt && d2 <= ub;                         // smooth.h:310 (non-strict)
                const unsigned row = __ballot_sync(FULL, in);
                if (OP != OP_COUNT && in) tile.pairs[npairs + __popc(row & ((1u << lane) - 1u))] = (uint16_t)(i << 5 | lane);
                if (PASS == 0 && lane == i) count += __popc(row);
                if (OP != OP_COUNT) npairs += __popc(row);
            }
            __syncwarp();
            // 2. one interaction per lane
            for (int base = 0; base < npairs; base += 32) {
                const bool valid = base + lane < npairs;
                const unsigned code = valid ? tile.pairs[base + lane] : 0u;
                const int i = valid ? (int)(code >> 5) : 32 + lane, j = (int)(code & 31u);
                const int ii = i & 31;
                const T px = tile.x[j], py = tile.y[j], pz = tile.z[j];
                const T d2 = wide ? dist2_nearest<T>((T)tile.c[2][ii], (T)tile.c[3][ii], (T)tile.c[4][ii], px, py, pz, L)
                                  : dist2<T>(tile.sx[ii], tile.sy[ii], tile.sz[ii], px, py, pz);
                const T ih2 = (T)tile.c[0][ii];
                const T norm = (T)tile.c[1][ii];
                double v[NC];
                if (OP == OP_RHO) {
                    // rho_i += (W * norm) * m_j, each term rounded as the reference rounds it (smooth.h:640-645)
                    T w = (T)kern_w(a.kernel_id, a.wendland_self, d2 * ih2);       // (float or double overload)
                    w *= norm;
                    v[0] = (double)(w * tile.m[j]);
                } else if (!OT::grad) {
                    T w = (T)kern_w(a.kernel_id, a.wendland_self, d2 * ih2);       // (float or double overload)
                    w *= norm;
                    const double wm = (double)w * tile.mr[j];
                    if (!OT::two_pass || PASS == 0) {
#pragma unroll
                        for (int c = 0; c < OT::qcols; ++c) v[c] = (double)(Q)(wm * (double)tile.q[c][j]);
                    } else {
                        v[0] = 0;
#pragma unroll
                        for (int c = 0; c < OT::qcols; ++c) {
                            Q tq = (Q)tile.c[5 + c][ii] - tile.q[c][j];
                            v[0] += (double)(Q)(wm * (double)tq * (double)tq);
                        }
                    }
                } else {
                    // raw UNWRAPPED separation (smooth.h:736-738, :791-793), wrapped r2 in the kernel
                    const T dx = (T)tile.c[2][ii] - px, dy = (T)tile.c[3][ii] - py, dz = (T)tile.c[4][ii] - pz;
                    T g = (T)kern_g(a.kernel_id, d2 * ih2, d2);
                    g *= norm;
                    const T dq0 = (T)((double)tile.q[0][j] - tile.c[5][ii]);
                    const T dq1 = (T)((double)tile.q[1][j] - tile.c[6][ii]);
                    const T dq2 = (T)((double)tile.q[2][j] - tile.c[7][ii]);
                    const double gm = (double)g * tile.mr[j];
                    if (OP == OP_DIV) {
                        const T dv = dx * dq0 + dy * dq1 + dz * dq2;
                        v[0] = (double)(Q)(gm * (double)dv);
                    } else {
                        const T c0 = dy * dq2 - dz * dq1, c1 = dz * dq0 - dx * dq2, c2 = dx * dq1 - dy * dq0;
                        v[0] = (double)(Q)(gm * (double)c0);
                        if (NC > 1) v[1 % NC] = (double)(Q)(gm * (double)c1);
                        if (NC > 2) v[2 % NC] = (double)(Q)(gm * (double)c2);
                    }
                }
                if (!valid) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) v[c] = 0;
                }
                // segmented sum over the lanes that hold the same query
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    const int ki = __shfl_down_sync(FULL, i, s);
                    const bool same = lane + s < 32 && ki == i;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const double o = __shfl_down_sync(FULL, v[c], s);
                        if (same) v[c] += o;
                    }
                }
                const int prev = __shfl_up_sync(FULL, i, 1);
                if (valid && (lane == 0 || prev != i)) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) tile.acc[c][i] += v[c];
                }
                __syncwarp();
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
    extern __shared__ __align__(16) unsigned char gather_smem[];
    const TreeView<T> &tv = a.tv;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t bucket = (int64_t)blockIdx.x * GATHER_WARPS + wib;
    if (bucket >= tv.nsplit) return;
    Tile<T, Q> &tile = reinterpret_cast<Tile<T, Q> *>(gather_smem)[wib];

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
    tile.ball2[lane] = ball2;
    tile.c[0][lane] = (double)ih2; tile.c[1][lane] = norm;
    tile.c[2][lane] = (double)qx; tile.c[3][lane] = (double)qy; tile.c[4][lane] = (double)qz;
    for (int c = 0; c < 3; ++c) { tile.c[5 + c][lane] = qi[c]; tile.acc[c][lane] = 0; }
    int count = 0;
    gather_sweep<T, Q, OP, PERIODIC, 0>(a, tile, lane, active, qx, qy, qz, ball2, count);
    if (OT::two_pass) {
        __syncwarp();
        for (int c = 0; c < 3; ++c) { tile.c[5 + c][lane] = tile.acc[c][lane]; tile.acc[c][lane] = 0; }   // the means
        gather_sweep<T, Q, OP, PERIODIC, 1>(a, tile, lane, active, qx, qy, qz, ball2, count);
    }
    __syncwarp();
    if (!active) return;
    if (OP == OP_COUNT) { reinterpret_cast<int32_t *>(a.out_t)[p] = count; return; }
    if (count > a.k + RESMOOTH_SAFE) atomicOr(a.overflow, 1);      // smooth.h:253-267
    if (OP == OP_RHO) {
        reinterpret_cast<T *>(a.out_t)[p] = (T)tile.acc[0][lane];
    } else if (OT::two_pass) {
        reinterpret_cast<Q *>(a.out_t)[p] = (Q)sqrt(tile.acc[0][lane]);         // smooth.h:866, :917
    } else {
#pragma unroll
        for (int c = 0; c < OT::ocols; ++c) reinterpret_cast<Q *>(a.out_t)[(int64_t)c * tv.n + p] = (Q)tile.acc[c][lane];
    }
}

template <typename T, typename Q, int OP>
static void launch_gather(Tree &t, GatherArgs<T, Q> &a)
{
    const unsigned grid = (unsigned)((t.nsplit + GATHER_WARPS - 1) / GATHER_WARPS);
    const size_t smem = sizeof(Tile<T, Q>) * GATHER_WARPS;
    auto kp = k_gather<T, Q, OP, true>;
    auto ko = k_gather<T, Q, OP, false>;
    PNB_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PNB_CUDA(cudaFuncSetAttribute(ko, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KernelTimer timer(1, t.stream);
    if (a.tv.periodic) PNB_LAUNCH(kp, grid, GATHER_WARPS * 32, smem, t.stream, a);
    else PNB_LAUNCH(ko, grid, GATHER_WARPS * 32, smem, t.stream, a);
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

void run_gather_counts(Tree &t, double period, const void *smooth_tree, int32_t *counts_tree)
{
    const KernelParams kp = make_kernel_params(0, 1);
    if (t.dtype == PNB_F64) {
        GatherArgs<double, double> a;
        a.tv = make_view<double>(t, period);
        a.k = 1; a.kernel_id = 0; a.wendland_self = kp.wendland_self;
        a.smooth_t = (const double *)smooth_tree; a.rho_t = nullptr; a.qty_t = nullptr;
        a.out_t = counts_tree; a.overflow = nullptr;
        launch_gather<double, double, OP_COUNT>(t, a);
    } else {
        GatherArgs<float, float> a;
        a.tv = make_view<float>(t, period);
        a.k = 1; a.kernel_id = 0; a.wendland_self = kp.wendland_self;
        a.smooth_t = (const float *)smooth_tree; a.rho_t = nullptr; a.qty_t = nullptr;
        a.out_t = counts_tree; a.overflow = nullptr;
        launch_gather<float, float, OP_COUNT>(t, a);
    }
    PNB_CUDA(cudaStreamSynchronize(t.stream));
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
