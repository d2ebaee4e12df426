// ball.cu -- K5: fixed-radius query around an arbitrary centre.
//
// Replaces typed_particles_in_sphere (pynbody/kdtree/kdmain.cpp:655-679), i.e. smBallGather with
// the StoreResultInList sink (smooth.h:245-251,270-324): ids of all particles with d2 <= r*r, with
// the periodic image chosen per cell (kd.h:68-154).  A bucket passes the cell test only if all its
// ancestors do, so the hierarchy is not needed for a single query: one warp tests one bucket's bounds
// and, on a hit, its 32 particles (one per lane); hits are compacted with a warp-aggregated atomic.
//
// The radius is the caller's and may approach half the box, where a per-cell image choice depends
// on how large the cells happen to be.  Each particle is therefore measured against its own nearest
// image of the centre (per axis the nearest of c, fl(c+L), fl(c-L)): identical arithmetic to the
// reference wherever its per-cell choice is the nearest image, and equal to the brute-force
// "wrap then measure" answer its own test demands (tests/kdtree_test.py:258-282) for any radius.
#include "traverse.cuh"
#include <algorithm>

namespace pnb {

template <typename T, bool PERIODIC>
__global__ void k_ball(TreeView<T> tv, T cx, T cy, T cz, T r2, uint32_t *__restrict__ out,
                       unsigned long long *__restrict__ counter, unsigned long long cap)
{
    const int lane = threadIdx.x & 31;
    const int64_t bucket = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (bucket >= tv.nsplit) return;
    Box6<T> b = load_box(tv.box, tv.nsplit + bucket);
    T sx, sy, sz;
    T gap2 = box_gap2<T, PERIODIC>(b, cx, cy, cz, tv.period, sx, sy, sz);
    if (!(gap2 <= r2 * Ex<T>::slack())) return;            // warp-uniform
    const int start = tv.bucket_start[bucket];
    const int cnt = tv.bucket_start[bucket + 1] - start;
    bool in = false;
    if (lane < cnt) {
        T d2 = PERIODIC ? dist2_nearest<T>(cx, cy, cz, tv.x[start + lane], tv.y[start + lane], tv.z[start + lane], tv.period)
                        : dist2<T>(cx, cy, cz, tv.x[start + lane], tv.y[start + lane], tv.z[start + lane]);
        in = d2 <= r2;                                      // smooth.h:310
    }
    unsigned m = __ballot_sync(FULL, in);
    if (!m) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(FULL, base, 0);
    if (in) {
        unsigned long long slot = base + __popc(m & ((1u << lane) - 1));
        if (slot < cap) out[slot] = tv.order[start + lane];
    }
}

template <typename T>
static int64_t ball_typed(Tree &t, double cx, double cy, double cz, double r, double period, int64_t *out_host,
                          int64_t cap)
{
    TreeView<T> tv = make_view<T>(t, period);
    // centre and radius are parsed in the tree dtype (kdmain.cpp:662-667), r*r evaluated in it
    const T tcx = (T)cx, tcy = (T)cy, tcz = (T)cz, tr = (T)r;
    const T r2 = tr * tr;
    DevBuf counter(sizeof(unsigned long long));
    DevBuf out(sizeof(uint32_t) * (size_t)std::max<int64_t>(cap, 1));
    PNB_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(unsigned long long), t.stream));
    const int BS = 128;
    const unsigned grid = (unsigned)((t.nsplit * 32 + BS - 1) / BS);
    if (tv.periodic)
        PNB_LAUNCH((k_ball<T, true>), grid, BS, 0, t.stream, tv, tcx, tcy, tcz, r2, out.as<uint32_t>(),
                   counter.as<unsigned long long>(), (unsigned long long)cap);
    else
        PNB_LAUNCH((k_ball<T, false>), grid, BS, 0, t.stream, tv, tcx, tcy, tcz, r2, out.as<uint32_t>(),
                   counter.as<unsigned long long>(), (unsigned long long)cap);
    unsigned long long found = 0;
    PNB_CUDA(cudaMemcpyAsync(&found, counter.p, sizeof(found), cudaMemcpyDeviceToHost, t.stream));
    PNB_CUDA(cudaStreamSynchronize(t.stream));
    const int64_t take = std::min<int64_t>((int64_t)found, cap);
    if (take > 0) {
        std::vector<uint32_t> ids((size_t)take);
        PNB_CUDA(cudaMemcpy(ids.data(), out.p, sizeof(uint32_t) * (size_t)take, cudaMemcpyDeviceToHost));
        std::sort(ids.begin(), ids.end());
        for (int64_t i = 0; i < take; ++i) out_host[i] = (int64_t)ids[(size_t)i];
    }
    return (int64_t)found;
}

// ---- many balls at once: which particles lie inside ANY of nq balls? (multi-GPU halo selection) -----
// One warp takes 32 balls (one per lane) and sweeps the tree once for all of them, like the gather
// kernel; accepted candidates of all lanes are OR-ed and the owning lanes set the particle flags.
template <typename T, bool PERIODIC>
__global__ void __launch_bounds__(128) k_mark_in_balls(TreeView<T> tv, const T *__restrict__ c, const T *__restrict__ rad,
                                                       int64_t nq, uint8_t *__restrict__ flags)
{
    __shared__ T tiles[4][96];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t q = ((int64_t)blockIdx.x * 4 + wib) * 32 + lane;
    if (((int64_t)blockIdx.x * 4 + wib) * 32 >= nq) return;
    T *tile = tiles[wib];
    const bool active = q < nq;
    T qx = 0, qy = 0, qz = 0, r2 = -1;
    if (active) {
        qx = c[q]; qy = c[nq + q]; qz = c[2 * nq + q];
        const T r = rad[q];
        r2 = Ex<T>::mul(r, r);
    }
    int64_t cp = 1;
    for (;;) {
        Box6<T> b = load_box(tv.box, cp);
        T sx, sy, sz;
        T gap2 = box_gap2<T, PERIODIC>(b, qx, qy, qz, tv.period, sx, sy, sz);
        bool pass = active && gap2 <= r2 * Ex<T>::slack();
        if (__any_sync(FULL, pass)) {
            if (cp < tv.nsplit) { cp = 2 * cp; continue; }
            const int64_t bucket = cp - tv.nsplit;
            const int start = tv.bucket_start[bucket];
            const int cnt = tv.bucket_start[bucket + 1] - start;
            __syncwarp();
            if (lane < cnt) {
                tile[lane] = tv.x[start + lane]; tile[32 + lane] = tv.y[start + lane]; tile[64 + lane] = tv.z[start + lane];
            }
            __syncwarp();
            unsigned mask = 0;
            const bool wide = PERIODIC && __any_sync(FULL, pass && shift_is_ambiguous<T, PERIODIC>(b, sx, sy, sz, tv.period));
            for (int j = 0; j < cnt; ++j) {
                T d2 = wide ? dist2_nearest<T>(qx, qy, qz, tile[j], tile[32 + j], tile[64 + j], tv.period)
                            : dist2<T>(sx, sy, sz, tile[j], tile[32 + j], tile[64 + j]);
                mask |= (unsigned)(d2 <= r2) << j;
            }
            if (!pass) mask = 0;
            mask = __reduce_or_sync(FULL, mask);
            if (lane < cnt && ((mask >> lane) & 1u)) flags[start + lane] = 1;
        }
        cp = next_node(cp);
        if (cp == 1) break;
    }
}

template <typename T>
static void mark_typed(Tree &t, const void *centres, const void *radii, int64_t nq, double period, uint8_t *flags)
{
    TreeView<T> tv = make_view<T>(t, period);
    tv.query_mask = nullptr;
    const unsigned grid = (unsigned)((nq + 127) / 128);
    if (tv.periodic) PNB_LAUNCH((k_mark_in_balls<T, true>), grid, 128, 0, t.stream, tv, (const T *)centres, (const T *)radii, nq, flags);
    else PNB_LAUNCH((k_mark_in_balls<T, false>), grid, 128, 0, t.stream, tv, (const T *)centres, (const T *)radii, nq, flags);
}

void run_mark_in_balls(Tree &t, const void *centres_cols, const void *radii, int64_t nq, double period, uint8_t *flags_tree)
{
    if (nq <= 0) return;
    if (t.dtype == PNB_F64) mark_typed<double>(t, centres_cols, radii, nq, period, flags_tree);
    else mark_typed<float>(t, centres_cols, radii, nq, period, flags_tree);
}

int64_t run_ball_query(Tree &t, double cx, double cy, double cz, double r, double period, int64_t *out_host,
                       int64_t cap)
{
    if (t.dtype == PNB_F64) return ball_typed<double>(t, cx, cy, cz, r, period, out_host, cap);
    return ball_typed<float>(t, cx, cy, cz, r, period, out_host, cap);
}

}  // namespace pnb
