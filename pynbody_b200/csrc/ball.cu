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
__device__ __forceinline__ T nearest_image_delta(T c, T x, T L)
{
    T d = Ex<T>::sub(c, x);
    if (PERIODIC) {
        T dp = Ex<T>::sub(Ex<T>::add(c, L), x), dm = Ex<T>::sub(Ex<T>::sub(c, L), x);
        if (fabs(dp) < fabs(d)) d = dp;
        if (fabs(dm) < fabs(d)) d = dm;
    }
    return d;
}

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
        T dx = nearest_image_delta<T, PERIODIC>(cx, tv.x[start + lane], tv.period);
        T dy = nearest_image_delta<T, PERIODIC>(cy, tv.y[start + lane], tv.period);
        T dz = nearest_image_delta<T, PERIODIC>(cz, tv.z[start + lane], tv.period);
        T d2 = Ex<T>::add(Ex<T>::add(Ex<T>::mul(dx, dx), Ex<T>::mul(dy, dy)), Ex<T>::mul(dz, dz));
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

int64_t run_ball_query(Tree &t, double cx, double cy, double cz, double r, double period, int64_t *out_host,
                       int64_t cap)
{
    if (t.dtype == PNB_F64) return ball_typed<double>(t, cx, cy, cz, r, period, out_host, cap);
    return ball_typed<float>(t, cx, cy, cz, r, period, out_host, cap);
}

}  // namespace pnb
