// knn.cu -- K2: exact k-nearest-neighbour search -> smoothing lengths (optionally fused density).
//
// Replaces smSmoothStep / smBallSearch / PriorityQueue (pynbody/kdtree/smooth.h:167-242,331-480;
// pq.h:57-204).  Result contract: for every particle the k smallest values of the reference's
// rounded d2 (k counts the particle itself) and h = 0.5*sqrt(d2_k) (smooth.h:429).
//
// Mapping to the GPU: one warp owns one 32-particle bucket of the tree; lane i owns query
// particle i of that bucket.  The warp walks the tree ONCE for all 32 queries -- own bucket first,
// then the sibling sub-trees bottom-up exactly in the reference's search order (smooth.h:209-241) --
// with a per-lane periodic box test (kd.h:68-154) combined by a warp vote.  A candidate bucket is
// staged in shared memory (3 coalesced lines) and every lane measures all 32 candidates against its
// own query; survivors (d2 < current k-th distance) are remembered in a 32-bit register mask and
// then inserted into the lane's private bounded max-heap.  The heaps live in shared memory in
// lane-interleaved layout (entry e of lane l at [e*32+l]: conflict-free), the current k-th distance
// and the query in registers.  Insertions of all lanes run together, so a sift-down step is useful
// work for every lane that has something pending.
#include "traverse.cuh"

namespace pnb {

constexpr int KNN_WARPS = 4;

template <typename T>
struct KnnArgs {
    TreeView<T> tv;
    int k;
    T *smooth_t;            // [n] tree order
    T *rho_t;               // [n] tree order (FUSE_RHO)
    int kernel_id;
    double wendland_self;
    int64_t *nn_idx;        // [n][k] file-order rows, or null
    T *nn_d2;               // [n][k]
};

// replace the root of lane's max-heap (stride-32 layout) by (v,id) and restore heap order
template <typename T, bool WITH_IDX>
__device__ __forceinline__ void heap_replace_top(T *hd, int32_t *hi, int k, T v, int32_t id)
{
    int p = 0;
    for (;;) {
        int c = 2 * p + 1;
        if (c >= k) break;
        T dc = hd[c * 32];
        if (c + 1 < k) {
            T dr = hd[(c + 1) * 32];
            if (dr > dc) { dc = dr; ++c; }
        }
        if (!(v < dc)) break;
        hd[p * 32] = dc;
        if (WITH_IDX) hi[p * 32] = hi[c * 32];
        p = c;
    }
    hd[p * 32] = v;
    if (WITH_IDX) hi[p * 32] = id;
}

template <typename T, bool WITH_IDX, bool PERIODIC>
__device__ __forceinline__ void knn_visit_bucket(const TreeView<T> &tv, int64_t bucket, bool pass, T sx, T sy, T sz,
                                                 T &top, T *hd, int32_t *hi, int k, T *tile, int lane)
{
    const int start = tv.bucket_start[bucket];
    const int cnt = tv.bucket_start[bucket + 1] - start;
    __syncwarp();                                   // previous tile fully consumed
    if (lane < cnt) {
        tile[lane] = tv.x[start + lane];
        tile[32 + lane] = tv.y[start + lane];
        tile[64 + lane] = tv.z[start + lane];
    }
    __syncwarp();
    // phase A: all candidates against this lane's query, branch-free
    unsigned mask = 0;
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
        T d2 = dist2<T>(sx, sy, sz, tile[j], tile[32 + j], tile[64 + j]);
        mask |= (unsigned)(d2 < top) << j;
    }
    if (!pass) mask = 0;
    // phase B: all lanes insert their survivors together
    while (__any_sync(FULL, mask != 0)) {
        if (mask) {
            int j = __ffs(mask) - 1;
            mask &= mask - 1;
            T d2 = dist2<T>(sx, sy, sz, tile[j], tile[32 + j], tile[64 + j]);
            if (d2 < top) {
                heap_replace_top<T, WITH_IDX>(hd, hi, k, d2, start + j);
                top = hd[0];
            }
        }
    }
}

template <typename T, bool WITH_IDX, bool FUSE_RHO, bool PERIODIC>
__global__ void __launch_bounds__(KNN_WARPS * 32) k_knn(KnnArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TreeView<T> &tv = a.tv;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t bucket = (int64_t)blockIdx.x * KNN_WARPS + wib;
    if (bucket >= tv.nsplit) return;                // whole warp exits together; no block barriers below
    const int k = a.k;

    const size_t per_warp = (size_t)k * 32 * sizeof(T) + (WITH_IDX ? (size_t)k * 32 * 4 : 0) + 96 * sizeof(T);
    unsigned char *base = smem_raw + per_warp * wib;
    T *hd = reinterpret_cast<T *>(base) + lane;
    int32_t *hi = WITH_IDX ? reinterpret_cast<int32_t *>(base + (size_t)k * 32 * sizeof(T)) + lane : nullptr;
    T *tile = reinterpret_cast<T *>(base + (size_t)k * 32 * sizeof(T) + (WITH_IDX ? (size_t)k * 32 * 4 : 0));

    const int start = tv.bucket_start[bucket];
    const int cnt = tv.bucket_start[bucket + 1] - start;
    const bool active = lane < cnt;
    T qx = 0, qy = 0, qz = 0;
    if (active) { qx = tv.x[start + lane]; qy = tv.y[start + lane]; qz = tv.z[start + lane]; }
    for (int e = 0; e < k; ++e) {
        hd[e * 32] = Ex<T>::inf();
        if (WITH_IDX) hi[e * 32] = -1;
    }
    // k-th smallest d2 so far; +inf until the heap is full (pq.h:185-192); idle lanes never match
    T top = active ? Ex<T>::inf() : (T)-1;
    const T L = tv.period;

    // own bucket first (smooth.h:200-208): the query lies inside its bounds, so no shift applies
    knn_visit_bucket<T, WITH_IDX, PERIODIC>(tv, bucket, active, qx, qy, qz, top, hd, hi, k, tile, lane);

    // then the sibling sub-tree at every level on the way up (smooth.h:209-241)
    int64_t cell = tv.nsplit + bucket;
    while (cell != 1) {
        int64_t cp = cell ^ 1;
        const int64_t stop = next_node(cp);
        for (;;) {
            Box6<T> b = load_box(tv.box, cp);
            T sx, sy, sz;
            T gap2 = box_gap2<T, PERIODIC>(b, qx, qy, qz, L, sx, sy, sz);
            bool pass = gap2 <= top * Ex<T>::slack();
            if (__any_sync(FULL, pass)) {
                if (cp < tv.nsplit) { cp = 2 * cp; continue; }
                knn_visit_bucket<T, WITH_IDX, PERIODIC>(tv, cp - tv.nsplit, pass, sx, sy, sz, top, hd, hi, k, tile, lane);
            }
            cp = next_node(cp);
            if (cp == stop) break;
        }
        cell >>= 1;
    }

    if (!active) return;
    const int64_t p = start + lane;
    const T h = (T)0.5 * Ex<T>::sqrt(top);          // smooth.h:429
    a.smooth_t[p] = h;

    if (FUSE_RHO) {
        // smDensity (smooth.h:626-647) over the heap: the gather set of radius 2h differs from the
        // heap only by entries at q = 2 where both kernels vanish.
        const T ih = (T)(1.0 / (double)h);
        const T ih2 = ih * ih;
        const T norm = (T)(0.31830988618379067154 * (double)ih * (double)ih2);
        T rho = 0;
        for (int e = 0; e < k; ++e) {
            T q2 = hd[e * 32] * ih2;
            T w = (T)kern_w(a.kernel_id, a.wendland_self, (double)q2);
            w *= norm;
            rho += w * tv.mass[hi[e * 32]];
        }
        a.rho_t[p] = rho;
    }
    if (WITH_IDX && a.nn_idx) {
        const int64_t row = (int64_t)tv.order[p] * k;
        for (int e = 0; e < k; ++e) {
            a.nn_idx[row + e] = (int64_t)tv.order[hi[e * 32]];
            a.nn_d2[row + e] = hd[e * 32];
        }
    }
}

KernelParams make_kernel_params(int kernel_id, int k)
{
    KernelParams kp;
    kp.kernel_id = kernel_id;
    kp.k = k;
    kp.wendland_self = (21 / 16.) * (1 - 0.0294 * pow(k * 0.01, -0.977));
    return kp;
}

template <typename T, bool WITH_IDX, bool FUSE_RHO>
static void launch_knn(Tree &t, KnnArgs<T> &a)
{
    const size_t per_warp = (size_t)a.k * 32 * sizeof(T) + (WITH_IDX ? (size_t)a.k * 32 * 4 : 0) + 96 * sizeof(T);
    const size_t smem = per_warp * KNN_WARPS;
    if (smem > 227 * 1024)
        throw Error(PNB_ERR_VALUE, "number of smoothing neighbours too large for the on-chip neighbour queues (k <= "
                                       + std::to_string((227 * 1024 / KNN_WARPS - 96 * sizeof(T)) / (32 * (sizeof(T) + (WITH_IDX ? 4 : 0))))
                                       + " for this dtype)");
    const unsigned grid = (unsigned)((t.nsplit + KNN_WARPS - 1) / KNN_WARPS);
    if (a.tv.periodic) {
        auto kern = k_knn<T, WITH_IDX, FUSE_RHO, true>;
        PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PNB_LAUNCH(kern, grid, KNN_WARPS * 32, smem, t.stream, a);
    } else {
        auto kern = k_knn<T, WITH_IDX, FUSE_RHO, false>;
        PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PNB_LAUNCH(kern, grid, KNN_WARPS * 32, smem, t.stream, a);
    }
}

template <typename T>
static void run_knn_typed(Tree &t, int k, double period, bool fuse_rho, const KernelParams &kp,
                          int64_t *nn_idx_dev, void *nn_d2_dev)
{
    if (!t.smooth_t.p) t.smooth_t.alloc(sizeof(T) * t.n);
    if (fuse_rho && !t.rho_t.p) t.rho_t.alloc(sizeof(T) * t.n);
    KnnArgs<T> a;
    a.tv = make_view<T>(t, period);
    a.k = k;
    a.smooth_t = t.smooth_t.as<T>();
    a.rho_t = t.rho_t.as<T>();
    a.kernel_id = kp.kernel_id;
    a.wendland_self = kp.wendland_self;
    a.nn_idx = nn_idx_dev;
    a.nn_d2 = (T *)nn_d2_dev;
    if (nn_idx_dev) launch_knn<T, true, false>(t, a);
    else if (fuse_rho) launch_knn<T, true, true>(t, a);
    else launch_knn<T, false, false>(t, a);
}

void run_knn(Tree &t, int k, double period, bool fuse_rho, const KernelParams &kp, int64_t *nn_idx_dev, void *nn_d2_dev)
{
    if (k > t.n)
        throw Error(PNB_ERR_VALUE, "Number of smoothing particles exceeds number of particles in tree");
    if (k < 1) throw Error(PNB_ERR_VALUE, "Number of smoothing particles must be positive");
    if (fuse_rho && !t.has_mass) fuse_rho = false;
    if (t.dtype == PNB_F64) run_knn_typed<double>(t, k, period, fuse_rho, kp, nn_idx_dev, nn_d2_dev);
    else run_knn_typed<float>(t, k, period, fuse_rho, kp, nn_idx_dev, nn_d2_dev);
    t.smooth_valid = true;
    t.smooth_k = k;
    t.smooth_period = period > 0 ? period : 0;
    t.rho_valid = fuse_rho && !nn_idx_dev;
    t.rho_kernel = kp.kernel_id;
}

}  // namespace pnb
