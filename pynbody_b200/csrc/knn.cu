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
// then inserted into the lane's private bounded neighbour queue, all lanes inserting together.
//
// The queue (the reference's max-heap, pq.h) is a "grouped max-cache": the k distances live in
// shared memory, lane-interleaved (entry e of lane l at [e*32+l]: conflict-free), split into 8 groups
// of S = ceil(k/8) slots; the maximum of every group, its slot, and the overall maximum (the current
// k-th distance) live in registers.  Replacing the current maximum is one store plus one rescan of a
// single group: S independent shared-memory loads (one latency) instead of the log2(k) dependent
// load/compare/store rounds of a binary heap sift-down -- the v1 profile (profiles/r1_knn_v1_*)
// showed that sift-down as 45 % of all issue slots at 9 of 32 lanes active.
#include "traverse.cuh"

namespace pnb {

constexpr int KNN_WARPS = 2;
constexpr int NGROUP = 8;

template <typename T>
struct KnnArgs {
    TreeView<T> tv;
    int k;
    int S;                  // slots per group, ceil(k / NGROUP)
    T *smooth_t;            // [n] tree order
    T *rho_t;               // [n] tree order (FUSE_RHO: all particles have mass `uniform_mass`)
    double uniform_mass;
    int kernel_id;
    double wendland_self;
    int64_t *nn_idx;        // [n][k] file-order rows, or null
    T *nn_d2;               // [n][k]
};

template <typename T>
struct TopK {
    T gm[NGROUP];           // maximum of each group
    unsigned long long gp;  // slot (within its group) of each group's maximum, 8 bits per group
    T top;                  // overall maximum = k-th smallest d2 so far (+inf until k are held, pq.h:185-192)
};

// replace the current maximum by (v, id)
template <typename T, bool WITH_IDX>
__device__ __forceinline__ void topk_replace_max(TopK<T> &q, T *hd, int32_t *hi, int S, T v, int32_t id)
{
    int g = 0;
#pragma unroll
    for (int i = NGROUP - 1; i >= 0; --i)
        if (q.gm[i] == q.top) g = i;
    const int s = (int)((q.gp >> (8 * g)) & 0xffull);
    T *grp = hd + (size_t)g * S * 32;
    grp[s * 32] = v;
    if (WITH_IDX) hi[(g * S + s) * 32] = id;
    // rescan the group: S independent loads
    T m = grp[0];
    int pos = 0;
#pragma unroll 4
    for (int t = 1; t < S; ++t) {
        T val = grp[t * 32];
        if (val > m) { m = val; pos = t; }
    }
#pragma unroll
    for (int i = 0; i < NGROUP; ++i)
        if (i == g) q.gm[i] = m;
    q.gp = (q.gp & ~(0xffull << (8 * g))) | ((unsigned long long)pos << (8 * g));
    T a0 = q.gm[0] > q.gm[1] ? q.gm[0] : q.gm[1], a1 = q.gm[2] > q.gm[3] ? q.gm[2] : q.gm[3];
    T a2 = q.gm[4] > q.gm[5] ? q.gm[4] : q.gm[5], a3 = q.gm[6] > q.gm[7] ? q.gm[6] : q.gm[7];
    a0 = a0 > a1 ? a0 : a1;
    a2 = a2 > a3 ? a2 : a3;
    q.top = a0 > a2 ? a0 : a2;
}

template <typename T, bool WITH_IDX>
__device__ __forceinline__ void knn_visit_bucket(const TreeView<T> &tv, int64_t bucket, bool pass, T sx, T sy, T sz,
                                                 TopK<T> &q, T *hd, int32_t *hi, int S, T *tile, int lane)
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
    const T top0 = q.top;
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
        T d2 = dist2<T>(sx, sy, sz, tile[j], tile[32 + j], tile[64 + j]);
        mask |= (unsigned)(d2 < top0) << j;
    }
    if (!pass) mask = 0;
    // phase B: all lanes insert their survivors together
    while (__any_sync(FULL, mask != 0)) {
        if (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const T d2 = dist2<T>(sx, sy, sz, tile[j], tile[32 + j], tile[64 + j]);
            if (d2 < q.top) topk_replace_max<T, WITH_IDX>(q, hd, hi, S, d2, start + j);
        }
    }
}

template <typename T, bool WITH_IDX>
__host__ __device__ inline size_t knn_smem_per_warp(int S)
{
    return (size_t)NGROUP * S * 32 * sizeof(T) + (WITH_IDX ? (size_t)NGROUP * S * 32 * 4 : 0) + 96 * sizeof(T);
}

template <typename T, bool WITH_IDX, bool FUSE_RHO, bool PERIODIC>
__global__ void __launch_bounds__(KNN_WARPS * 32) k_knn(KnnArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TreeView<T> &tv = a.tv;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t bucket = (int64_t)blockIdx.x * KNN_WARPS + wib;
    if (bucket >= tv.nsplit) return;                // whole warp exits together; no block barriers below
    const int k = a.k, S = a.S, nslot = NGROUP * S;

    unsigned char *base = smem_raw + knn_smem_per_warp<T, WITH_IDX>(S) * wib;
    T *hd = reinterpret_cast<T *>(base) + lane;
    int32_t *hi = WITH_IDX ? reinterpret_cast<int32_t *>(base + (size_t)nslot * 32 * sizeof(T)) + lane : nullptr;
    T *tile = reinterpret_cast<T *>(base + (size_t)nslot * 32 * sizeof(T) + (WITH_IDX ? (size_t)nslot * 32 * 4 : 0));

    const int start = tv.bucket_start[bucket];
    const int cnt = tv.bucket_start[bucket + 1] - start;
    const bool active = lane < cnt && (!tv.query_mask || tv.query_mask[start + lane]);
    if (!__any_sync(FULL, active)) return;          // no query of this bucket is selected
    T qx = 0, qy = 0, qz = 0;
    if (active) { qx = tv.x[start + lane]; qy = tv.y[start + lane]; qz = tv.z[start + lane]; }

    // slots [0,k) start at +inf (replaced first), padding slots [k, 8S) at -inf (never the maximum)
    for (int e = 0; e < nslot; ++e) {
        hd[e * 32] = e < k ? Ex<T>::inf() : -Ex<T>::inf();
        if (WITH_IDX) hi[e * 32] = -1;
    }
    TopK<T> q;
#pragma unroll
    for (int g = 0; g < NGROUP; ++g) q.gm[g] = (g * S < k) ? Ex<T>::inf() : -Ex<T>::inf();
    q.gp = 0;
    q.top = active ? Ex<T>::inf() : (T)-1;          // idle lanes never match
    const T L = tv.period;

    // own bucket first (smooth.h:200-208): the query lies inside its bounds, so no shift applies
    knn_visit_bucket<T, WITH_IDX>(tv, bucket, active, qx, qy, qz, q, hd, hi, S, tile, lane);

    // then the sibling sub-tree at every level on the way up (smooth.h:209-241)
    int64_t cell = tv.nsplit + bucket;
    while (cell != 1) {
        int64_t cp = cell ^ 1;
        const int64_t stop = next_node(cp);
        for (;;) {
            Box6<T> b = load_box(tv.box, cp);
            T sx, sy, sz;
            T gap2 = box_gap2<T, PERIODIC>(b, qx, qy, qz, L, sx, sy, sz);
            bool pass = gap2 <= q.top * Ex<T>::slack() && gap2 < Ex<T>::inf();   // (empty nodes: inverted bounds)
            if (__any_sync(FULL, pass)) {
                if (cp < tv.nsplit) { cp = 2 * cp; continue; }
                knn_visit_bucket<T, WITH_IDX>(tv, cp - tv.nsplit, pass, sx, sy, sz, q, hd, hi, S, tile, lane);
            }
            cp = next_node(cp);
            if (cp == stop) break;
        }
        cell >>= 1;
    }

    if (!active) return;
    const int64_t p = start + lane;
    const T h = (T)0.5 * Ex<T>::sqrt(q.top);        // smooth.h:429
    a.smooth_t[p] = h;

    if (FUSE_RHO) {
        // smDensity (smooth.h:626-647) over the queue: the gather set of radius 2h differs from the
        // k nearest only by entries at q = 2 where both kernels vanish.  All masses are equal here.
        const T ih = (T)(1.0 / (double)h);
        const T ih2 = ih * ih;
        const T norm = (T)(0.31830988618379067154 * (double)ih * (double)ih2);
        const T m = (T)a.uniform_mass;
        T rho = 0;
        for (int e = 0; e < k; ++e) {
            T q2 = hd[e * 32] * ih2;
            T w = (T)kern_w(a.kernel_id, a.wendland_self, (double)q2);
            w *= norm;
            rho += w * m;
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
    const size_t smem = knn_smem_per_warp<T, WITH_IDX>(a.S) * KNN_WARPS;
    if (smem > 227 * 1024 || a.S > 255)
        throw Error(PNB_ERR_VALUE, "number of smoothing neighbours too large for the on-chip neighbour queues");
    const unsigned grid = (unsigned)((t.nsplit + KNN_WARPS - 1) / KNN_WARPS);
    KernelTimer timer(0, t.stream);
    if (a.tv.periodic) {
        auto kern = k_knn<T, WITH_IDX, FUSE_RHO, true>;
        PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PNB_LAUNCH(kern, grid, KNN_WARPS * 32, smem, t.stream, a);
    } else {
        auto kern = k_knn<T, WITH_IDX, FUSE_RHO, false>;
        PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PNB_LAUNCH(kern, grid, KNN_WARPS * 32, smem, t.stream, a);
    }
    timer.stop();
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
    a.S = (k + NGROUP - 1) / NGROUP;
    a.smooth_t = t.smooth_t.as<T>();
    a.rho_t = t.rho_t.as<T>();
    a.uniform_mass = t.uniform_mass;
    a.kernel_id = kp.kernel_id;
    a.wendland_self = kp.wendland_self;
    a.nn_idx = nn_idx_dev;
    a.nn_d2 = (T *)nn_d2_dev;
    if (nn_idx_dev) launch_knn<T, true, false>(t, a);
    else if (fuse_rho) launch_knn<T, false, true>(t, a);
    else launch_knn<T, false, false>(t, a);
}

void run_knn(Tree &t, int k, double period, bool fuse_rho, const KernelParams &kp, int64_t *nn_idx_dev, void *nn_d2_dev)
{
    if (k > t.n)
        throw Error(PNB_ERR_VALUE, "Number of smoothing particles exceeds number of particles in tree");
    if (k < 1) throw Error(PNB_ERR_VALUE, "Number of smoothing particles must be positive");
    // the density can ride along only when every particle has the same mass (no neighbour ids are kept)
    if (fuse_rho && !(t.has_mass && t.mass_is_uniform)) fuse_rho = false;
    if (nn_idx_dev) fuse_rho = false;
    if (t.dtype == PNB_F64) run_knn_typed<double>(t, k, period, fuse_rho, kp, nn_idx_dev, nn_d2_dev);
    else run_knn_typed<float>(t, k, period, fuse_rho, kp, nn_idx_dev, nn_d2_dev);
    t.smooth_valid = true;                  // (for the particles selected by the query mask, if one is set)
    t.smooth_k = k;
    t.smooth_period = period > 0 ? period : 0;
    t.rho_valid = fuse_rho;
    t.rho_kernel = kp.kernel_id;
}

}  // namespace pnb
