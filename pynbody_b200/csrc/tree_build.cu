// tree_build.cu -- K1: kd-tree build on the GPU.
//
// Replaces kdBuildTree / kdBuildNode / kdSelect / kdUpPass (pynbody/kdtree/kd.cpp:31-238).  The tree is
// an implicit complete binary tree in heap order (root = 1, children 2i / 2i+1, all buckets at one
// depth, kd.h:12-23) over a permutation of the particles; every node splits its particle range along
// the longest axis of its tight bounds.  It is built breadth-first with data-parallel passes instead of
// a recursive quickselect:
//
//   1. three radix sorts give, per axis, the list of particle ids ordered by that coordinate;
//   2. per level, every node reads its tight bounds straight off the ends of its three list
//      segments (no up-pass needed), picks its split axis, and the other two lists are
//      stable-partitioned about the split rank with one global prefix sum;
//   3. after the last level the x-list *is* the tree order; particles are gathered once into
//      tree-ordered structure-of-arrays x[], y[], z[], mass[] so that a 32-particle bucket is
//      32 consecutive elements (one coalesced 128/256-byte line per coordinate).
//
// Two range shapes share the code:
//   PACKED     (the device tree the kernels traverse) node (level, j) owns ranks
//              [j*cap, min(n,(j+1)*cap)), cap = 32 << (levels-level): every bucket holds exactly 32
//              particles (one per warp lane) except the last one; trailing buckets may be empty and
//              carry inverted (+inf,-inf) bounds so no search ever opens them.
//   REFERENCE  (only for KDTree.serialize(), export.cu) the reference's shape: inclusive range [lo,hi]
//              split at m=(lo+hi)/2 (kd.cpp:176-199), leaf count from kdCountNodes (kd.cpp:98-111).
// The device tree takes its split axis from the node's TIGHT box; the reference-shaped tree takes it, like the
// reference (kd.cpp:165-174), from the box the node INHERITED from its parent, so that iDim / fSplit / bounds /
// ranges of the exported KDNode records equal the reference's own.  Neighbour sets do not depend on tree shape.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <algorithm>

namespace pnb {

// ---- index arithmetic of the implicit tree ----------------------------------------------------------
// inclusive rank range [lo,hi] of node (level, j); empty nodes have hi < lo
__host__ __device__ inline void seg_of_node(const TreeShape &s, int level, int64_t j, int64_t &lo, int64_t &hi)
{
    if (s.packed) {
        const int64_t cap = (int64_t)s.leafcap << (s.levels - level);
        lo = j * cap;
        hi = lo + cap - 1;
        if (hi > s.n - 1) hi = s.n - 1;
        if (lo > s.n) lo = s.n;
        return;
    }
    lo = 0; hi = s.n - 1;
    for (int b = level - 1; b >= 0; --b) {
        int64_t m = (lo + hi) >> 1;
        if ((j >> b) & 1) lo = m + 1; else hi = m;
    }
}
// last rank that goes to the lower child of a node owning [lo,hi]
__host__ __device__ inline int64_t split_rank(const TreeShape &s, int level, int64_t lo, int64_t hi)
{
    if (s.packed) {
        const int64_t m = lo + ((int64_t)s.leafcap << (s.levels - level - 1)) - 1;
        return m < hi ? m : hi;
    }
    return (lo + hi) >> 1;
}
__host__ __device__ inline void seg_of_pos(const TreeShape &s, int level, int64_t p, int64_t &lo, int64_t &hi, int64_t &node)
{
    if (s.packed) {                                  // buckets of 32: node ranges are powers of two, no division
        const int sh = 5 + s.levels - level;
        const int64_t cap = (int64_t)1 << sh;
        const int64_t j = p >> sh;
        lo = j << sh;
        hi = lo + cap - 1;
        if (hi > s.n - 1) hi = s.n - 1;
        node = ((int64_t)1 << level) + j;
        return;
    }
    lo = 0; hi = s.n - 1; node = 1;
    for (int i = 0; i < level; ++i) {
        int64_t m = (lo + hi) >> 1;
        if (p <= m) { hi = m; node = 2 * node; } else { lo = m + 1; node = 2 * node + 1; }
    }
}

// order-preserving map float -> unsigned
__device__ inline uint32_t orderable(float v)
{
    uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ inline uint64_t orderable(double v)
{
    uint64_t u = (uint64_t)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

template <typename T, typename K>
__global__ void k_make_keys(const T *__restrict__ c, int64_t n, K *__restrict__ keys, uint32_t *__restrict__ ids)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = orderable(c[i]);
    ids[i] = (uint32_t)i;
}

template <typename T> __device__ inline T pos_inf();
template <> __device__ inline float pos_inf<float>() { return __int_as_float(0x7f800000); }
template <> __device__ inline double pos_inf<double>() { return __longlong_as_double(0x7ff0000000000000ll); }

// one thread per node of `level`: tight bounds from the list ends, split axis, split value
// `ibox` (reference-shaped tree only): the box every node INHERITS from its parent -- the parent's inherited box
// clipped at the parent's split value, the root's being the bounds of all particles -- which is what the reference
// picks its split axis from (kd.cpp:165-174, 182-190); the tight bounds are what it stores afterwards (kdUpPass,
// kd.cpp:68-96).
template <typename T>
__global__ void k_node_bounds(TreeShape s, int level, const uint32_t *__restrict__ L, const T *__restrict__ c,
                              T *__restrict__ box, int8_t *__restrict__ dim_out, T *__restrict__ split_out, bool leaf,
                              int32_t *__restrict__ bucket_start, T *__restrict__ ibox)
{
    const int64_t n = s.n;
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t nn = (int64_t)1 << level;
    if (j >= nn) return;
    int64_t lo, hi;
    seg_of_node(s, level, j, lo, hi);
    int64_t node = nn + j;
    T *b = box + node * 6;
    if (leaf) {
        bucket_start[j] = (int32_t)lo;
        if (j == nn - 1) bucket_start[nn] = (int32_t)n;
    }
    if (hi < lo) {                                   // empty: inverted bounds keep every search out
        const T inf = pos_inf<T>();
        b[0] = b[1] = b[2] = inf; b[3] = b[4] = b[5] = -inf;
        if (!leaf) { dim_out[node] = 0; split_out[node] = 0; }
        return;
    }
    T mn[3], mx[3];
    for (int d = 0; d < 3; ++d) {
        mn[d] = c[d * n + L[d * n + lo]];
        mx[d] = c[d * n + L[d * n + hi]];
    }
    b[0] = mn[0]; b[1] = mn[1]; b[2] = mn[2]; b[3] = mx[0]; b[4] = mx[1]; b[5] = mx[2];
    if (leaf) return;
    int d = 0;
    if (ibox) {
        T *ib = ibox + node * 6;
        if (level == 0) { ib[0] = mn[0]; ib[1] = mn[1]; ib[2] = mn[2]; ib[3] = mx[0]; ib[4] = mx[1]; ib[5] = mx[2]; }
        for (int e = 1; e < 3; ++e)                                   // extents in double, as KDNode::bnd holds them
            if ((double)ib[3 + e] - (double)ib[e] > (double)ib[3 + d] - (double)ib[d]) d = e;
        const T sp = c[d * n + L[d * n + split_rank(s, level, lo, hi)]];
        T *lc = ibox + 2 * node * 6, *uc = lc + 6;
        for (int e = 0; e < 6; ++e) { lc[e] = ib[e]; uc[e] = ib[e]; }
        lc[3 + d] = sp;                                                // kd.cpp:182-190
        uc[d] = sp;
        dim_out[node] = (int8_t)d;
        split_out[node] = sp;
        return;
    }
    for (int e = 1; e < 3; ++e)
        if (mx[e] - mn[e] > mx[d] - mn[d]) d = e;
    dim_out[node] = (int8_t)d;
    split_out[node] = c[d * n + L[d * n + split_rank(s, level, lo, hi)]];
}

// one thread per list position: the particle at rank p of its node's split-axis list goes to the
// lower child iff p <= split rank
constexpr int LEVEL_ITEMS = 4;       // list entries per thread in the per-level passes: four independent load chains in flight

__global__ void k_mark_sides(TreeShape s, int level, const uint32_t *__restrict__ L, const int8_t *__restrict__ dim,
                             uint8_t *__restrict__ side)
{
    const int64_t base = (int64_t)blockIdx.x * blockDim.x * LEVEL_ITEMS + threadIdx.x;
    uint32_t id[LEVEL_ITEMS];
    uint8_t sd[LEVEL_ITEMS];
#pragma unroll
    for (int k = 0; k < LEVEL_ITEMS; ++k) {
        const int64_t p = base + (int64_t)k * blockDim.x;
        id[k] = 0xffffffffu;
        if (p >= s.n) continue;
        int64_t lo, hi, node;
        seg_of_pos(s, level, p, lo, hi, node);
        const int d = dim[node];
        id[k] = L[(int64_t)d * s.n + p];
        sd[k] = (uint8_t)(p > split_rank(s, level, lo, hi));
    }
#pragma unroll
    for (int k = 0; k < LEVEL_ITEMS; ++k)
        if (id[k] != 0xffffffffu) side[id[k]] = sd[k];
}

struct LeftFlag {
    const uint32_t *L;
    const uint8_t *side;
    __host__ __device__ uint32_t operator()(int64_t i) const { return side[L[i]] ? 0u : 1u; }
};

// stable partition of all three lists about each node's split rank, using the global prefix sum S of
// "goes to the lower child" flags over the concatenated lists
__global__ void k_partition(TreeShape s, int level, const uint32_t *__restrict__ L, const uint8_t *__restrict__ side,
                            const int8_t *__restrict__ dim, const uint32_t *__restrict__ S, uint32_t *__restrict__ Lnew)
{
    const int64_t n = s.n;
    const int64_t base = (int64_t)blockIdx.x * blockDim.x * LEVEL_ITEMS + threadIdx.x;
    // phase 1: the independent loads of all items (list entry, node's split axis, prefix sums)
    uint32_t id[LEVEL_ITEMS], sv[LEVEL_ITEMS], sl[LEVEL_ITEMS];
    int64_t lo_[LEVEL_ITEMS], m_[LEVEL_ITEMS], p_[LEVEL_ITEMS];
    int dd[LEVEL_ITEMS];
    bool keep[LEVEL_ITEMS];
#pragma unroll
    for (int k = 0; k < LEVEL_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * blockDim.x;
        dd[k] = -1;
        if (i >= 3 * n) continue;
        const int64_t d = (i >= n) + (i >= 2 * n), p = i - d * n;
        int64_t lo, hi, node;
        seg_of_pos(s, level, p, lo, hi, node);
        dd[k] = (int)d; p_[k] = p; lo_[k] = lo; m_[k] = split_rank(s, level, lo, hi);
        id[k] = L[i];
        keep[k] = s.packed && dim[node] == d;                 // the list of the split axis is partitioned as it stands
        sv[k] = S[i];
        sl[k] = S[d * n + lo];
    }
    // phase 2: the dependent gathers of the side flags
    uint8_t sd[LEVEL_ITEMS];
#pragma unroll
    for (int k = 0; k < LEVEL_ITEMS; ++k) sd[k] = (dd[k] >= 0 && !keep[k]) ? side[id[k]] : 0;
#pragma unroll
    for (int k = 0; k < LEVEL_ITEMS; ++k) {
        if (dd[k] < 0) continue;
        const int64_t i = base + (int64_t)k * blockDim.x;
        if (keep[k]) { Lnew[i] = id[k]; continue; }
        const uint32_t left_before = sv[k] - sl[k];           // modular arithmetic: exact because < 2^32
        const int64_t dest = sd[k] ? (m_[k] + 1) + (p_[k] - lo_[k]) - (int64_t)left_before : lo_[k] + (int64_t)left_before;
        Lnew[(int64_t)dd[k] * n + dest] = id[k];
    }
}

// ---- fused stable partition (packed device tree): one pass per level, no prefix-sum array --------------------
// A tile is PF_TILE consecutive entries of one list.  Node ranges of the packed tree are powers of two and tiles are
// aligned, so a tile either holds whole nodes (range <= PF_TILE: everything is decided inside the CTA) or lies inside
// one node; in that case the number of lower-child entries in the node's earlier tiles comes from a decoupled
// look-back over per-tile status words (aggregate / inclusive prefix; the chain starts at the node's first tile).
// Tiles are handed out by a ticket counter, so every predecessor of a waiting tile is already running.  Status words
// carry the level as a tag: nothing is reset between levels.  Compared with mark + CUB scan over 3n flags + partition
// this reads every list entry and side flag once and writes no 3n-long prefix array (per level 0.45 -> see
// profiles/r2_tree_build_variants.md).
constexpr int PF_THREADS = 256, PF_ITEMS = 8, PF_TILE = PF_THREADS * PF_ITEMS, PF_CHUNKS = PF_TILE / 32;

__device__ __forceinline__ unsigned long long pf_load(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pf_store(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(PF_THREADS)
k_partition_fused(TreeShape s, int level, const uint32_t *__restrict__ L, const uint8_t *__restrict__ side,
                  const int8_t *__restrict__ dim, uint32_t *__restrict__ Lnew, unsigned long long *status,
                  unsigned int *ticket)
{
    __shared__ unsigned s_tile;
    __shared__ uint32_t s_pre[PF_CHUNKS + 1];
    __shared__ uint32_t s_base;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const int64_t n = s.n;
    const int64_t tiles_per_list = (n + PF_TILE - 1) / PF_TILE;
    const int d = (int)(tile / tiles_per_list);
    const int64_t tl = tile - (int64_t)d * tiles_per_list;
    const int64_t p0 = tl * PF_TILE;
    const int sh = 5 + s.levels - level;                    // log2 of the node capacity at this level
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t *Ld = L + (int64_t)d * n;
    uint32_t *Lo = Lnew + (int64_t)d * n;
    const int64_t first_node = (int64_t)1 << level;

    uint32_t id[PF_ITEMS];
    bool keep[PF_ITEMS];
#pragma unroll
    for (int k = 0; k < PF_ITEMS; ++k) {
        const int64_t p = p0 + k * PF_THREADS + threadIdx.x;
        id[k] = 0xffffffffu;
        keep[k] = true;
        if (p < n) {
            id[k] = Ld[p];
            keep[k] = dim[first_node + (p >> sh)] == d;    // the list of the split axis is partitioned as it stands
        }
    }
    bool left[PF_ITEMS];
#pragma unroll
    for (int k = 0; k < PF_ITEMS; ++k) left[k] = (id[k] != 0xffffffffu && !keep[k]) ? side[id[k]] == 0 : false;
    unsigned bal[PF_ITEMS];
#pragma unroll
    for (int k = 0; k < PF_ITEMS; ++k) {
        bal[k] = __ballot_sync(0xffffffffu, left[k]);
        if (lane == 0) s_pre[k * (PF_THREADS / 32) + w + 1] = __popc(bal[k]);
    }
    if (threadIdx.x == 0) s_pre[0] = 0;
    __syncthreads();
    if (w == 0) {                                           // inclusive scan of the PF_CHUNKS chunk counts
        constexpr int PER = PF_CHUNKS / 32;
        uint32_t v[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { v[i] = s_pre[1 + lane * PER + i]; sum += v[i]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        uint32_t run = inc - sum;
#pragma unroll
        for (int i = 0; i < PER; ++i) { run += v[i]; s_pre[1 + lane * PER + i] = run; }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        // a node spanning several tiles: lower-child entries of the node's earlier tiles
        if (((int64_t)1 << sh) > PF_TILE) {
            const bool tile_keeps = dim[first_node + (p0 >> sh)] == d;
            uint32_t base = 0;
            if (!tile_keeps) {
                const unsigned long long tag = (unsigned long long)(level + 1) << 34;
                const int64_t first_tl = ((p0 >> sh) << sh) / PF_TILE;
                unsigned long long *st = status + (int64_t)d * tiles_per_list;
                if (tl == first_tl) {
                    if (lane == 0) pf_store(st + tl, tag | (2ull << 32) | total);
                } else {
                    if (lane == 0) pf_store(st + tl, tag | (1ull << 32) | total);
                    int64_t look = tl - 1;                  // window [look-31, look]
                    for (;;) {
                        const int64_t mine = look - lane;
                        unsigned long long word = tag | (2ull << 32);          // before the node's first tile: prefix 0
                        if (mine >= first_tl) {
                            do { word = pf_load(st + mine); } while ((word >> 34) != (unsigned long long)(level + 1));
                        }
                        const bool incl = ((word >> 32) & 3ull) == 2ull;
                        const unsigned im = __ballot_sync(0xffffffffu, incl);
                        const int stop = im ? __ffs(im) - 1 : 32;             // nearest tile with an inclusive prefix
                        uint32_t add = lane <= stop ? (uint32_t)word : 0u;
#pragma unroll
                        for (int o = 16; o; o >>= 1) add += __shfl_xor_sync(0xffffffffu, add, o);
                        base += add;
                        if (im) break;
                        look -= 32;
                    }
                    if (lane == 0) pf_store(st + tl, tag | (2ull << 32) | (unsigned long long)(base + total));
                }
            }
            if (lane == 0) s_base = base;
        }
    }
    __syncthreads();
    const bool wide = ((int64_t)1 << sh) > PF_TILE;
    const uint32_t base = wide ? s_base : 0u;
    const int64_t cap = (int64_t)1 << sh;
#pragma unroll
    for (int k = 0; k < PF_ITEMS; ++k) {
        if (id[k] == 0xffffffffu) continue;
        const int64_t p = p0 + k * PF_THREADS + threadIdx.x;
        if (keep[k]) { Lo[p] = id[k]; continue; }
        const int c = k * (PF_THREADS / 32) + w;
        const int64_t lo = (p >> sh) << sh;
        int64_t hi = lo + cap - 1;
        if (hi > n - 1) hi = n - 1;
        int64_t m = lo + (cap >> 1) - 1;
        if (m > hi) m = hi;
        uint32_t before = s_pre[c] + __popc(bal[k] & ((1u << lane) - 1u));
        if (wide) before += base;
        else before -= s_pre[(int)((lo - p0) >> 5)];
        const int64_t dest = left[k] ? lo + (int64_t)before : (m + 1) + (p - lo) - (int64_t)before;
        Lo[dest] = id[k];
    }
}

template <typename T>
__global__ void k_gather_tree_order(int64_t n, const uint32_t *__restrict__ order, const T *__restrict__ c,
                                    const T *__restrict__ mass, T *__restrict__ x, T *__restrict__ y,
                                    T *__restrict__ z, T *__restrict__ m)
{
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t id = order[p];
    x[p] = c[id]; y[p] = c[n + id]; z[p] = c[2 * n + id];
    if (mass) m[p] = mass[id];
}

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

template <typename T, typename K>
static void build_typed(Tree &t, DevBuf &c, const T *mass, bool want_soa)
{
    const int64_t n = t.n;
    const TreeShape shape = t.shape;
    const int levels = shape.levels;
    cudaStream_t s = t.stream;
    const int BS = 256;

    DevBuf La(sizeof(uint32_t) * 3 * n), Lb(sizeof(uint32_t) * 3 * n);
    {
        DevBuf keys_in(sizeof(K) * n), keys_out(sizeof(K) * n), ids_in(sizeof(uint32_t) * n);
        size_t tmp_bytes = 0;
        PNB_CUDA((cub::DeviceRadixSort::SortPairs<K, uint32_t, int64_t>(
            nullptr, tmp_bytes, keys_in.as<K>(), keys_out.as<K>(), ids_in.as<uint32_t>(), La.as<uint32_t>(), n, 0,
            (int)sizeof(K) * 8, s)));
        DevBuf tmp(tmp_bytes);
        for (int d = 0; d < 3; ++d) {
            PNB_LAUNCH((k_make_keys<T, K>), nblk(n, BS), BS, 0, s, c.as<T>() + d * n, n, keys_in.as<K>(),
                       ids_in.as<uint32_t>());
            PNB_CUDA((cub::DeviceRadixSort::SortPairs<K, uint32_t, int64_t>(
                tmp.p, tmp_bytes, keys_in.as<K>(), keys_out.as<K>(), ids_in.as<uint32_t>(),
                La.as<uint32_t>() + d * n, n, 0, (int)sizeof(K) * 8, s)));
            g_launches += 4;   // CUB's histogram + onesweep passes (approximate count)
        }
        PNB_CUDA(cudaStreamSynchronize(s));            // keys/tmp go back to the block cache on scope exit
    }

    t.box.alloc(sizeof(T) * 6 * 2 * t.nsplit);
    if (!(shape.packed && want_soa && g_opt_tree_build == 1)) t.bucket_start.alloc(sizeof(int32_t) * (t.nsplit + 1));
    DevBuf dim(sizeof(int8_t) * 2 * t.nsplit), split(sizeof(T) * 2 * t.nsplit);
    // packed device tree: fused one-pass partition per level; reference-shaped tree (arbitrary ranges) and
    // pnb_set_option("tree_build", 3): flags -> CUB prefix sum over the three lists -> partition
    const bool fused = shape.packed && g_opt_tree_build != 3;
    DevBuf side(sizeof(uint8_t) * n), S(fused ? 0 : sizeof(uint32_t) * 3 * n);
    const int64_t pf_tiles = 3 * ((n + PF_TILE - 1) / PF_TILE);
    DevBuf pf_state;
    if (fused) {
        pf_state.alloc(sizeof(unsigned long long) * (pf_tiles + 32));
        PNB_CUDA(cudaMemsetAsync(pf_state.p, 0, pf_state.bytes, s));
    }
    DevBuf inherited;                                  // reference-shaped tree: split axis from the inherited boxes
    if (!shape.packed) inherited.alloc(sizeof(T) * 6 * 2 * t.nsplit);
    PNB_CUDA(cudaMemsetAsync(t.box.p, 0, sizeof(T) * 6 * 2 * t.nsplit, s));

    uint32_t *L = La.as<uint32_t>(), *Ln = Lb.as<uint32_t>();
    size_t scan_bytes = 0;
    if (!fused) {
        LeftFlag f{L, side.as<uint8_t>()};
        auto it = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), f);
        PNB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, it, S.as<uint32_t>(), 3 * n, s));
    }
    DevBuf scan_tmp(scan_bytes);

    // The packed device tree finishes its lowest levels per sub-tree in shared memory (tree_morton.cu: exact
    // medians by an in-CTA sort, tight bounds, tree-ordered structure of arrays) instead of one global pass +
    // prefix sum per level; the reference-shaped tree (export) runs every level here.
    const int lseg = (shape.packed && want_soa && g_opt_tree_build == 1) ? std::min(levels, segment_levels(t.dtype)) : 0;
    const int global_levels = levels - lseg;
    for (int level = 0; level < global_levels; ++level) {
        int64_t nn = (int64_t)1 << level;
        PNB_LAUNCH((k_node_bounds<T>), nblk(nn, 128), 128, 0, s, shape, level, L, c.as<T>(), t.box.as<T>(),
                   dim.as<int8_t>(), split.as<T>(), false, (int32_t *)nullptr, inherited.as<T>());
        PNB_LAUNCH(k_mark_sides, nblk(n, BS * LEVEL_ITEMS), BS, 0, s, shape, level, L, dim.as<int8_t>(), side.as<uint8_t>());
        if (fused) {
            // status words [pf_tiles], then one ticket counter per level (level < 32)
            unsigned long long *status = pf_state.as<unsigned long long>();
            unsigned int *ticket = reinterpret_cast<unsigned int *>(status + pf_tiles) + level;
            PNB_LAUNCH(k_partition_fused, (unsigned)pf_tiles, PF_THREADS, 0, s, shape, level, L, side.as<uint8_t>(),
                       dim.as<int8_t>(), Ln, status, ticket);
        } else {
            LeftFlag f{L, side.as<uint8_t>()};
            auto it = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), f);
            PNB_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp.p, scan_bytes, it, S.as<uint32_t>(), 3 * n, s));
            g_launches += 2;
            PNB_LAUNCH(k_partition, nblk(3 * n, BS * LEVEL_ITEMS), BS, 0, s, shape, level, L, side.as<uint8_t>(), dim.as<int8_t>(), S.as<uint32_t>(), Ln);
        }
        uint32_t *tmp = L; L = Ln; Ln = tmp;
    }
    if (lseg > 0) {
        build_tree_segments(t, c.p, mass, L, lseg);
    } else {
        PNB_LAUNCH((k_node_bounds<T>), nblk(t.nsplit, 128), 128, 0, s, shape, levels, L, c.as<T>(), t.box.as<T>(),
                   dim.as<int8_t>(), split.as<T>(), true, t.bucket_start.as<int32_t>(), (T *)nullptr);
        t.order.alloc(sizeof(uint32_t) * n);
        PNB_CUDA(cudaMemcpyAsync(t.order.p, L, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s));
        if (want_soa) {                                    // tree-ordered structure of arrays
            t.x.alloc(sizeof(T) * n); t.y.alloc(sizeof(T) * n); t.z.alloc(sizeof(T) * n);
            if (mass) t.mass.alloc(sizeof(T) * n);
            t.has_mass = mass != nullptr;
            if (t.mass_ready) PNB_CUDA(cudaStreamWaitEvent(s, t.mass_ready, 0));   // masses uploaded behind the sorts
            PNB_LAUNCH((k_gather_tree_order<T>), nblk(n, BS), BS, 0, s, n, t.order.as<uint32_t>(), c.as<T>(), mass,
                       t.x.as<T>(), t.y.as<T>(), t.z.as<T>(), t.mass.as<T>());
        }
    }
    T rb[6];
    PNB_CUDA(cudaMemcpyAsync(rb, t.box.as<T>() + 6, sizeof(rb), cudaMemcpyDeviceToHost, s));
    PNB_CUDA(cudaStreamSynchronize(s));
    for (int d = 0; d < 3; ++d) { t.root_min[d] = (double)rb[d]; t.root_max[d] = (double)rb[3 + d]; }
    t.split_dim = std::move(dim);
    t.split_val = std::move(split);
}

// packed device tree: 32-particle buckets
TreeShape packed_shape(int64_t n)
{
    TreeShape s;
    s.n = n; s.packed = 1; s.leafcap = 32; s.levels = 0;
    const int64_t nb = (n + 31) / 32;
    while (((int64_t)1 << s.levels) < nb) ++s.levels;
    return s;
}
// the reference's shape for a given leafsize (kd.cpp:98-111)
TreeShape reference_shape(int64_t n, int leafsize)
{
    TreeShape s;
    s.n = n; s.packed = 0; s.leafcap = leafsize; s.levels = 0;
    int64_t m = n;
    while (m > leafsize) { m >>= 1; ++s.levels; }
    return s;
}

void build_tree(Tree &t, DevBuf &cols, const void *mass_dev, const TreeShape &shape, bool want_soa)
{
    if (t.n < 1) throw Error(PNB_ERR_VALUE, "kd-tree needs at least one particle");
    if (t.n >= ((int64_t)1 << 31)) throw Error(PNB_ERR_VALUE, "at most 2^31-1 particles per device tree");
    t.shape = shape;
    t.levels = shape.levels;
    t.nsplit = (int64_t)1 << shape.levels;
    if (t.dtype == PNB_F64)
        build_typed<double, uint64_t>(t, cols, (const double *)mass_dev, want_soa);
    else
        build_typed<float, uint32_t>(t, cols, (const float *)mass_dev, want_soa);
}

// host-side copy of the index arithmetic for export.cu
void shape_seg_of_node(const TreeShape &s, int level, int64_t j, int64_t &lo, int64_t &hi) { seg_of_node(s, level, j, lo, hi); }

}  // namespace pnb
