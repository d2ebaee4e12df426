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

// ---- float64 coordinates: sort on 32-bit keys, then put the (rare) ties right ---------------------------------
// A 64-bit radix sort takes eight passes.  key32 = floor((v - min) * 2^32 / (max - min)) is a monotone map of the
// coordinate (each step is monotone under rounding), so a four-pass sort on it orders everything except particles that
// share a key -- closer than 2^-32 of the extent -- and those runs are finished by rank counting on the exact
// (coordinate, id) pairs: the result is the list the stable 64-bit sort gives.  Long runs (coordinates on a lattice)
// raise a flag and that axis falls back to the 64-bit sort.
constexpr int TIE_RUN_MAX = 48;

__global__ void k_minmax_orderable(const double *__restrict__ c, int64_t n, unsigned long long *__restrict__ mm)
{
    unsigned long long lo = ~0ull, hi = 0ull;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long u = orderable(c[i]);
        lo = u < lo ? u : lo;
        hi = u > hi ? u : hi;
    }
    for (int o = 16; o; o >>= 1) {
        const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(mm, lo); atomicMax(mm + 1, hi); }
}

__device__ inline double from_orderable(unsigned long long u)
{
    u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)u);
}

__global__ void k_make_keys32(const double *__restrict__ c, int64_t n, const unsigned long long *__restrict__ mm,
                              uint32_t *__restrict__ keys, uint32_t *__restrict__ ids, int *__restrict__ flag)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mn = from_orderable(mm[0]), mx = from_orderable(mm[1]);
    const double ext = mx - mn;
    const double v = c[i];
    uint32_t key = 0;
    if (ext > 0 && ext < 1.7e308 && v == v) {
        double t = (v - mn) * (4294967295.0 / ext);          // monotone in v; <= 2^32 - 1 up to rounding
        t = t < 0.0 ? 0.0 : (t > 4294967295.0 ? 4294967295.0 : t);
        key = (uint32_t)t;                                     // truncation: monotone
    } else if (!(v == v) || !(ext < 1.7e308)) {
        *flag = 1;                                             // NaN / infinite extent: leave it to the 64-bit sort
    }
    keys[i] = key;
    ids[i] = (uint32_t)i;
}

__global__ void k_fix_ties(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ ids, const double *__restrict__ c,
                           int64_t n, uint32_t *__restrict__ out, int *__restrict__ flag)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t k = keys[p], id = ids[p];
    const bool le = p > 0 && keys[p - 1] == k, re = p + 1 < n && keys[p + 1] == k;
    if (!le && !re) { out[p] = id; return; }
    int64_t a = p, b = p;
    while (a > 0 && keys[a - 1] == k && p - a <= TIE_RUN_MAX) --a;
    while (b + 1 < n && keys[b + 1] == k && b - p <= TIE_RUN_MAX) ++b;
    if (b - a + 1 > TIE_RUN_MAX) { *flag = 1; return; }
    const unsigned long long me = orderable(c[id]);
    int rank = 0;
    for (int64_t q = a; q <= b; ++q) {
        const uint32_t oid = ids[q];
        const unsigned long long o = orderable(c[oid]);
        rank += (o < me) || (o == me && oid < id);
    }
    out[a + rank] = id;
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
    // only the upper halves are read and written: the flag of a particle is "went to the upper child AT THIS LEVEL"
    // (side[] starts at 0xff; lower-half particles keep an older level's tag)
    uint32_t id[LEVEL_ITEMS];
#pragma unroll
    for (int k = 0; k < LEVEL_ITEMS; ++k) {
        const int64_t p = base + (int64_t)k * blockDim.x;
        id[k] = 0xffffffffu;
        if (p >= s.n) continue;
        int64_t lo, hi, node;
        seg_of_pos(s, level, p, lo, hi, node);
        if (!(p > split_rank(s, level, lo, hi))) continue;
        id[k] = L[(int64_t)dim[node] * s.n + p];
    }
#pragma unroll
    for (int k = 0; k < LEVEL_ITEMS; ++k)
        if (id[k] != 0xffffffffu) side[id[k]] = (uint8_t)level;
}

struct LeftFlag {
    const uint32_t *L;
    const uint8_t *side;
    uint8_t level;
    __host__ __device__ uint32_t operator()(int64_t i) const { return side[L[i]] == level ? 0u : 1u; }
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
    for (int k = 0; k < LEVEL_ITEMS; ++k) sd[k] = (dd[k] >= 0 && !keep[k]) ? (uint8_t)(side[id[k]] == (uint8_t)level) : 0;
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

template <int MINB>
__global__ void __launch_bounds__(PF_THREADS, MINB)
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
    for (int k = 0; k < PF_ITEMS; ++k) left[k] = (id[k] != 0xffffffffu && !keep[k]) ? side[id[k]] != (uint8_t)level : false;
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

// ---- lowest levels on chip, from the presorted lists (packed device tree) --------------------------------------
// Once a node holds at most PS_SEG particles its three list segments fit into shared memory.  One CTA then finishes
// the sub-tree: particles get LOCAL ids (their rank in the segment's x-list; global id -> local id through one
// scatter / two gathers of a scratch array) and every further level -- bounds off the list ends, longest axis, side
// flags, stable partition of the other two lists -- runs on 16-bit local ids in shared memory with the same index
// arithmetic as the global passes, so the tree is the SAME tree.  Only the list ends touch the coordinates (a few
// hundred gathers per segment); the kernel also writes the leaf bounds and the particle order, replacing the
// per-level passes of these levels and the leaf-bounds pass.
constexpr int PS_LEVELS = 6, PS_SEG = 32 << PS_LEVELS, PS_THREADS = 256, PS_ITEMS = PS_SEG / PS_THREADS,
              PS_CHUNKS = PS_SEG / 32;

static size_t partition_segment_smem()
{
    return (size_t)PS_SEG * (sizeof(uint32_t) + 6 * sizeof(uint16_t) + 1) + 3 * (PS_CHUNKS + 1) * sizeof(uint32_t)
           + PS_CHUNKS * sizeof(int) + 64;
}

template <typename T>
__global__ void __launch_bounds__(PS_THREADS, 4)
k_partition_segment(TreeShape s, int ls, const uint32_t *__restrict__ L, const T *__restrict__ c,
                    uint16_t *__restrict__ loc, T *__restrict__ box, int8_t *__restrict__ dim_out, T *__restrict__ split_out,
                    int32_t *__restrict__ bucket_start, uint32_t *__restrict__ order)
{
    extern __shared__ __align__(16) unsigned char ps_raw[];
    const int64_t n = s.n;
    const int SEG = 32 << ls;                       // particles per segment (ls <= PS_LEVELS)
    const int gl = s.levels - ls;                   // level of this CTA's root node
    const int64_t seg = blockIdx.x, lo0 = seg * SEG;
    const int cnt = (int)max((int64_t)0, min((int64_t)SEG, n - lo0));
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t nsplit = (int64_t)1 << s.levels;

    uint32_t *gid = reinterpret_cast<uint32_t *>(ps_raw);                   // [PS_SEG] global id by local id
    uint16_t *la = reinterpret_cast<uint16_t *>(gid + PS_SEG);              // [3][PS_SEG] lists of local ids
    uint16_t *lb = la + 3 * PS_SEG;                                         // the other buffer
    uint32_t *pre = reinterpret_cast<uint32_t *>(lb + 3 * PS_SEG);          // [3][PS_CHUNKS + 1] chunk prefix sums
    int *ndim = reinterpret_cast<int *>(pre + 3 * (PS_CHUNKS + 1));         // split axis of every node of the level
    uint8_t *side = reinterpret_cast<uint8_t *>(ndim + PS_CHUNKS);          // [PS_SEG] by local id

    // local ids = ranks in the x-list (all loads of a thread in flight together)
    {
        uint32_t g[PS_ITEMS];
#pragma unroll
        for (int k = 0; k < PS_ITEMS; ++k) {
            const int i = k * PS_THREADS + tid;
            g[k] = i < cnt ? L[lo0 + i] : 0u;
        }
#pragma unroll
        for (int k = 0; k < PS_ITEMS; ++k) {
            const int i = k * PS_THREADS + tid;
            if (i < cnt) { gid[i] = g[k]; loc[g[k]] = (uint16_t)i; la[i] = (uint16_t)i; }
        }
    }
    __syncthreads();                                // (block-wide visibility of the scratch writes)
    {
        uint32_t gy[PS_ITEMS], gz[PS_ITEMS];
#pragma unroll
        for (int k = 0; k < PS_ITEMS; ++k) {
            const int i = k * PS_THREADS + tid;
            gy[k] = i < cnt ? L[n + lo0 + i] : 0u;
            gz[k] = i < cnt ? L[2 * n + lo0 + i] : 0u;
        }
        uint16_t ly[PS_ITEMS], lz[PS_ITEMS];
#pragma unroll
        for (int k = 0; k < PS_ITEMS; ++k) {
            const int i = k * PS_THREADS + tid;
            ly[k] = i < cnt ? loc[gy[k]] : (uint16_t)0;
            lz[k] = i < cnt ? loc[gz[k]] : (uint16_t)0;
        }
#pragma unroll
        for (int k = 0; k < PS_ITEMS; ++k) {
            const int i = k * PS_THREADS + tid;
            if (i < cnt) { la[PS_SEG + i] = ly[k]; la[2 * PS_SEG + i] = lz[k]; }
        }
    }
    __syncthreads();

    const T inf = pos_inf<T>();
    for (int l = 0; l <= ls; ++l) {
        const int sh = 5 + ls - l;                  // log2 of the node capacity at this level
        const int cap = 1 << sh, nn = 1 << l;
        const bool leaf = l == ls;
        // bounds straight off the list ends, split axis and value (k_node_bounds): six lanes per node, one list end each
        for (int t = tid; t < nn * 8; t += PS_THREADS) {
            const int j = t >> 3, r = t & 7;
            const int lo = j << sh, hi = min(cnt, lo + cap) - 1;
            const int64_t node = ((int64_t)1 << (gl + l)) + seg * nn + j;
            T *b = box + node * 6;
            const bool empty = hi < lo;
            T v = r < 3 ? inf : -inf;
            if (!empty && r < 6) {
                const int d = r < 3 ? r : r - 3;
                v = c[(int64_t)d * n + gid[la[d * PS_SEG + (r < 3 ? lo : hi)]]];
            }
            if (r < 6) b[r] = v;
            // the eight lanes of a node share its bounds by shuffle (nodes never straddle a warp: 8 divides 32)
            const unsigned grp = 0xffu << (lane & 24);
            T mn[3], mx[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                mn[d] = __shfl_sync(grp, v, (lane & 24) + d);
                mx[d] = __shfl_sync(grp, v, (lane & 24) + 3 + d);
            }
            if (r == 0) {
                if (leaf) {
                    const int64_t bj = seg * nn + j;
                    bucket_start[bj] = (int32_t)min(n, lo0 + lo);
                    if (bj == nsplit - 1) bucket_start[nsplit] = (int32_t)n;
                } else if (empty) {
                    dim_out[node] = 0; split_out[node] = 0; ndim[j] = 0;
                } else {
                    int d = 0;
                    for (int e = 1; e < 3; ++e)
                        if (mx[e] - mn[e] > mx[d] - mn[d]) d = e;
                    const int mr = min(lo + (cap >> 1) - 1, hi);
                    dim_out[node] = (int8_t)d;
                    split_out[node] = c[(int64_t)d * n + gid[la[d * PS_SEG + mr]]];
                    ndim[j] = d;
                }
            }
        }
        if (leaf) break;
        __syncthreads();
        // side flags from the split-axis list (k_mark_sides)
        for (int p = tid; p < cnt; p += PS_THREADS) {
            const int j = p >> sh, lo = j << sh, hi = min(cnt, lo + cap) - 1;
            const int mr = min(lo + (cap >> 1) - 1, hi);
            side[la[ndim[j] * PS_SEG + p]] = (uint8_t)(p > mr);
        }
        __syncthreads();
        // stable partition of the other two lists (k_partition_fused, whole nodes inside the CTA)
        uint16_t id[3][PS_ITEMS];
        unsigned bal[3][PS_ITEMS];
        bool left[3][PS_ITEMS];
#pragma unroll
        for (int e = 0; e < 3; ++e)
#pragma unroll
            for (int k = 0; k < PS_ITEMS; ++k) {
                const int p = k * PS_THREADS + tid;
                id[e][k] = 0; left[e][k] = false;
                if (p < cnt) {
                    id[e][k] = la[e * PS_SEG + p];
                    left[e][k] = ndim[p >> sh] != e && side[id[e][k]] == 0;
                }
                bal[e][k] = __ballot_sync(0xffffffffu, left[e][k]);
                if (lane == 0) pre[e * (PS_CHUNKS + 1) + k * (PS_THREADS / 32) + w + 1] = __popc(bal[e][k]);
            }
        if (tid < 3) pre[tid * (PS_CHUNKS + 1)] = 0;
        __syncthreads();
        if (w < 3) {                                // warp e: inclusive scan of list e's chunk counts
            constexpr int PER = PS_CHUNKS / 32;
            uint32_t *pp = pre + w * (PS_CHUNKS + 1) + 1;
            uint32_t v[PER], sum = 0;
#pragma unroll
            for (int i = 0; i < PER; ++i) { v[i] = pp[lane * PER + i]; sum += v[i]; }
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += up;
            }
            uint32_t run = inc - sum;
#pragma unroll
            for (int i = 0; i < PER; ++i) { run += v[i]; pp[lane * PER + i] = run; }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 3; ++e)
#pragma unroll
            for (int k = 0; k < PS_ITEMS; ++k) {
                const int p = k * PS_THREADS + tid;
                if (p >= cnt) continue;
                const int j = p >> sh;
                if (ndim[j] == e) { lb[e * PS_SEG + p] = id[e][k]; continue; }
                const int lo = j << sh, hi = min(cnt, lo + cap) - 1;
                const int mr = min(lo + (cap >> 1) - 1, hi);
                const uint32_t *pp = pre + e * (PS_CHUNKS + 1);
                const int before = (int)(pp[k * (PS_THREADS / 32) + w] - pp[lo >> 5]) + __popc(bal[e][k] & ((1u << lane) - 1u));
                const int dest = left[e][k] ? lo + before : (mr + 1) + (p - lo) - before;
                lb[e * PS_SEG + dest] = id[e][k];
            }
        __syncthreads();
        uint16_t *t2 = la; la = lb; lb = t2;
    }
    __syncthreads();
    // tree order = the x-list
    for (int p = tid; p < cnt; p += PS_THREADS) order[lo0 + p] = gid[la[p]];
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

// float64: the three lists through the 32-bit-key sort; done[d] is set for every axis that needed no fallback
static void presort_f64(DevBuf &c, int64_t n, DevBuf &keys_in, DevBuf &keys_out, DevBuf &ids_in, DevBuf &tmp, size_t tmp_bytes,
                        DevBuf &La, cudaStream_t s, bool done[3])
{
    const int BS = 256;
    DevBuf ids_out(sizeof(uint32_t) * n), aux(64);
    // aux: [0..5] min / max per axis (orderable), then three flags
    unsigned long long init[8] = {~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull, 0ull, 0ull};
    PNB_CUDA(cudaMemcpyAsync(aux.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
    unsigned long long *mm = aux.as<unsigned long long>();
    int *flags = reinterpret_cast<int *>(mm + 6);
    size_t need = 0;
    PNB_CUDA((cub::DeviceRadixSort::SortPairs<uint32_t, uint32_t, int64_t>(
        nullptr, need, keys_in.as<uint32_t>(), keys_out.as<uint32_t>(), ids_in.as<uint32_t>(), ids_out.as<uint32_t>(), n, 0, 32, s)));
    if (need > tmp_bytes) return;                           // (never: the 64-bit sort needs more)
    for (int d = 0; d < 3; ++d) {
        const double *cd = c.as<double>() + (int64_t)d * n;
        PNB_LAUNCH(k_minmax_orderable, 148 * 8, BS, 0, s, cd, n, mm + 2 * d);
        PNB_LAUNCH(k_make_keys32, nblk(n, BS), BS, 0, s, cd, n, mm + 2 * d, keys_in.as<uint32_t>(), ids_in.as<uint32_t>(), flags + d);
        PNB_CUDA((cub::DeviceRadixSort::SortPairs<uint32_t, uint32_t, int64_t>(
            tmp.p, need, keys_in.as<uint32_t>(), keys_out.as<uint32_t>(), ids_in.as<uint32_t>(), ids_out.as<uint32_t>(), n, 0, 32, s)));
        PNB_LAUNCH(k_fix_ties, nblk(n, BS), BS, 0, s, keys_out.as<uint32_t>(), ids_out.as<uint32_t>(), cd, n,
                   La.as<uint32_t>() + (int64_t)d * n, flags + d);
        g_launches += 3;
    }
    int hflags[3];
    PNB_CUDA(cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, s));
    PNB_CUDA(cudaStreamSynchronize(s));
    for (int d = 0; d < 3; ++d) done[d] = hflags[d] == 0;
}

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
        bool presorted[3] = {false, false, false};
        if (sizeof(K) == 8 && g_opt_tree_build != 3) presort_f64(c, n, keys_in, keys_out, ids_in, tmp, tmp_bytes, La, s, presorted);
        for (int d = 0; d < 3; ++d) {
            if (presorted[d]) continue;
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
    PNB_CUDA(cudaMemsetAsync(side.p, 0xff, sizeof(uint8_t) * n, s));      // "never went up" (levels are < 64)

    uint32_t *L = La.as<uint32_t>(), *Ln = Lb.as<uint32_t>();
    size_t scan_bytes = 0;
    if (!fused) {
        LeftFlag f{L, side.as<uint8_t>(), 0};
        auto it = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), f);
        PNB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, it, S.as<uint32_t>(), 3 * n, s));
    }
    DevBuf scan_tmp(scan_bytes);

    // The packed device tree finishes its lowest levels per sub-tree in shared memory (tree_morton.cu: exact
    // medians by an in-CTA sort, tight bounds, tree-ordered structure of arrays) instead of one global pass +
    // prefix sum per level; the reference-shaped tree (export) runs every level here.
    const int lseg = (shape.packed && want_soa && g_opt_tree_build == 1) ? std::min(levels, segment_levels(t.dtype)) : 0;
    // default (2): the lowest PS_LEVELS levels from the presorted lists in shared memory (k_partition_segment);
    // 4: every level by the fused global pass
    const int lps = (fused && want_soa && g_opt_tree_build == 2) ? std::min(levels, PS_LEVELS) : 0;
    const int global_levels = levels - lseg - lps;
    for (int level = 0; level < global_levels; ++level) {
        int64_t nn = (int64_t)1 << level;
        PNB_LAUNCH((k_node_bounds<T>), nblk(nn, 128), 128, 0, s, shape, level, L, c.as<T>(), t.box.as<T>(),
                   dim.as<int8_t>(), split.as<T>(), false, (int32_t *)nullptr, inherited.as<T>());
        PNB_LAUNCH(k_mark_sides, nblk(n, BS * LEVEL_ITEMS), BS, 0, s, shape, level, L, dim.as<int8_t>(), side.as<uint8_t>());
        if (fused) {
            // status words [pf_tiles], then one ticket counter per level (level < 32)
            unsigned long long *status = pf_state.as<unsigned long long>();
            unsigned int *ticket = reinterpret_cast<unsigned int *>(status + pf_tiles) + level;
            // six resident CTAs per SM (40 registers): the pass is latency-bound (4 CTAs: 11.75 ms per build, 6 or 8: 10.78 ms)
            PNB_LAUNCH(k_partition_fused<6>, (unsigned)pf_tiles, PF_THREADS, 0, s, shape, level, L, side.as<uint8_t>(),
                       dim.as<int8_t>(), Ln, status, ticket);
        } else {
            LeftFlag f{L, side.as<uint8_t>(), (uint8_t)level};
            auto it = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), f);
            PNB_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp.p, scan_bytes, it, S.as<uint32_t>(), 3 * n, s));
            g_launches += 2;
            PNB_LAUNCH(k_partition, nblk(3 * n, BS * LEVEL_ITEMS), BS, 0, s, shape, level, L, side.as<uint8_t>(), dim.as<int8_t>(), S.as<uint32_t>(), Ln);
        }
        uint32_t *tmp = L; L = Ln; Ln = tmp;
    }
    if (lseg > 0) {
        build_tree_segments(t, c.p, mass, L, lseg);
    } else if (lps > 0 || (fused && want_soa && g_opt_tree_build == 2)) {
        t.order.alloc(sizeof(uint32_t) * n);
        {
            DevBuf loc(sizeof(uint16_t) * n);
            auto kern = k_partition_segment<T>;
            const size_t smem = partition_segment_smem();
            PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PNB_LAUNCH(kern, (unsigned)(t.nsplit >> lps), PS_THREADS, smem, s, shape, lps, L, c.as<T>(), loc.as<uint16_t>(),
                       t.box.as<T>(), dim.as<int8_t>(), split.as<T>(), t.bucket_start.as<int32_t>(), t.order.as<uint32_t>());
            t.x.alloc(sizeof(T) * n); t.y.alloc(sizeof(T) * n); t.z.alloc(sizeof(T) * n);
            if (mass) t.mass.alloc(sizeof(T) * n);
            t.has_mass = mass != nullptr;
            if (t.mass_ready) PNB_CUDA(cudaStreamWaitEvent(s, t.mass_ready, 0));   // masses uploaded behind the sorts
            PNB_LAUNCH((k_gather_tree_order<T>), nblk(n, BS), BS, 0, s, n, t.order.as<uint32_t>(), c.as<T>(), mass,
                       t.x.as<T>(), t.y.as<T>(), t.z.as<T>(), t.mass.as<T>());
            PNB_CUDA(cudaStreamSynchronize(s));            // `loc` goes back to the block cache on scope exit
        }
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
