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
// staged in shared memory (3 coalesced lines).  The first k candidates of a search go straight into
// the queue's leaves.  After that the bucket is tested with lane = candidate: the queries in reach of
// it are taken in turn, their (shifted) position and k-th distance broadcast by shuffle, and the 32
// verdicts of one query come back as one ballot; the survivors (d2 < current k-th distance) of each
// query are then inserted into its lane's private neighbour queue, all lanes inserting together.
// The kernel is persistent: warps take buckets in tree order from a global counter.
// (-DKNN_STATS builds a variant that counts buckets, survivors, insert rounds and node tests.)
//
// NEIGHBOUR LISTS.  With KEEP_IDS the id of every queued neighbour is written straight to its final place in
// the tree's resident neighbour list (Tree::nn_ids, uint32 tree positions, laid out [bucket][slot][lane] so that
// a warp reads or writes one 128-byte line per slot): the queue's slot e of lane l IS list entry e of particle
// 32*bucket+l.  The ids never enter shared memory (which bounds occupancy), cost one fire-and-forget global
// store per insertion, and stay in L2 while a bucket is being worked on.  Every later reduction over the same
// neighbours (density with unequal masses, mean, dispersion, div, curl; gather_list.cu) and nn() read these
// lists instead of walking the tree again.
//
// The queue (the reference's max-heap, pq.h) is a TOURNAMENT TREE per lane in shared memory,
// lane-interleaved (element e of lane l at [e*32+l]: conflict-free): K2 = 2^ceil(log2 k) leaf slots
// hold the keys, K2-1 one-byte internal nodes hold the leaf index of the larger child's winner; the
// root's winner is the current k-th distance and is kept in registers.  Replacing the maximum touches
// one root-to-leaf path whose addresses depend only on the leaf index, so all sibling loads of the path
// are issued together (two shared-memory round trips) and the rest is a register compare chain
// (profiles/r1_knn_v1_summary.txt, r1_knn_v2_summary.txt for the binary heap / grouped max-cache it
// replaced).
//
// LAYOUT (k <= 128, i.e. tree depth known at compile time).  ONE block per SM with as many warps as its shared memory holds
// (21 for k = 64): [candidate tiles of the first warps | leaves of every warp | winner bytes of every warp | remaining tiles].
// The leaves start at a multiple of one warp's leaf region (K2 * 128 bytes) of the shared window -- the gap in front of them
// holds candidate tiles -- and the winner bytes of a warp at a multiple of K2 * 32, in IN-ORDER positions ((a << l) | (1 << (l-1))
// for the node above leaves [a 2^l, (a+1) 2^l)) and swizzled [pos / 4][lane][pos % 4] (lane = bank): every node and sibling
// address on a replacement path is then ONE logic instruction away from the leaf index, and the accesses go through 32-bit
// shared-window addresses (queue_replace_max_fast; 113 SASS instructions per insert round against 140 with heap positions
// and generic pointers).  Larger k keeps the generic formulation (heap positions, two warps per block).
// PERIODIC boxes: wrapped images are tested until the queues are full; afterwards a warp none of whose query balls reaches a
// face of the period runs the plain cell test.  Bucket ranges are computed (bucket j = ranks [32 j, 32 j + 32)), and the bounds /
// coordinates the walk will need next are prefetched to L1 by otherwise idle lanes.
//
// float64 positions: shared memory, not arithmetic, bounds the number of resident warps, and latency
// hiding is what the kernel lives on.  The queue therefore keeps only the HIGH 32 bits of each d2 in
// shared memory (non-negative doubles order like their bit patterns); the low words sit in an
// L2-resident scratch array in global memory and are read only for the new maximum and to break ties
// between equal high words.  Every ordering decision is still made on the exact 64-bit value.
#include "traverse.cuh"
#include <cmath>
#include <limits>

namespace pnb {
#ifdef KNN_STATS
__device__ unsigned long long g_stats[8];
#define STAT(i, v) do { if (lane == 0) atomicAdd(&g_stats[i], (unsigned long long)(v)); } while (0)
#else
#define STAT(i, v)
#endif

constexpr int KNN_WARPS = 2;          // k > 128: warps per block (several blocks per SM)
constexpr int KNN_MAX_WARPS = 21;     // k <= 128: ONE block per SM with as many warps as its shared memory holds
constexpr int MAXLOG = 10;              // k <= 1024

template <typename T>
struct KnnArgs {
    TreeView<T> tv;
    int k;
    int LOG;                // log2(K2), K2 = leaf slots of the tournament tree
    T *smooth_t;            // [n] tree order
    T *rho_t;               // [n] tree order (FUSE_RHO: all particles have mass `uniform_mass`)
    double uniform_mass;
    int kernel_id;
    double wendland_self;
    uint32_t *nn_ids;       // [nsplit][K2][32] neighbour list (KEEP_IDS), tree positions; padding slots hold ~0u
    T *nn_d2;               // same layout: the queued distances, written once at the end of a search, or null
    unsigned long long *next_bucket;   // work counter
    uint32_t *lo_scratch;   // [resident warps][K2][32] low words of the queued d2 (float64 only)
    // periodic boxes: a ball that stays inside (face_lo, face_hi) on every axis cannot contain a wrapped image of any
    // particle (all particles lie within one period: face_lo = root_max - L rounded up, face_hi = root_min + L rounded
    // down); face_eps covers the rounding of q +- r
    T face_lo[3], face_hi[3], face_eps;
};

// key held in shared memory for a queued distance
template <typename T> struct KeyOf { typedef T type; };
template <> struct KeyOf<double> { typedef uint32_t type; };

// per-lane view of the queue
template <typename T, typename W_t>
struct Queue {
    typedef typename KeyOf<T>::type Key;
    Key *hd;                // leaves [K2], stride 32
    uint32_t *lo;           // low words [K2], stride 32, global memory (float64)
    W_t *W;                 // internal nodes [K2] (heap order, node 0 unused), stride 32: winner leaf index
    uint32_t *hi;           // neighbour ids [K2], stride 32, GLOBAL memory: the tree's resident list (KEEP_IDS)
    int LOG, K2;
    T top;                  // float: the root winner's value = k-th smallest d2 so far.  double: an UPPER BOUND of it
                            // (exact high word, low word all ones) -- the exact low word is fetched only on demand
    uint32_t top_hi, top_lo;  // double: exact high word; low word if lo_known
    bool lo_known;
    int etop;               // the root winner's leaf
    uint32_t hb, wb;        // k <= 128: 32-bit shared-window addresses of this lane's leaf 0 / winner byte 0; both regions
                            // are aligned to their size, so any leaf or node address is one logic operation away from
                            // the leaf index (queue_replace_max_fast)
};

// ---- internal nodes of the tournament --------------------------------------------------------------------------
// The node at height l (1 = parents of leaves) above leaves [a 2^l, (a+1) 2^l) sits at
//   k <= 128 (depth known at compile time): in-order position (a << l) | (1 << (l-1))
//   otherwise:                               heap position (K2 >> l) + a
template <int LOGC>
__device__ __forceinline__ int wpos(int K2, int l, int a)
{
    return LOGC >= 0 ? ((a << l) | (1 << (l - 1))) : ((K2 >> l) + a);
}
// byte offset of the winner byte at position pos from the lane's byte 0 (Queue::W).  Generic depth: [pos][lane].
// Compile-time depth: four consecutive positions of one lane share a 32-bit word, [pos / 4][lane][pos % 4], so lane l
// only ever touches bank l and the byte accesses of a replacement are free of bank conflicts whatever leaves the
// lanes are working on ([pos][lane] put four lanes into each word: 1.7 G conflicts in 7 G wavefronts, ncu).
template <int LOGC>
__device__ __forceinline__ int wofs(int pos)
{
    return LOGC >= 0 ? (((pos >> 2) << 7) | (pos & 3)) : pos * 32;
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }

// shared memory through 32-bit window addresses (no generic-address arithmetic in the insert path)
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

// is d2 strictly below the current k-th distance?  (exact)
template <typename W_t>
__device__ __forceinline__ bool below_top(Queue<float, W_t> &q, float d2) { return d2 < q.top; }
template <typename W_t>
__device__ __forceinline__ bool below_top(Queue<double, W_t> &q, double d2)
{
    const uint32_t h = (uint32_t)__double2hiint(d2);
    if (h != q.top_hi) return h < q.top_hi;         // non-negative doubles order like their bit patterns
    if (!q.lo_known) { q.top_lo = q.lo[q.etop * 32]; q.lo_known = true; }   // equal high words: rare
    return (uint32_t)__double2loint(d2) < q.top_lo;
}
// phase-A filter, evaluated by every lane for the query of lane i: can d2 be below that query's k-th distance?
// (double: decided on the high words -- one shuffle instead of two, an integer compare; an upper bound of the exact test)
template <typename W_t>
__device__ __forceinline__ bool maybe_below(const Queue<float, W_t> &q, float d2, int i) { return d2 < __shfl_sync(FULL, q.top, i); }
template <typename W_t>
__device__ __forceinline__ bool maybe_below(const Queue<double, W_t> &q, double d2, int i)
{
    return (uint32_t)__double2hiint(d2) <= __shfl_sync(FULL, q.top_hi, i);
}
template <typename W_t>
__device__ __forceinline__ float top_exact(Queue<float, W_t> &q) { return q.top; }
template <typename W_t>
__device__ __forceinline__ double top_exact(Queue<double, W_t> &q)
{
    if (!q.lo_known) { q.top_lo = q.lo[q.etop * 32]; q.lo_known = true; }
    return __hiloint2double((int)q.top_hi, (int)q.top_lo);
}

__device__ __forceinline__ float queue_value(const Queue<float, uint8_t> &q, int e) { return q.hd[e * 32]; }
__device__ __forceinline__ float queue_value(const Queue<float, uint16_t> &q, int e) { return q.hd[e * 32]; }
template <typename W_t>
__device__ __forceinline__ double queue_value(const Queue<double, W_t> &q, int e)
{
    return __hiloint2double((int)q.hd[e * 32], (int)q.lo[e * 32]);
}

// replace the current maximum by (v, id) and replay its path to the root
// LOGC >= 0: tree depth known at compile time (k <= 32 / 64 / 128: the path is fully unrolled without
// predicates); LOGC < 0: depth read from the queue (any k <= 1024)
template <typename W_t, int LOGC, bool WITH_IDX>
__device__ __forceinline__ void queue_replace_max(Queue<float, W_t> &q, float v, uint32_t id)
{
    constexpr int NL = LOGC >= 0 ? LOGC : MAXLOG;
    const int e = q.etop;
    q.hd[e * 32] = v;
    if (WITH_IDX) q.hi[e * 32] = id;
    const int K2c = LOGC >= 0 ? (1 << LOGC) : q.K2;              // heap index of leaf e is K2c + e
    int si[NL > 0 ? NL : 1];
    float sv[NL > 0 ? NL : 1];
    // sibling winners along the path: addresses depend only on e, so all loads go out together
    if (NL > 0) si[0] = e ^ 1;
#pragma unroll
    for (int l = 1; l < NL; ++l)
        if (LOGC >= 0 || l < q.LOG) si[l] = (int)q.W[((K2c >> l) + ((e >> l) ^ 1)) * 32];
#pragma unroll
    for (int l = 0; l < NL; ++l)
        if (LOGC >= 0 || l < q.LOG) sv[l] = q.hd[si[l] * 32];
    float cv = v;
    int ce = e;
#pragma unroll
    for (int l = 0; l < NL; ++l)
        if (LOGC >= 0 || l < q.LOG) {
            if (sv[l] > cv) { cv = sv[l]; ce = si[l]; }
            q.W[((K2c >> (l + 1)) + (e >> (l + 1))) * 32] = (W_t)ce;
        }
    q.top = cv;
    q.etop = ce;
}

template <typename W_t, int LOGC, bool WITH_IDX>
__device__ __forceinline__ void queue_replace_max(Queue<double, W_t> &q, double v, uint32_t id)
{
    constexpr int NL = LOGC >= 0 ? LOGC : MAXLOG;
    const int e = q.etop;
    const uint32_t vhi = (uint32_t)__double2hiint(v), vlo = (uint32_t)__double2loint(v);
    q.hd[e * 32] = vhi;
    q.lo[e * 32] = vlo;
    if (WITH_IDX) q.hi[e * 32] = id;
    const int K2c = LOGC >= 0 ? (1 << LOGC) : q.K2;
    int si[NL > 0 ? NL : 1], win[NL > 0 ? NL : 1];
    uint32_t sk[NL > 0 ? NL : 1];
    if (NL > 0) si[0] = e ^ 1;
#pragma unroll
    for (int l = 1; l < NL; ++l)
        if (LOGC >= 0 || l < q.LOG) si[l] = (int)q.W[((K2c >> l) + ((e >> l) ^ 1)) * 32];
#pragma unroll
    for (int l = 0; l < NL; ++l)
        if (LOGC >= 0 || l < q.LOG) sk[l] = q.hd[si[l] * 32];
    // the path is replayed on the high words alone, without branches; equal high words anywhere on the
    // path (rare) send the lane through the exact replay below
    uint32_t ck = vhi;
    int ce = e;
    bool tie = false;
#pragma unroll
    for (int l = 0; l < NL; ++l)
        if (LOGC >= 0 || l < q.LOG) {
            const bool gt = sk[l] > ck;
            tie |= sk[l] == ck;
            ck = gt ? sk[l] : ck;
            ce = gt ? si[l] : ce;
            win[l] = ce;
        }
    uint32_t clo = vlo;
    bool clo_known = ce == e;                       // low word of the path winner is in a register
    if (tie) {
        ck = vhi; clo = vlo; ce = e; clo_known = true;
#pragma unroll
        for (int l = 0; l < NL; ++l)
            if (LOGC >= 0 || l < q.LOG) {
                bool gt = sk[l] > ck;
                if (sk[l] == ck) {                  // decide on the exact 64-bit values
                    if (!clo_known) { clo = q.lo[ce * 32]; clo_known = true; }
                    gt = q.lo[si[l] * 32] > clo;
                }
                if (gt) { ck = sk[l]; ce = si[l]; clo_known = false; }
                win[l] = ce;
            }
    }
#pragma unroll
    for (int l = 0; l < NL; ++l)
        if (LOGC >= 0 || l < q.LOG) q.W[((K2c >> (l + 1)) + (e >> (l + 1))) * 32] = (W_t)win[l];
    q.top_hi = ck;
    q.top_lo = clo;
    q.lo_known = clo_known;
    q.top = __hiloint2double((int)ck, -1);          // upper bound for pruning and the phase-A filter
    q.etop = ce;
}

// ---- k <= 128: the same replacement with the address arithmetic stripped to one instruction per access ----------
// Round-2 profile of the insert round (profiles/r2_knn_summary.txt): 140 SASS instructions, a third of them address
// arithmetic (heap-position shifts per level, the shared window base rematerialised in every round).  Here the lane's
// leaves live at hb | e*128 and its winner bytes at wb | pos*32 with pos in in-order layout, both regions aligned to
// their size: the path node at height l is (e32 & ~mask_l) | wb (+ a constant folded into the access), its sibling is
// that address with bit l flipped.
template <int LOGC>
struct PathAddr {
    // swizzled offset of position p (compile-time): wofs<LOGC>(p)
    static __device__ __forceinline__ constexpr uint32_t sw(uint32_t p) { return ((p >> 2) << 7) | (p & 3u); }
    uint32_t t[LOGC > 0 ? LOGC : 1];         // t[l-1]: lane's winner byte 0 | swizzled first leaf under the path node at height l
    __device__ __forceinline__ PathAddr(uint32_t wb, uint32_t e)
    {
        const uint32_t S = ((e & ~3u) << 5) | (e & 3u);                 // sw(e)
#pragma unroll
        for (int l = 1; l <= LOGC; ++l) t[l - 1] = (S & ~sw((1u << l) - 1u)) | wb;
    }
    // (the constants are folded into the accesses' immediate offsets)
    __device__ __forceinline__ uint32_t own(int l) const { return t[l - 1] + sw(1u << (l - 1)); }
    __device__ __forceinline__ uint32_t sibling(int l) const { return (t[l - 1] ^ sw(1u << l)) + sw(1u << (l - 1)); }   // l < LOGC
};

template <int LOGC, bool WITH_IDX>
__device__ __forceinline__ void queue_replace_max_fast(Queue<float, uint8_t> &q, float v, uint32_t id)
{
    const uint32_t e = (uint32_t)q.etop;
    const uint32_t leaf = q.hb | (e << 7);
    sts_u32(leaf, __float_as_uint(v));
    if (WITH_IDX) q.hi[e * 32] = id;
    const PathAddr<LOGC> pa(q.wb, e);
    uint32_t si[LOGC > 0 ? LOGC : 1];
    float sv[LOGC > 0 ? LOGC : 1];
    si[0] = e ^ 1u;
#pragma unroll
    for (int l = 1; l < LOGC; ++l) si[l] = lds_u8(pa.sibling(l));
    if (LOGC > 0) sv[0] = __uint_as_float(lds_u32(leaf ^ 128u));
#pragma unroll
    for (int l = 1; l < LOGC; ++l) sv[l] = __uint_as_float(lds_u32(q.hb + (si[l] << 7)));
    float cv = v;
    uint32_t ce = e;
#pragma unroll
    for (int l = 0; l < LOGC; ++l) {
        if (sv[l] > cv) { cv = sv[l]; ce = si[l]; }
        sts_u8(pa.own(l + 1), ce);
    }
    q.top = cv;
    q.etop = (int)ce;
}

template <int LOGC, bool WITH_IDX>
__device__ __forceinline__ void queue_replace_max_fast(Queue<double, uint8_t> &q, double v, uint32_t id)
{
    const uint32_t e = (uint32_t)q.etop;
    const uint32_t vhi = (uint32_t)__double2hiint(v), vlo = (uint32_t)__double2loint(v);
    const uint32_t leaf = q.hb | (e << 7);
    sts_u32(leaf, vhi);
    q.lo[e * 32] = vlo;
    if (WITH_IDX) q.hi[e * 32] = id;
    const PathAddr<LOGC> pa(q.wb, e);
    uint32_t si[LOGC > 0 ? LOGC : 1], win[LOGC > 0 ? LOGC : 1], sk[LOGC > 0 ? LOGC : 1];
    si[0] = e ^ 1u;
#pragma unroll
    for (int l = 1; l < LOGC; ++l) si[l] = lds_u8(pa.sibling(l));
    if (LOGC > 0) sk[0] = lds_u32(leaf ^ 128u);
#pragma unroll
    for (int l = 1; l < LOGC; ++l) sk[l] = lds_u32(q.hb + (si[l] << 7));
    // replay on the high words alone, without branches; equal high words anywhere on the path (rare) send the lane
    // through the exact replay below
    uint32_t ck = vhi, ce = e;
    bool tie = false;
#pragma unroll
    for (int l = 0; l < LOGC; ++l) {
        const bool gt = sk[l] > ck;
        tie |= sk[l] == ck;
        ck = gt ? sk[l] : ck;
        ce = gt ? si[l] : ce;
        win[l] = ce;
    }
    q.lo_known = false;                             // (fetched on demand: below_top on equal high words, top_exact)
    if (tie) {
        uint32_t clo = vlo;
        bool clo_known = true;
        ck = vhi; ce = e;
#pragma unroll
        for (int l = 0; l < LOGC; ++l) {
            bool gt = sk[l] > ck;
            if (sk[l] == ck) {                      // decide on the exact 64-bit values
                if (!clo_known) { clo = q.lo[ce * 32]; clo_known = true; }
                gt = q.lo[si[l] * 32] > clo;
            }
            if (gt) { ck = sk[l]; ce = si[l]; clo_known = false; }
            win[l] = ce;
        }
    }
#pragma unroll
    for (int l = 0; l < LOGC; ++l) sts_u8(pa.own(l + 1), win[l]);
    q.top_hi = ck;
    q.top = __hiloint2double((int)ck, -1);          // upper bound for pruning (kept in registers: deriving it from top_hi
                                                    // at the points of use makes nvcc guard every warp vote with BRA.DIV)
    q.etop = (int)ce;
}

// depth known at compile time (k <= 128): the stripped path; otherwise the generic one
template <typename T, typename W_t, int LOGC, bool WITH_IDX>
struct Replace {
    static_assert(LOGC < 0, "compile-time depths use the in-order winner layout (queue_replace_max_fast)");
    static __device__ __forceinline__ void run(Queue<T, W_t> &q, T v, uint32_t id) { queue_replace_max<W_t, LOGC, WITH_IDX>(q, v, id); }
};
template <typename T, bool WITH_IDX> struct Replace<T, uint8_t, 5, WITH_IDX> {
    static __device__ __forceinline__ void run(Queue<T, uint8_t> &q, T v, uint32_t id) { queue_replace_max_fast<5, WITH_IDX>(q, v, id); }
};
template <typename T, bool WITH_IDX> struct Replace<T, uint8_t, 6, WITH_IDX> {
    static __device__ __forceinline__ void run(Queue<T, uint8_t> &q, T v, uint32_t id) { queue_replace_max_fast<6, WITH_IDX>(q, v, id); }
};
template <typename T, bool WITH_IDX> struct Replace<T, uint8_t, 7, WITH_IDX> {
    static __device__ __forceinline__ void run(Queue<T, uint8_t> &q, T v, uint32_t id) { queue_replace_max_fast<7, WITH_IDX>(q, v, id); }
};

// ---- start of a search: the first k candidates go straight into the leaves, then the tree is built once ----
__device__ __forceinline__ void leaf_store(Queue<float, uint8_t> &q, int e, float d2) { q.hd[e * 32] = d2; }
__device__ __forceinline__ void leaf_store(Queue<float, uint16_t> &q, int e, float d2) { q.hd[e * 32] = d2; }
template <typename W_t>
__device__ __forceinline__ void leaf_store(Queue<double, W_t> &q, int e, double d2)
{
    q.hd[e * 32] = (uint32_t)__double2hiint(d2);
    q.lo[e * 32] = (uint32_t)__double2loint(d2);
}
// does leaf a hold a larger value than leaf b?  (exact)
__device__ __forceinline__ bool leaf_greater(const Queue<float, uint8_t> &q, int a, int b) { return q.hd[a * 32] > q.hd[b * 32]; }
__device__ __forceinline__ bool leaf_greater(const Queue<float, uint16_t> &q, int a, int b) { return q.hd[a * 32] > q.hd[b * 32]; }
template <typename W_t>
__device__ __forceinline__ bool leaf_greater(const Queue<double, W_t> &q, int a, int b)
{
    const uint32_t ka = q.hd[a * 32], kb = q.hd[b * 32];
    if (ka != kb) return ka > kb;
    return q.lo[a * 32] > q.lo[b * 32];
}
__device__ __forceinline__ void set_top_from_leaf(Queue<float, uint8_t> &q, int e) { q.top = q.hd[e * 32]; q.etop = e; }
__device__ __forceinline__ void set_top_from_leaf(Queue<float, uint16_t> &q, int e) { q.top = q.hd[e * 32]; q.etop = e; }
template <typename W_t>
__device__ __forceinline__ void set_top_from_leaf(Queue<double, W_t> &q, int e)
{
    q.top_hi = q.hd[e * 32];
    q.lo_known = false;
    q.top = __hiloint2double((int)q.top_hi, -1);
    q.etop = e;
}

// play every match of the tournament once (K2 - 1 comparisons); leaves [k, K2) hold the smallest key
template <typename T, typename W_t, int LOGC>
__device__ __forceinline__ void queue_build(Queue<T, W_t> &q)
{
    const int K2 = LOGC >= 0 ? (1 << LOGC) : q.K2;
    const int LOG = LOGC >= 0 ? LOGC : q.LOG;
    if (K2 == 1) { set_top_from_leaf(q, 0); return; }
#pragma unroll 4
    for (int a = 0; a < K2 / 2; ++a) {                      // parents of leaves
        const int x = 2 * a, y = x + 1;
        q.W[wofs<LOGC>(wpos<LOGC>(K2, 1, a))] = (W_t)(leaf_greater(q, y, x) ? y : x);
    }
    for (int l = 2; l <= LOG; ++l) {
#pragma unroll 4
        for (int a = 0; a < (K2 >> l); ++a) {
            const int x = (int)q.W[wofs<LOGC>(wpos<LOGC>(K2, l - 1, 2 * a))], y = (int)q.W[wofs<LOGC>(wpos<LOGC>(K2, l - 1, 2 * a + 1))];
            q.W[wofs<LOGC>(wpos<LOGC>(K2, l, a))] = (W_t)(leaf_greater(q, y, x) ? y : x);
        }
    }
    set_top_from_leaf(q, (int)q.W[wofs<LOGC>(wpos<LOGC>(K2, LOG, 0))]);
}

// cell test with the periodic wrap switched off at run time (warp-uniform) once no query ball of the warp can reach
// across a face of the period
template <typename T, bool PERIODIC>
__device__ __forceinline__ T box_gap2_w(bool wrap, const Box6<T> &b, T qx, T qy, T qz, T L, T &sx, T &sy, T &sz)
{
    if (PERIODIC && wrap) return box_gap2<T, true>(b, qx, qy, qz, L, sx, sy, sz);
    return box_gap2<T, false>(b, qx, qy, qz, L, sx, sy, sz);
}

template <typename T, typename W_t, int LOGC, bool WITH_IDX>
__device__ __forceinline__ void knn_visit_bucket(const TreeView<T> &tv, int64_t bucket, bool pass, T sx, T sy, T sz,
                                                 Queue<T, W_t> &q, T *tile, int lane, bool wide, T qx, T qy, T qz,
                                                 int &filled, int k, bool active)
{
    // every builder makes bucket j the ranks [32 j, 32 j + 32) cut at n (bucket_start[j] = min(n, 32 j)): computed, not
    // loaded -- a dependent global load less in front of every candidate tile
    const int start = (int)min(tv.n, (int64_t)32 * bucket);
    const int cnt = (int)min(tv.n, (int64_t)32 * (bucket + 1)) - start;
    __syncwarp();                                   // previous tile fully consumed
    if (lane < cnt) {
        tile[lane] = tv.x[start + lane];
        tile[32 + lane] = tv.y[start + lane];
        tile[64 + lane] = tv.z[start + lane];
    }
    __syncwarp();
    unsigned done = 0;                              // candidates consumed by the fill below
    if (filled >= 0) {
        // the queue is not full yet (warp-uniform: nothing is pruned while the k-th distance is infinite, so
        // every lane has seen the same candidates): store the distances into the next leaves
        const int take = min(cnt, k - filled);
        for (int j = 0; j < take; ++j) {
            const T d2 = wide ? dist2_nearest<T>(qx, qy, qz, tile[j], tile[32 + j], tile[64 + j], tv.period)
                              : dist2<T>(sx, sy, sz, tile[j], tile[32 + j], tile[64 + j]);
            leaf_store(q, filled + j, d2);
            if (WITH_IDX) q.hi[(filled + j) * 32] = (uint32_t)(start + j);
        }
        filled += take;
        if (filled < k) return;
        queue_build<T, W_t, LOGC>(q);
        filled = -1;
        if (!active) { q.top = (T)-1; q.top_hi = 0u; q.top_lo = 0u; q.lo_known = true; }   // idle lanes never match
        if (take == cnt) return;
        done = (1u << take) - 1u;
    }
    // phase A, transposed: lane j holds candidate j; the queries the bucket is in reach of (typically a third
    // of the lanes) are taken in turn, their shifted position and k-th distance broadcast by shuffle, and the
    // 32 verdicts of one query come back as one ballot -- half the instructions of testing all 32 x 32 pairs
    unsigned mask = 0;
    const T top0 = q.top;
    if (!wide) {
        const T cx = lane < cnt ? tile[lane] : (T)0, cy = lane < cnt ? tile[32 + lane] : (T)0,
                cz = lane < cnt ? tile[64 + lane] : (T)0;
        unsigned todo = __ballot_sync(FULL, pass);
        const bool valid = lane < cnt;
        while (todo) {
            const int i = __ffs(todo) - 1;
            todo &= todo - 1;
            const T ux = __shfl_sync(FULL, sx, i), uy = __shfl_sync(FULL, sy, i), uz = __shfl_sync(FULL, sz, i);
            const T d2 = dist2<T>(ux, uy, uz, cx, cy, cz);
            const bool mb = maybe_below(q, d2, i);                 // (contains a shuffle: every lane takes part)
            const unsigned row = __ballot_sync(FULL, valid & mb);
            if (lane == i) mask = row;
        }
    } else {                                        // some lane's shift is ambiguous here: image per particle
        for (int j = 0; j < cnt; ++j) {
            T d2 = dist2_nearest<T>(qx, qy, qz, tile[j], tile[32 + j], tile[64 + j], tv.period);
            mask |= (unsigned)(sizeof(T) == 8 ? d2 <= top0 : d2 < top0) << j;
        }
    }
    mask &= ~done;
    if (!pass) mask = 0;
#ifdef KNN_STATS
    {
        unsigned tot = __popc(mask);
        for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
        STAT(0, 1);            // buckets visited
        STAT(1, tot);          // survivors (lane sum)
    }
#endif
    // phase B: all lanes insert their survivors together
    while (__any_sync(FULL, mask != 0)) {
        STAT(2, 1);            // rounds
        bool ins = false;
        if (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const T d2 = wide ? dist2_nearest<T>(qx, qy, qz, tile[j], tile[32 + j], tile[64 + j], tv.period)
                              : dist2<T>(sx, sy, sz, tile[j], tile[32 + j], tile[64 + j]);
            if (below_top(q, d2)) { Replace<T, W_t, LOGC, WITH_IDX>::run(q, d2, (uint32_t)(start + j)); ins = true; }
        }
#ifdef KNN_STATS
        unsigned b = __ballot_sync(FULL, ins);
        STAT(3, __popc(b));    // inserts (lane sum)
        STAT(4, b != 0);       // rounds with at least one insert
#endif
        (void)ins;
    }
}

template <typename T, typename W_t, bool WITH_IDX>
__host__ __device__ inline size_t knn_smem_per_warp(int K2)
{
    size_t b = (size_t)K2 * 32 * sizeof(typename KeyOf<T>::type) + (size_t)K2 * 32 * sizeof(W_t) + 96 * sizeof(T);
    return (b + 127) & ~(size_t)127;
}

// padding leaves [k, K2) hold the smallest key; until k candidates have been seen the k-th distance is "+inf"
__device__ __forceinline__ float padding_key(float) { return -1.0f; }
__device__ __forceinline__ uint32_t padding_key(double) { return 0u; }
__device__ __forceinline__ float init_top(float) { return 3.0e38f; }
__device__ __forceinline__ double init_top(double) { return __hiloint2double(0x7fe00000, 0); }

template <typename T, typename W_t, int LOGC, bool WITH_IDX, bool FUSE_RHO, bool PERIODIC>
__global__ void __launch_bounds__(LOGC >= 0 ? KNN_MAX_WARPS * 32 : KNN_WARPS * 32, 1) k_knn(KnnArgs<T> a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef typename KeyOf<T>::type Key;
    const TreeView<T> &tv = a.tv;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warps = blockDim.x >> 5;
    const int k = a.k, LOG = LOGC >= 0 ? LOGC : a.LOG, K2 = 1 << LOG;

    Queue<T, W_t> q;
    q.LOG = LOG; q.K2 = K2;
    q.hi = nullptr;
    q.hb = q.wb = 0;
    T *tile;
    if (LOGC >= 0) {
        // one block per SM: [pad][leaves of every warp][winner bytes of every warp][candidate tiles]; the leaves start
        // at a multiple of one warp's leaf region (K2 * 128 bytes) in the shared window, so every warp's leaf and winner
        // regions are aligned to their own size (queue_replace_max_fast)
        static_assert(LOGC < 0 || sizeof(Key) == 4, "leaf stride");
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
        const uint32_t A = (uint32_t)K2 * 128u;
        const uint32_t pad = (A - (sbase & (A - 1u))) & (A - 1u);
        unsigned char *leaves = smem_raw + pad + (size_t)wib * A;
        unsigned char *wins = smem_raw + pad + (size_t)warps * A + (size_t)wib * K2 * 32;
        q.hd = reinterpret_cast<Key *>(leaves) + lane;
        q.W = reinterpret_cast<W_t *>(wins) + lane * 4;      // swizzled winner bytes: wofs()
        q.hb = sbase + pad + (uint32_t)wib * A + (uint32_t)lane * 4u;
        q.wb = sbase + pad + (uint32_t)warps * A + (uint32_t)wib * (uint32_t)K2 * 32u + (uint32_t)lane * 4u;
        // candidate tiles: as many as fit into the alignment gap in front of the leaves, the rest behind the winner bytes
        // (this is what lets the 21st warp in: 8 KB of alignment would otherwise be lost)
        const uint32_t tile_bytes = 96u * (uint32_t)sizeof(T);
        const uint32_t nfront = pad / tile_bytes;
        const uint32_t behind = pad + (uint32_t)warps * (A + (uint32_t)K2 * 32u);
        if ((uint32_t)wib < nfront) tile = reinterpret_cast<T *>(smem_raw + (size_t)wib * tile_bytes);
        else tile = reinterpret_cast<T *>(smem_raw + behind + (size_t)((uint32_t)wib - nfront) * tile_bytes);
        {
            uint32_t dyn;
            asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
            const uint32_t nback = (uint32_t)warps > nfront ? (uint32_t)warps - nfront : 0u;
            if (behind + nback * tile_bytes > dyn) __trap();       // the host sized the block for another window base
        }
    } else {
        unsigned char *base = smem_raw + knn_smem_per_warp<T, W_t, WITH_IDX>(K2) * wib;
        q.hd = reinterpret_cast<Key *>(base) + lane;
        unsigned char *after = base + (size_t)K2 * 32 * sizeof(Key);
        q.W = reinterpret_cast<W_t *>(after) + lane;
        after += (size_t)K2 * 32 * sizeof(W_t);
        tile = reinterpret_cast<T *>(after);      // 16-byte aligned: every array above is a multiple of 32 bytes
    }
    q.lo = a.lo_scratch ? a.lo_scratch + ((size_t)blockIdx.x * warps + wib) * (size_t)K2 * 32 + lane : nullptr;
    const T L = tv.period;

    for (;;) {
        unsigned long long b64 = 0;
        if (lane == 0) b64 = atomicAdd(a.next_bucket, 1ull);
        const int64_t bucket = (int64_t)__shfl_sync(FULL, b64, 0);
        if (bucket >= tv.nsplit) return;

        const int start = (int)min(tv.n, (int64_t)32 * bucket);
        const int cnt = (int)min(tv.n, (int64_t)32 * (bucket + 1)) - start;
        const bool active = lane < cnt && (!tv.query_mask || tv.query_mask[start + lane]);
        if (!__any_sync(FULL, active)) continue;    // no query of this bucket is selected
        T qx = 0, qy = 0, qz = 0;
        if (active) { qx = tv.x[start + lane]; qy = tv.y[start + lane]; qz = tv.z[start + lane]; }
        if (WITH_IDX) q.hi = a.nn_ids + (size_t)bucket * K2 * 32 + lane;

        for (int e = k; e < K2; ++e) {              // padding leaves hold the smallest key and never win
            q.hd[e * 32] = padding_key(T());
            if (sizeof(T) == 8) q.lo[e * 32] = 0u;
            if (WITH_IDX) q.hi[e * 32] = ~0u;
        }
        int filled = 0;                             // leaves filled so far; -1 once the tournament has been built
        q.etop = 0;
        q.top = init_top(T());                      // "+inf" until k are held (pq.h:185-192)
        q.top_hi = 0x7fe00000u; q.top_lo = 0u; q.lo_known = true;

        // own bucket first (smooth.h:200-208): the query lies inside its bounds, so no shift applies
        {
            const bool amb = active && shift_is_ambiguous<T, PERIODIC>(load_box(tv.box, tv.nsplit + bucket), qx, qy, qz, L);
            knn_visit_bucket<T, W_t, LOGC, WITH_IDX>(tv, bucket, active, qx, qy, qz, q, tile, lane,
                                                     PERIODIC && __any_sync(FULL, amb), qx, qy, qz, filled, k, active);
        }

        // then the sibling sub-tree at every level on the way up (smooth.h:209-241).  Inside a sub-tree the walk is
        // depth-first with the NEARER child first (the reference goes lower child first): both children of a node are
        // tested together (their bounds are adjacent in memory), the warp votes on which one lies nearer to most of
        // its queries, and one bit per level remembers that the other child is still to come.  Near buckets tighten
        // every lane's k-th distance early, so later buckets yield fewer candidates that have to be inserted only to
        // be pushed out again (CPU model of the walk, scripts/knn_walk_model.py: -10 % buckets visited, -18 % insert
        // rounds).  The result does not depend on the order (exact ties at the k-th distance aside).
        int64_t cell = tv.nsplit + bucket;
        // PERIODIC: the wrapped images are tested until the queues are full; from then on the k-th distance only
        // shrinks, and if no active query's ball reaches a face of the period the rest of the walk runs the plain
        // (cheaper) cell test -- an interior warp never sees a wrapped image
        bool wrap = PERIODIC, undecided = PERIODIC;
        // the walk tests one sibling per level on the way up: lane l asks for the bounds of the one l levels up now
        // (a 48-byte record can straddle two lines)
        {
            const int64_t up = cell >> lane;
            if (lane < 31 && up > 1) {
                const T *b = tv.box + (up ^ 1) * 6;
                prefetch_l1(b);
                prefetch_l1(b + 5);
            }
        }
        while (cell != 1) {
            int64_t cp = cell ^ 1;
            int depth = 0;
            uint32_t pend = 0, pend_right = 0;      // bit d: the node at depth d has a child left to visit / it is the right one
            T sx, sy, sz;
            bool pass;
            {
                const Box6<T> b = load_box(tv.box, cp);
                STAT(5, 1);    // nodes tested
                const T gap2 = box_gap2_w<T, PERIODIC>(wrap, b, qx, qy, qz, L, sx, sy, sz);
                // (top * slack overflows to +inf while the queue still holds its initial huge keys; empty
                // nodes have inverted bounds and an infinite gap)
                pass = active && gap2 <= q.top * Ex<T>::slack() && gap2 < Ex<T>::inf();
            }
            for (;;) {
                bool descend = false;
                if (__any_sync(FULL, pass)) {
                    if (cp >= tv.nsplit) {
                        bool amb = false;
                        if (PERIODIC && wrap) amb = pass && shift_is_ambiguous<T, PERIODIC>(load_box(tv.box, cp), sx, sy, sz, L);
                        knn_visit_bucket<T, W_t, LOGC, WITH_IDX>(tv, cp - tv.nsplit, pass, sx, sy, sz, q, tile, lane,
                                                                 PERIODIC && __any_sync(FULL, amb), qx, qy, qz, filled, k, active);
                    } else {
                        const Box6<T> bl = load_box(tv.box, 2 * cp), br = load_box(tv.box, 2 * cp + 1);
                        // one level ahead, by otherwise idle lanes: the four grandchildren's bounds (contiguous), or,
                        // under a parent of leaves, the coordinates of both candidate buckets (contiguous per axis)
                        if (4 * cp < 2 * tv.nsplit) {
                            if (2 * cp < tv.nsplit) {
                                if (lane < 3) prefetch_l1(reinterpret_cast<const char *>(tv.box + 4 * cp * 6) + min(lane * 96, 4 * 6 * (int)sizeof(T) - 1));
                            } else if (lane < 12) {
                                const int ax = lane >> 2;
                                const T *c = ax == 0 ? tv.x : (ax == 1 ? tv.y : tv.z);
                                const int64_t first = 32 * (2 * cp - tv.nsplit) + (lane & 3) * (128 / (int)sizeof(T));
                                if (first < tv.n && ((lane & 3) * 128 < 64 * (int)sizeof(T))) prefetch_l1(c + first);
                            }
                        }
                        STAT(5, 2);
                        T lx_, ly_, lz_, rx_, ry_, rz_;
                        const T gl = box_gap2_w<T, PERIODIC>(wrap, bl, qx, qy, qz, L, lx_, ly_, lz_);
                        const T gr = box_gap2_w<T, PERIODIC>(wrap, br, qx, qy, qz, L, rx_, ry_, rz_);
                        const T bound = q.top * Ex<T>::slack();
                        const bool pl = active && gl <= bound && gl < Ex<T>::inf();
                        const bool pr = active && gr <= bound && gr < Ex<T>::inf();
                        const unsigned vl = __ballot_sync(FULL, pl), vr = __ballot_sync(FULL, pr);
                        if (vl | vr) {
                            const unsigned nearer_r = __ballot_sync(FULL, active && gr < gl);
                            const unsigned nearer_l = __ballot_sync(FULL, active && gl < gr);
                            const bool right_first = vr && (!vl || __popc(nearer_r) > __popc(nearer_l));
                            if (vl && vr) {
                                pend |= 1u << depth;
                                if (right_first) pend_right &= ~(1u << depth); else pend_right |= 1u << depth;
                            }
                            cp = 2 * cp + (right_first ? 1 : 0);
                            ++depth;
                            pass = right_first ? pr : pl;
                            sx = right_first ? rx_ : lx_; sy = right_first ? ry_ : ly_; sz = right_first ? rz_ : lz_;
                            descend = true;
                        }
                    }
                }
                if (descend) continue;
                // back up to the nearest level that still owes a child; re-test it against the tighter bound
                bool found = false;
                while (depth > 0) {
                    --depth;
                    cp >>= 1;
                    if ((pend >> depth) & 1u) {
                        pend &= ~(1u << depth);
                        cp = 2 * cp + ((pend_right >> depth) & 1u);
                        ++depth;
                        found = true;
                        break;
                    }
                }
                if (!found) break;
                const Box6<T> b = load_box(tv.box, cp);
                STAT(5, 1);
                const T gap2 = box_gap2_w<T, PERIODIC>(wrap, b, qx, qy, qz, L, sx, sy, sz);
                pass = active && gap2 <= q.top * Ex<T>::slack() && gap2 < Ex<T>::inf();
            }
            cell >>= 1;
            if (PERIODIC && undecided && filled < 0) {
                undecided = false;
                const T r = Ex<T>::sqrt(q.top) * (T)1.0001 + a.face_eps;        // q.top: (an upper bound of) the k-th d2
                const bool reaches = active && (qx - r <= a.face_lo[0] || qx + r >= a.face_hi[0] || qy - r <= a.face_lo[1]
                                                || qy + r >= a.face_hi[1] || qz - r <= a.face_lo[2] || qz + r >= a.face_hi[2]);
                wrap = __any_sync(FULL, reaches);
            }
        }

        if (active) {
            const int64_t p = start + lane;
            const T h = (T)0.5 * Ex<T>::sqrt(top_exact(q));    // smooth.h:429
            a.smooth_t[p] = h;
            if (FUSE_RHO) {
                // smDensity (smooth.h:626-647) over the queue: the gather set of radius 2h differs from the
                // k nearest only by entries at q = 2 where both kernels vanish.  All masses are equal here.
                const T ih = (T)(1.0 / (double)h);
                const T ih2 = ih * ih;
                const T norm = (T)(0.31830988618379067154 * (double)ih * (double)ih2);
                const T m = (T)a.uniform_mass;
                T rho = 0;
                for (int e0 = 0; e0 < k; e0 += 8) {                 // eight queue entries in flight (float64: their low
                    T d2v[8];                                       // words come from the L2-resident scratch)
#pragma unroll
                    for (int u = 0; u < 8; ++u) d2v[u] = e0 + u < k ? queue_value(q, e0 + u) : (T)0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (e0 + u >= k) break;
                        T q2 = d2v[u] * ih2;
                        T w = (T)kern_w(a.kernel_id, a.wendland_self, q2);
                        w *= norm;
                        rho += w * m;
                    }
                }
                a.rho_t[p] = rho;
            }
        }
        if (WITH_IDX && a.nn_d2) {                  // the queued distances, beside the ids (nn() only)
            T *d2o = a.nn_d2 + (size_t)bucket * K2 * 32 + lane;
            for (int e = 0; e < k; ++e) d2o[e * 32] = queue_value(q, e);
        }
        __syncwarp();
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

template <typename T, typename W_t, int LOGC, bool WITH_IDX, bool FUSE_RHO, bool PERIODIC>
static void launch_knn_p(Tree &t, KnnArgs<T> &a)
{
    auto kern = k_knn<T, W_t, LOGC, WITH_IDX, FUSE_RHO, PERIODIC>;
    const int K2 = 1 << a.LOG;
    const size_t per_warp = knn_smem_per_warp<T, W_t, WITH_IDX>(K2);
    int warps;
    size_t smem;
    if (LOGC >= 0) {
        // one block per SM; up to K2 * 128 bytes of alignment padding in front (see k_knn)
        // The dynamic shared memory starts behind the per-block reservation (1 KB), so the leaves start `gap` bytes in;
        // that gap holds candidate tiles.  (Sized for that base; the kernel traps if its window base disagrees.)
        const size_t A = (size_t)K2 * 128, tile_bytes = 96 * sizeof(T), per_queue = A + (size_t)K2 * 32;
        int dev0 = 0, reserved = 1024;
        PNB_CUDA(cudaGetDevice(&dev0));
        if (cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, dev0) != cudaSuccess) { cudaGetLastError(); reserved = 1024; }
        const size_t gap = (A - (size_t)reserved % A) % A, nfront = gap / tile_bytes;
        auto need = [&](int w) { return gap + per_queue * w + ((size_t)w > nfront ? (size_t)w - nfront : 0) * tile_bytes; };
        warps = KNN_MAX_WARPS;
        while (warps > 1 && need(warps) > (size_t)227 * 1024) --warps;
        if (g_opt_knn_blocks_per_sm > 0) warps = std::min(warps, 2 * g_opt_knn_blocks_per_sm);
        smem = need(warps);
    } else {
        // two warps per block unless one warp's queues already need more than half the shared memory
        warps = per_warp * KNN_WARPS <= 227 * 1024 ? KNN_WARPS : 1;
        smem = per_warp * warps;
    }
    if (smem > 227 * 1024 || warps < 1)
        throw Error(PNB_ERR_VALUE, "number of smoothing neighbours too large for the on-chip neighbour queues "
                                   "(at most 1024)");
    PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, per_sm = 0;
    PNB_CUDA(cudaGetDevice(&dev));
    PNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PNB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    // a search running beside other work (multi-GPU: interior queries on one thread, halo pipeline on another) leaves
    // part of each SM's shared memory to the kernels of the other thread
    if (LOGC < 0 && g_opt_knn_blocks_per_sm > 0) per_sm = std::min(per_sm, g_opt_knn_blocks_per_sm);
    const int64_t want = (t.nsplit + warps - 1) / warps;
    const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sms * per_sm);    // persistent: one resident wave
    DevBuf counter(sizeof(unsigned long long)), lo;
    PNB_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(unsigned long long), t.stream));
    a.next_bucket = counter.as<unsigned long long>();
    a.lo_scratch = nullptr;
    if (sizeof(T) == 8) {
        lo.alloc(sizeof(uint32_t) * (size_t)grid * warps * K2 * 32);
        a.lo_scratch = lo.as<uint32_t>();
    }
#ifdef KNN_STATS
    unsigned long long zero[8] = {0};
    PNB_CUDA(cudaMemcpyToSymbol(g_stats, zero, sizeof(zero)));
#endif
    KernelTimer timer(0, t.stream);
    PNB_LAUNCH(kern, grid, warps * 32, smem, t.stream, a);
    timer.stop();
#ifdef KNN_STATS
    unsigned long long st[8];
    PNB_CUDA(cudaMemcpyFromSymbol(st, g_stats, sizeof(st)));
    const double nb = (double)t.nsplit;
    fprintf(stderr, "[knn stats] per home bucket: visited %.1f  survivors/lane %.1f  rounds %.1f  inserts/lane %.1f  "
            "rounds_with_insert %.1f  nodes %.1f\n", st[0] / nb, st[1] / nb / 32, st[2] / nb, st[3] / nb / 32, st[4] / nb, st[5] / nb);
#endif
}

template <typename T, typename W_t, int LOGC, bool WITH_IDX, bool FUSE_RHO>
static void launch_knn(Tree &t, KnnArgs<T> &a)
{
    if (a.tv.periodic) launch_knn_p<T, W_t, LOGC, WITH_IDX, FUSE_RHO, true>(t, a);
    else launch_knn_p<T, W_t, LOGC, WITH_IDX, FUSE_RHO, false>(t, a);
}

template <typename T, typename W_t, int LOGC>
static void dispatch_knn(Tree &t, KnnArgs<T> &a, bool keep_ids, bool fuse_rho)
{
    if (keep_ids) {
        if (fuse_rho) launch_knn<T, W_t, LOGC, true, true>(t, a);
        else launch_knn<T, W_t, LOGC, true, false>(t, a);
    } else if (fuse_rho) launch_knn<T, W_t, LOGC, false, true>(t, a);
    else launch_knn<T, W_t, LOGC, false, false>(t, a);
}

template <typename T>
static void run_knn_typed(Tree &t, int k, double period, bool fuse_rho, const KernelParams &kp, bool keep_ids, bool keep_d2)
{
    if (!t.smooth_t.p) t.smooth_t.alloc(sizeof(T) * t.n);
    if (fuse_rho && !t.rho_t.p) t.rho_t.alloc(sizeof(T) * t.n);
    KnnArgs<T> a;
    a.tv = make_view<T>(t, period);
    a.k = k;
    a.LOG = 0;
    while ((1 << a.LOG) < k) ++a.LOG;
    if (a.LOG > MAXLOG) throw Error(PNB_ERR_VALUE, "number of smoothing neighbours too large (at most 1024)");
    const size_t slots = (size_t)t.nsplit * ((size_t)1 << a.LOG) * 32;
    t.nn_valid = false;
    if (keep_ids) {
        try {
            if (t.nn_ids.bytes != slots * sizeof(uint32_t)) t.nn_ids.alloc(slots * sizeof(uint32_t));
            if (keep_d2) { if (t.nn_d2.bytes != slots * sizeof(T)) t.nn_d2.alloc(slots * sizeof(T)); }
            else t.nn_d2.release();
        } catch (const Error &e) {
            if (e.code != PNB_ERR_MEMORY || keep_d2) throw;      // nn() needs its lists; populate() can do without
            t.nn_ids.release();
            t.nn_d2.release();
            keep_ids = false;
        }
    }
    if (!keep_ids) {
        t.nn_ids.release();
        t.nn_d2.release();
    }
    for (int d = 0; d < 3; ++d) {
        const double lo = t.root_max[d] - period, hi = t.root_min[d] + period;
        T flo = (T)lo, fhi = (T)hi;
        if ((double)flo < lo) flo = std::nextafter(flo, std::numeric_limits<T>::infinity());     // rounded towards the inside
        if ((double)fhi > hi) fhi = std::nextafter(fhi, -std::numeric_limits<T>::infinity());
        a.face_lo[d] = flo;
        a.face_hi[d] = fhi;
    }
    {
        double scale = std::fabs(period);
        for (int d = 0; d < 3; ++d) scale = std::max(scale, std::max(std::fabs(t.root_min[d]), std::fabs(t.root_max[d])) + std::fabs(period));
        a.face_eps = (T)(16.0 * (double)std::numeric_limits<T>::epsilon() * scale);
    }
    a.smooth_t = t.smooth_t.as<T>();
    a.rho_t = t.rho_t.as<T>();
    a.uniform_mass = t.uniform_mass;
    a.kernel_id = kp.kernel_id;
    a.wendland_self = kp.wendland_self;
    a.nn_ids = keep_ids ? t.nn_ids.as<uint32_t>() : nullptr;
    a.nn_d2 = keep_ids && keep_d2 ? t.nn_d2.as<T>() : nullptr;
    switch (a.LOG) {                                   // the usual neighbour counts get fully unrolled queues
    case 5: dispatch_knn<T, uint8_t, 5>(t, a, keep_ids, fuse_rho); break;
    case 6: dispatch_knn<T, uint8_t, 6>(t, a, keep_ids, fuse_rho); break;
    case 7: dispatch_knn<T, uint8_t, 7>(t, a, keep_ids, fuse_rho); break;
    default:
        if (a.LOG <= 8) dispatch_knn<T, uint8_t, -1>(t, a, keep_ids, fuse_rho);
        else dispatch_knn<T, uint16_t, -1>(t, a, keep_ids, fuse_rho);
    }
    if (keep_ids) {
        t.nn_valid = true;
        t.nn_k = k;
        t.nn_K2 = 1 << a.LOG;
    }
}

// Neighbour lists cost n * K2 * 4 bytes (k = 64: 256 B per particle).  They are kept when they take at most a
// quarter of the device's memory (16.7 M particles: 4.3 GB; 134 M: 34 GB); pnb_set_option "neighbour_lists" /
// PNB_NEIGHBOUR_LISTS: 0 disables them, 1 forces them.  Without lists -- or if the allocation fails -- the
// reductions fall back to the tree-walking gather (gather.cu).  (No per-call cudaMemGetInfo: a driver query
// per step serialises against anything else talking to the driver, e.g. a monitoring nvidia-smi.)
bool neighbour_lists_fit(const Tree &t, int k, bool with_d2)
{
    if (g_opt_neighbour_lists == 0) return false;
    if (g_opt_neighbour_lists == 1) return true;
    static size_t total = 0;
    if (!total) {
        size_t free_b = 0;
        if (cudaMemGetInfo(&free_b, &total) != cudaSuccess) { cudaGetLastError(); total = 0; return false; }
    }
    int LOG = 0;
    while ((1 << LOG) < k) ++LOG;
    const size_t per = 4 + (with_d2 ? tsize(t.dtype) : 0);
    const size_t need = (size_t)t.nsplit * ((size_t)1 << LOG) * 32 * per;
    return need <= total / 4;
}

void run_knn(Tree &t, int k, double period, bool fuse_rho, const KernelParams &kp, bool keep_ids, bool keep_d2)
{
    if (k > t.n)
        throw Error(PNB_ERR_VALUE, "Number of smoothing particles exceeds number of particles in tree");
    if (k < 1) throw Error(PNB_ERR_VALUE, "Number of smoothing particles must be positive");
    // the density can ride along only when every particle has the same mass
    if (fuse_rho && !(t.has_mass && t.mass_is_uniform)) fuse_rho = false;
    if (t.dtype == PNB_F64) run_knn_typed<double>(t, k, period, fuse_rho, kp, keep_ids, keep_d2);
    else run_knn_typed<float>(t, k, period, fuse_rho, kp, keep_ids, keep_d2);
    t.smooth_valid = true;                  // (for the particles selected by the query mask, if one is set)
    t.smooth_k = k;
    t.smooth_period = period > 0 ? period : 0;
    t.rho_valid = fuse_rho;
    t.rho_kernel = kp.kernel_id;
}

}  // namespace pnb
