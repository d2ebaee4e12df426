// tree_morton.cu -- K1: the device tree, built Morton-first.
//
// Replaces kdBuildTree / kdBuildNode / kdSelect / kdUpPass (pynbody/kdtree/kd.cpp:31-238) for the tree the
// kernels traverse.  (The reference-SHAPED tree that KDTree.serialize() exposes is a different object with the
// reference's split rule; it is built on demand by tree_build.cu / export.cu.)
//
// The tree is the same implicit complete binary tree as before -- heap order, root = 1, every bucket at depth
// `levels`, bucket j = tree ranks [32j, 32j+32) -- but it is built in two regimes:
//
//   coarse  one 63-bit Morton key per particle (21 bits per axis over the cube around the particle set), ONE
//           radix sort of (key, id) pairs.  Consecutive runs of SEG = 4096 (float32) / 2048 (float64) particles
//           of the sorted order are the sub-trees at depth `levels - LSEG`; the levels above them split the
//           Morton order by rank.  This replaces 12 of the 19 breadth-first levels of the old build (each a
//           pass over all particles plus a global prefix sum) by a single sort.
//   fine    one CTA per segment loads its particles into shared memory and builds the lower LSEG = 7 / 6
//           levels as a genuine kd-tree there: per level the tight bounds of every node, split along the
//           longest axis, exact median by rank via a bitonic sort of (coordinate, slot) pairs inside each node.
//           Bucket quality -- which is what the neighbour search pays for -- is that of a median-split
//           kd-tree; only the segment outlines follow the Morton curve.  The CTA then writes its particles in
//           tree order as structure-of-arrays (x[], y[], z[], mass[], order[]): a bucket is 32 consecutive
//           elements = one 128/256-byte line per coordinate.
//
// Node bounds are tight everywhere: computed from the particles inside the segments, as unions of children
// above them.  Trailing empty nodes carry inverted (+inf, -inf) bounds so that no search opens them.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace pnb {

namespace {

template <typename T> struct Ord;
template <> struct Ord<float> {
    typedef uint32_t U;
    static __device__ __forceinline__ U enc(float v) { U u = __float_as_uint(v); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
    static __device__ __forceinline__ float dec(U u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};
template <> struct Ord<double> {
    typedef unsigned long long U;
    static __device__ __forceinline__ U enc(double v)
    {
        U u = (U)__double_as_longlong(v);
        return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
    }
    static __device__ __forceinline__ double dec(U u)
    {
        return __longlong_as_double((long long)((u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u));
    }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
};

// ---- global bounds: [0..2] = min xyz, [3..5] = max xyz, order-preserving unsigned encoding -------------------
template <typename T>
__global__ void k_bounds_init(typename Ord<T>::U *b)
{
    if (threadIdx.x < 3) b[threadIdx.x] = ~(typename Ord<T>::U)0;
    else if (threadIdx.x < 6) b[threadIdx.x] = 0;
}
template <typename T>
__global__ void k_bounds(const T *__restrict__ c, int64_t n, typename Ord<T>::U *__restrict__ b)
{
    typedef typename Ord<T>::U U;
    T mn[3], mx[3];
    for (int d = 0; d < 3; ++d) { mn[d] = Ord<T>::inf(); mx[d] = -Ord<T>::inf(); }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        for (int d = 0; d < 3; ++d) {
            const T v = c[d * n + i];
            mn[d] = fmin(mn[d], v); mx[d] = fmax(mx[d], v);
        }
    for (int d = 0; d < 3; ++d)
        for (int o = 16; o; o >>= 1) {
            mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    if ((threadIdx.x & 31) == 0)
        for (int d = 0; d < 3; ++d) {
            atomicMin(&b[d], (U)Ord<T>::enc(mn[d]));
            atomicMax(&b[3 + d], (U)Ord<T>::enc(mx[d]));
        }
}

__device__ __forceinline__ uint64_t spread21(uint32_t v)       // bit i -> bit 3i
{
    uint64_t x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

template <typename T>
__global__ void k_morton_keys(const T *__restrict__ c, int64_t n, const typename Ord<T>::U *__restrict__ b,
                              uint64_t *__restrict__ keys, uint32_t *__restrict__ ids)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double mn[3], ext = 0;
    for (int d = 0; d < 3; ++d) {
        mn[d] = (double)Ord<T>::dec(b[d]);
        ext = fmax(ext, (double)Ord<T>::dec(b[3 + d]) - mn[d]);
    }
    const double scale = ext > 0 ? 2097152.0 / ext : 0.0;       // 2^21 cells across the longest extent
    uint64_t key = 0;
    for (int d = 0; d < 3; ++d) {
        double q = ((double)c[d * n + i] - mn[d]) * scale;
        uint32_t u = q >= 2097151.0 ? 2097151u : (q > 0 ? (uint32_t)q : 0u);
        key |= spread21(u) << d;
    }
    keys[i] = key;
    ids[i] = (uint32_t)i;
}

// ---- fine levels: one CTA per segment -------------------------------------------------------------------------
template <typename T> struct SegCfg;
template <> struct SegCfg<float> { static constexpr int LSEG = 7; };     // 4096 particles, 1024 threads
template <> struct SegCfg<double> { static constexpr int LSEG = 6; };    // 2048 particles,  512 threads

template <typename T>
struct SegArgs {
    int64_t n, nsplit;
    int levels, lseg;                 // lseg = min(levels, LSEG): levels built inside a segment
    const T *cols;                    // [3][n] file order
    const T *mass;                    // [n] file order or null
    const uint32_t *ids;              // [n] Morton order -> file index
    T *x, *y, *z, *m;                 // tree order
    uint32_t *order;                  // tree position -> file index
    T *box;                           // [2*nsplit][6]
};

// ---- the sorting network of one level, 4 slots per thread -------------------------------------------------------
// Slot e = 4 * tid + r lives in register r of thread tid: compare-exchange partners at strides 1 and 2 are in the same
// thread, at strides 4 .. 64 in another lane of the same warp (one shuffle per register), at strides >= 128 in another
// warp (those stages go through shared memory).  `up` for a pair: ascending iff bit k2 of the slot index inside its
// node is clear (node size m; for k2 == m every pair is ascending).  Equal keys never swap.
template <typename T>
__device__ __forceinline__ void cex(T &ka, int &ia, T &kb, int &ib, bool up)
{
    const bool sw = up ? (ka > kb) : (ka < kb);
    if (sw) { const T t = ka; ka = kb; kb = t; const int u = ia; ia = ib; ib = u; }
}
template <typename T>
__device__ __forceinline__ void reg_stages(T (&K)[4], int (&I)[4], int e0, int k2, int jstart, int mmask)
{
    for (int j = jstart; j >= 4; j >>= 1) {
        const int lm = j >> 2;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int e = e0 + r;
            const bool up = ((e & mmask) & k2) == 0, upper = (e & j) != 0;
            const T kb = __shfl_xor_sync(0xffffffffu, K[r], lm);
            const int ib = __shfl_xor_sync(0xffffffffu, I[r], lm);
            const bool take_min = up != upper;
            const bool sw = take_min ? (kb < K[r]) : (kb > K[r]);
            if (sw) { K[r] = kb; I[r] = ib; }
        }
    }
    if (jstart >= 2) {
        const bool up0 = ((e0 & mmask) & k2) == 0, up1 = (((e0 + 1) & mmask) & k2) == 0;
        cex(K[0], I[0], K[2], I[2], up0);
        cex(K[1], I[1], K[3], I[3], up1);
    }
    {
        const bool up0 = ((e0 & mmask) & k2) == 0, up2 = (((e0 + 2) & mmask) & k2) == 0;
        cex(K[0], I[0], K[1], I[1], up0);
        cex(K[2], I[2], K[3], I[3], up2);
    }
}

template <typename T>
__global__ void __launch_bounds__((32 << SegCfg<T>::LSEG) / 4) k_build_segment(SegArgs<T> a)
{
    typedef typename Ord<T>::U U;
    constexpr int SEG = 32 << SegCfg<T>::LSEG;
    constexpr int NT = SEG / 4;
    constexpr int MAXSUB = SEG / 64;                 // nodes of the deepest split level
    extern __shared__ __align__(16) unsigned char seg_smem[];
    T *px = reinterpret_cast<T *>(seg_smem);
    T *py = px + SEG, *pz = py + SEG, *sk = pz + SEG;
    U *sbox = reinterpret_cast<U *>(sk + SEG);       // [MAXSUB][6]
    uint16_t *perm = reinterpret_cast<uint16_t *>(sbox + MAXSUB * 6);

    const int tid = threadIdx.x, lane = tid & 31;
    const int segn = 32 << a.lseg;                   // slots of this segment (<= SEG)
    const int64_t base = (int64_t)blockIdx.x * segn;
    const int cnt = (int)max((int64_t)0, min((int64_t)segn, a.n - base));
    const int ltop = a.levels - a.lseg;              // depth of the segment roots
    const T inf = Ord<T>::inf();
    const int e0 = 4 * tid;                          // this thread's slots: e0 .. e0 + 3

    for (int i = tid; i < SEG; i += NT) {
        if (i < cnt) {
            const uint32_t id = a.ids[base + i];
            px[i] = a.cols[id]; py[i] = a.cols[a.n + id]; pz[i] = a.cols[2 * a.n + id];
        } else {
            px[i] = py[i] = pz[i] = inf;             // padding sorts to the end and is skipped by the bounds
        }
    }
    T K[4];
    int I[4];                                        // local particle held by each slot (>= cnt: padding)
#pragma unroll
    for (int r = 0; r < 4; ++r) I[r] = e0 + r;
    __syncthreads();

    for (int l = 0; l <= a.lseg; ++l) {
        const int m = segn >> l;                     // slots per node of this level
        const int nsub = 1 << l;
        const bool leaf = l == a.lseg;
        const int node_of_thread = e0 >> (5 + a.lseg - l);
        // 1. tight bounds of every node of the level: per thread, then across the lanes that share the node
        if (!leaf) {
            for (int i = tid; i < nsub * 6; i += NT) sbox[i] = (i % 6) < 3 ? ~(U)0 : (U)0;
            __syncthreads();
        }
        T mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int li = I[r];
            if (li < cnt) {
                const T x = px[li], y = py[li], z = pz[li];
                mn[0] = fmin(mn[0], x); mn[1] = fmin(mn[1], y); mn[2] = fmin(mn[2], z);
                mx[0] = fmax(mx[0], x); mx[1] = fmax(mx[1], y); mx[2] = fmax(mx[2], z);
            }
        }
        const int group = min(32, m >> 2);            // lanes per node inside a warp (a power of two >= 8)
        for (int o = group >> 1; o; o >>= 1)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
                mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
            }
        if (e0 < segn && (lane & (group - 1)) == 0) {
            if (leaf) {
                const int64_t node = a.nsplit + (int64_t)blockIdx.x * (segn / 32) + (e0 >> 5);
                T *b = a.box + node * 6;
                b[0] = mn[0]; b[1] = mn[1]; b[2] = mn[2]; b[3] = mx[0]; b[4] = mx[1]; b[5] = mx[2];
            } else {
                U *b = sbox + node_of_thread * 6;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    atomicMin(&b[d], (U)Ord<T>::enc(mn[d]));
                    atomicMax(&b[3 + d], (U)Ord<T>::enc(mx[d]));
                }
            }
        }
        if (leaf) break;
        __syncthreads();
        // 2. node bounds out (one thread per node), split axis = longest extent of the tight box
        int axis = 0;
        if (e0 < segn) {
            T bmn[3], bmx[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) { bmn[d] = Ord<T>::dec(sbox[node_of_thread * 6 + d]); bmx[d] = Ord<T>::dec(sbox[node_of_thread * 6 + 3 + d]); }
            if (bmx[1] - bmn[1] > bmx[axis] - bmn[axis]) axis = 1;
            if (bmx[2] - bmn[2] > bmx[axis] - bmn[axis]) axis = 2;
            if ((e0 & (m - 1)) == 0) {
                const int64_t node = ((int64_t)1 << (ltop + l)) + (int64_t)blockIdx.x * nsub + node_of_thread;
                T *b = a.box + node * 6;
                b[0] = bmn[0]; b[1] = bmn[1]; b[2] = bmn[2]; b[3] = bmx[0]; b[4] = bmx[1]; b[5] = bmx[2];
            }
        }
        // 3. sort every node's slots by the coordinate along its axis: the lower half is the lower child
        const T *pc = axis == 0 ? px : axis == 1 ? py : pz;
#pragma unroll
        for (int r = 0; r < 4; ++r) K[r] = (e0 < segn && I[r] < SEG) ? pc[I[r]] : inf;
        const int mmask = m - 1;
        for (int k2 = 2; k2 <= min(m, 128); k2 <<= 1) reg_stages<T>(K, I, e0, k2, k2 >> 1, mmask);
        for (int k2 = 256; k2 <= m; k2 <<= 1) {
            __syncthreads();                              // (sbox and the previous round's reads are done with)
#pragma unroll
            for (int r = 0; r < 4; ++r) { sk[e0 + r] = K[r]; perm[e0 + r] = (uint16_t)I[r]; }
            __syncthreads();
            for (int j = k2 >> 1; j >= 128; j >>= 1) {
                for (int t = tid; t < segn / 2; t += NT) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int p = i | j;
                    const bool up = ((i & mmask) & k2) == 0;
                    const T ka = sk[i], kb = sk[p];
                    if ((ka > kb) == up && ka != kb) {
                        sk[i] = kb; sk[p] = ka;
                        const uint16_t t16 = perm[i]; perm[i] = perm[p]; perm[p] = t16;
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) { K[r] = sk[e0 + r]; I[r] = perm[e0 + r]; }
            reg_stages<T>(K, I, e0, k2, 64, mmask);
        }
        __syncthreads();
    }
    // tree-ordered structure of arrays
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int e = e0 + r, li = I[r];
        if (e < cnt) {
            const uint32_t id = a.ids[base + li];
            a.x[base + e] = px[li]; a.y[base + e] = py[li]; a.z[base + e] = pz[li];
            a.order[base + e] = id;
            if (a.mass) a.m[base + e] = a.mass[id];
        }
    }
}

// levels above the segments: union of the children's bounds, bottom-up, one block
template <typename T>
__global__ void k_upper_bounds(T *__restrict__ box, int ltop)
{
    for (int level = ltop - 1; level >= 0; --level) {
        const int64_t nn = (int64_t)1 << level;
        for (int64_t j = threadIdx.x; j < nn; j += blockDim.x) {
            const int64_t node = nn + j;
            const T *l = box + 2 * node * 6, *r = l + 6;
            T *o = box + node * 6;
#pragma unroll
            for (int d = 0; d < 3; ++d) { o[d] = fmin(l[d], r[d]); o[3 + d] = fmax(l[3 + d], r[3 + d]); }
        }
        __syncthreads();
    }
}

__global__ void k_bucket_starts(int64_t n, int64_t nsplit, int32_t *__restrict__ bs)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j <= nsplit) bs[j] = (int32_t)min(n, 32 * j);
}

template <typename T>
size_t segment_smem()
{
    constexpr int SEG = 32 << SegCfg<T>::LSEG;
    return (size_t)SEG * 4 * sizeof(T) + (size_t)(SEG / 64) * 6 * sizeof(typename Ord<T>::U) + (size_t)SEG * 2 + 16;
}

// The lower `lseg` levels of the device tree, one CTA per sub-tree: `ids` lists the particles (file indices) so
// that the sub-tree rooted at node (levels - lseg, j) owns entries [j * 32 << lseg, ...).  Writes the bounds of
// the nodes at depth >= levels - lseg, the bucket offsets and the tree-ordered structure of arrays; t.box must
// have been allocated.
template <typename T>
void build_segments(Tree &t, const T *cols, const T *mass, const uint32_t *ids, int lseg)
{
    const int64_t n = t.n;
    cudaStream_t s = t.stream;
    const int BS = 256;
    t.bucket_start.alloc(sizeof(int32_t) * (t.nsplit + 1));
    t.order.alloc(sizeof(uint32_t) * n);
    t.x.alloc(sizeof(T) * n); t.y.alloc(sizeof(T) * n); t.z.alloc(sizeof(T) * n);
    if (mass) t.mass.alloc(sizeof(T) * n);
    t.has_mass = mass != nullptr;
    SegArgs<T> a;
    a.n = n; a.nsplit = t.nsplit; a.levels = t.levels;
    a.lseg = lseg;
    a.cols = cols; a.mass = mass; a.ids = ids;
    a.x = t.x.as<T>(); a.y = t.y.as<T>(); a.z = t.z.as<T>(); a.m = t.mass.as<T>();
    a.order = t.order.as<uint32_t>(); a.box = t.box.as<T>();
    const size_t smem = segment_smem<T>();
    PNB_CUDA(cudaFuncSetAttribute(k_build_segment<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (t.mass_ready) PNB_CUDA(cudaStreamWaitEvent(s, t.mass_ready, 0));   // masses uploaded behind the sorts
    const unsigned nseg = (unsigned)(t.nsplit >> lseg);
    PNB_LAUNCH((k_build_segment<T>), nseg, (32 << SegCfg<T>::LSEG) / 4, smem, s, a);
    PNB_LAUNCH(k_bucket_starts, (unsigned)((t.nsplit + 1 + BS - 1) / BS), BS, 0, s, n, t.nsplit, t.bucket_start.as<int32_t>());
}

template <typename T>
void build_morton_typed(Tree &t, DevBuf &c, const T *mass)
{
    typedef typename Ord<T>::U U;
    const int64_t n = t.n;
    cudaStream_t s = t.stream;
    const int BS = 256;
    const unsigned nb = (unsigned)((n + BS - 1) / BS);

    DevBuf bounds(sizeof(U) * 6), ids_sorted(sizeof(uint32_t) * n);
    PNB_LAUNCH((k_bounds_init<T>), 1, 32, 0, s, bounds.as<U>());
    PNB_LAUNCH((k_bounds<T>), std::min<unsigned>(nb, 148 * 8), BS, 0, s, c.as<T>(), n, bounds.as<U>());
    {
        DevBuf keys_in(sizeof(uint64_t) * n), keys_out(sizeof(uint64_t) * n), ids_in(sizeof(uint32_t) * n);
        PNB_LAUNCH((k_morton_keys<T>), nb, BS, 0, s, c.as<T>(), n, bounds.as<U>(), keys_in.as<uint64_t>(), ids_in.as<uint32_t>());
        size_t tmp_bytes = 0;
        PNB_CUDA((cub::DeviceRadixSort::SortPairs<uint64_t, uint32_t, int64_t>(
            nullptr, tmp_bytes, keys_in.as<uint64_t>(), keys_out.as<uint64_t>(), ids_in.as<uint32_t>(),
            ids_sorted.as<uint32_t>(), n, 0, 63, s)));
        DevBuf tmp(tmp_bytes);
        PNB_CUDA((cub::DeviceRadixSort::SortPairs<uint64_t, uint32_t, int64_t>(
            tmp.p, tmp_bytes, keys_in.as<uint64_t>(), keys_out.as<uint64_t>(), ids_in.as<uint32_t>(),
            ids_sorted.as<uint32_t>(), n, 0, 63, s)));
        g_launches += 9;                               // CUB's histogram + onesweep passes (approximate count)
        PNB_CUDA(cudaStreamSynchronize(s));            // keys/tmp go back to the block cache on scope exit
    }

    t.box.alloc(sizeof(T) * 6 * 2 * t.nsplit);
    PNB_CUDA(cudaMemsetAsync(t.box.p, 0, sizeof(T) * 6 * 2, s));      // node 0 is unused
    const int lseg = std::min(t.levels, SegCfg<T>::LSEG);
    build_segments<T>(t, c.as<T>(), mass, ids_sorted.as<uint32_t>(), lseg);
    const int ltop = t.levels - lseg;
    if (ltop > 0) PNB_LAUNCH((k_upper_bounds<T>), 1, 1024, 0, s, t.box.as<T>(), ltop);

    T rb[6];
    PNB_CUDA(cudaMemcpyAsync(rb, t.box.as<T>() + 6, sizeof(rb), cudaMemcpyDeviceToHost, s));
    PNB_CUDA(cudaStreamSynchronize(s));
    for (int d = 0; d < 3; ++d) { t.root_min[d] = (double)rb[d]; t.root_max[d] = (double)rb[3 + d]; }
}

}  // namespace

int segment_levels(int dtype) { return dtype == PNB_F64 ? SegCfg<double>::LSEG : SegCfg<float>::LSEG; }
void build_tree_segments(Tree &t, const void *cols, const void *mass_dev, const uint32_t *ids, int lseg)
{
    if (t.dtype == PNB_F64) build_segments<double>(t, (const double *)cols, (const double *)mass_dev, ids, lseg);
    else build_segments<float>(t, (const float *)cols, (const float *)mass_dev, ids, lseg);
}

// cols: file-order coordinates as dense columns T[3][n]; mass_dev: T[n] in file order or null
void build_tree_morton(Tree &t, DevBuf &cols, const void *mass_dev)
{
    if (t.n < 1) throw Error(PNB_ERR_VALUE, "kd-tree needs at least one particle");
    if (t.n >= ((int64_t)1 << 31)) throw Error(PNB_ERR_VALUE, "at most 2^31-1 particles per device tree");
    t.shape = packed_shape(t.n);
    t.levels = t.shape.levels;
    t.nsplit = (int64_t)1 << t.levels;
    if (t.dtype == PNB_F64) build_morton_typed<double>(t, cols, (const double *)mass_dev);
    else build_morton_typed<float>(t, cols, (const float *)mass_dev);
}

}  // namespace pnb
