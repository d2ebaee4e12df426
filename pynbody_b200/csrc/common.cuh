// common.cuh -- shared definitions of libpnb_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <stdexcept>
#include <vector>

#include "../../include/pnb_b200.h"

namespace pnb {

// ------------------------------------------------------------------------------------------------
// errors: C++ exceptions inside the library, translated to status codes at the C boundary
// ------------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string &m);

#define PNB_CUDA(call)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            throw ::pnb::Error(PNB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// every kernel launch of the library goes through this so bench.py can report gpu_launches
extern thread_local int64_t g_launches;
#define PNB_LAUNCH(kernel, grid, block, smem, stream, ...)                                        \
    do {                                                                                          \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                               \
        ++::pnb::g_launches;                                                                      \
        PNB_CUDA(cudaGetLastError());                                                             \
    } while (0)

// ------------------------------------------------------------------------------------------------
// streams
// ------------------------------------------------------------------------------------------------
// Every host thread that calls into the library works on its OWN non-blocking stream (one per device, created on
// first use): two threads -- e.g. the interior neighbour search of a domain and the halo pipeline of the multi-GPU
// layer -- run their kernels side by side, and nothing serialises behind the legacy default stream any more.
// Ordering against the caller's work (torch tensors are produced on the default stream):
//   * on entry, ApiScope makes the thread's stream wait for everything already enqueued on the default stream;
//   * on exit it drains the thread's stream, so outputs are complete when the call returns.
// pnb_set_option("streams", 0) / PNB_STREAMS=0 goes back to the default stream.
cudaStream_t lib_stream();
struct ApiScope {
    ApiScope();
    ~ApiScope();
};
// converts to the calling thread's stream wherever a cudaStream_t is expected (Tree::stream)
struct StreamRef {
    operator cudaStream_t() const { return lib_stream(); }
};

// ------------------------------------------------------------------------------------------------
// device buffer with RAII (stream-ordered allocator)
// ------------------------------------------------------------------------------------------------
// Device memory comes from a small caching allocator (api.cu): blocks are cudaMalloc'ed once, kept
// on per-size free lists when released and handed out again to the next request of the same size.
// A block goes back to the cache only when the stream that used it has been drained (every entry point ends
// with ApiScope's drain, and scratch buffers released mid-call are preceded by an explicit synchronisation), so
// the next user -- on whatever thread or stream -- finds it idle.  A steady-state step performs no driver
// allocation at all (cudaMalloc/cudaFree synchronise the device and were measured to stall for hundreds of ms).
void *cache_alloc(size_t bytes);
void cache_free(void *p, size_t bytes);
void cache_trim();
size_t cached_bytes();      // idle bytes the cache could hand out or give back
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;     // requested size
    size_t cap = 0;       // allocated size (cache size class)
    DevBuf() = default;
    explicit DevBuf(size_t b) { alloc(b); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes), cap(o.cap) { o.p = nullptr; o.bytes = o.cap = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; bytes = o.bytes; cap = o.cap; o.p = nullptr; o.bytes = o.cap = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    // Size classes: multiples of 512 B up to 1 MiB, then 8 geometric steps per octave (<= 12.5 % slack), so
    // that work buffers of slightly different particle counts (multi-GPU domains, successive snapshots) land
    // in the same class and are re-used instead of going back to cudaMalloc.
    static size_t size_class(size_t b) {
        if (b <= ((size_t)1 << 20)) return b == 0 ? 512 : (b + 511) & ~(size_t)511;
        int lg = 63 - __builtin_clzll((unsigned long long)b);
        const size_t step = (size_t)1 << (lg - 3);
        return (b + step - 1) & ~(step - 1);
    }
    void alloc(size_t b) {
        release();
        cap = size_class(b);
        p = cache_alloc(cap);
        bytes = b;
    }
    void release() { if (p) { cache_free(p, cap); p = nullptr; bytes = cap = 0; } }
    template <typename U> U *as() const { return reinterpret_cast<U *>(p); }
};

// a borrowed array registered through pnb_set_array (kdmain.cpp:511-579)
struct ArrRef {
    void *ptr = nullptr;
    int64_t stride0 = 0, stride1 = 0;
    int ncols = 0, dtype = -1, mem = 0;
    bool set() const { return ptr != nullptr; }
};

// ------------------------------------------------------------------------------------------------
// the tree handle
// ------------------------------------------------------------------------------------------------
// how the implicit tree's nodes map to ranges of the particle permutation (tree_build.cu)
struct TreeShape {
    int64_t n = 0;
    int levels = 0;     // buckets live at heap level `levels`
    int packed = 1;     // 1: fixed-capacity buckets (device tree); 0: the reference's median ranges
    int leafcap = 32;   // bucket capacity (packed) / leafsize (reference)
};
TreeShape packed_shape(int64_t n);
TreeShape reference_shape(int64_t n, int leafsize);
void shape_seg_of_node(const TreeShape &s, int level, int64_t j, int64_t &lo, int64_t &hi);

struct Tree {
    TreeShape shape;
    int dtype = PNB_F64;          // PNB_F32 / PNB_F64: dtype of pos/mass/smooth/rho ("T")
    int device = -1;              // CUDA device the tree's buffers live on; entry points reject calls from another device
    int64_t n = 0;
    int user_leafsize = 16;
    int levels = 0;               // depth of the device tree: buckets live at heap level `levels`
    int64_t nsplit = 1;           // number of buckets = 1 << levels
    StreamRef stream;             // the calling thread's stream (see ApiScope)
    cudaEvent_t mass_ready = nullptr;   // set while the masses are still on their way to the device (pnb_tree_create)

    // particles in tree order (T), structure of arrays
    DevBuf x, y, z, mass;
    bool has_mass = false;
    bool mass_is_uniform = false; // every particle has the same mass (then density can ride along the kNN pass)
    double uniform_mass = 0;
    DevBuf order;                 // uint32[n]: tree position -> file index
    DevBuf inv_order;             // uint32[n]: file index -> tree position (made on demand: pnb_knn_rows)
    DevBuf box;                   // T[2*nsplit][6]: tight bounds of every node (heap index, root = 1)
    DevBuf bucket_start;          // int32[nsplit + 1]
    DevBuf split_dim, split_val;  // int8 / T [2*nsplit]: split axis and value of internal nodes (export only)
    double root_min[3] = {0, 0, 0}, root_max[3] = {0, 0, 0};

    // results kept resident in tree order between populate calls
    DevBuf smooth_t, rho_t;       // T[n]
    bool smooth_valid = false;    // smooth_t holds this tree's own kNN smoothing lengths ...
    int smooth_k = 0;             // ... for this k
    double smooth_period = 0;     // ... and this effective period (<=0: open)
    bool rho_valid = false;       // rho_t was computed from smooth_t (fused with the kNN pass)
    int rho_kernel = -1;

    DevBuf query_mask;            // uint8[n] tree order, or empty: restricts populate / knn to a subset

    // neighbour lists of the last kNN pass (knn.cu): entry e of the particle at tree position 32*b+l is at
    // [(b*nn_K2 + e)*32 + l]; valid together with smooth_t (same k, same period, same query mask)
    DevBuf nn_ids;                // uint32[nsplit][nn_K2][32] tree positions of the neighbours (~0u: padding slot)
    DevBuf nn_d2;                 // T, same layout: their squared distances (kept only for nn())
    bool nn_valid = false;
    DevBuf pos4;                  // packed {x, y, z, m} records in tree order for the list reductions (gather_list.cu)
    int nn_k = 0, nn_K2 = 0;

    ArrRef arr[5];
};

inline size_t tsize(int dtype) { return dtype == PNB_F64 ? 8 : 4; }

// ---- stage timing (pnb_last_device_ms) ----------------------------------------------------------
struct StageTimer {
    int stage;
    cudaStream_t stream;
    cudaEvent_t a = nullptr, b = nullptr;
    int64_t launches0;
    StageTimer(int stage, cudaStream_t s);
    void stop();
    ~StageTimer();
};

// times one kernel launch with CUDA events on its stream (pnb_last_kernel_ms)
struct KernelTimer {
    int which;
    cudaStream_t stream;
    cudaEvent_t a = nullptr, b = nullptr;
    KernelTimer(int which, cudaStream_t s);
    void stop();        // records the end event, waits for it, stores the duration
    ~KernelTimer();
};

// ---- host <-> device movement of strided arrays (io.cu) ------------------------------------------
// Bring `ncols` columns of a strided [n][ncols] array (host or device) into a dense device buffer
// laid out column-major ([ncols][n]) in FILE order.
void fetch_columns(const void *ptr, int64_t stride0, int64_t stride1, int ncols, int dtype, int mem,
                   int64_t n, void *dev_out, cudaStream_t stream);
// Write a dense column-major device buffer ([ncols][n], file order) back to a strided destination.
void store_columns(void *ptr, int64_t stride0, int64_t stride1, int ncols, int dtype, int mem,
                   int64_t n, const void *dev_in, cudaStream_t stream);
bool host_is_pinned(const void *p);
// tree-order <-> file-order permutations on the device (element size 4 or 8, ncols columns)
void permute_to_tree(const Tree &t, const void *file_in, void *tree_out, int ncols, int elem, cudaStream_t s);
void permute_to_file(const Tree &t, const void *tree_in, void *file_out, int ncols, int elem, cudaStream_t s);

// ---- stages --------------------------------------------------------------------------------------
// cols: file-order coordinates as dense columns T[3][n]; mass_dev: T[n] in file order or null
void build_tree(Tree &t, DevBuf &cols, const void *mass_dev, const TreeShape &shape, bool want_soa);
// the device tree (32-particle buckets): one Morton sort + the lower levels as a kd-tree in shared memory (tree_morton.cu)
void build_tree_morton(Tree &t, DevBuf &cols, const void *mass_dev);
// the lowest levels of the packed device tree, built per sub-tree in shared memory (tree_morton.cu): `ids` orders the
// particles so that sub-tree j at depth levels - lseg owns entries [j * (32 << lseg), ...)
int segment_levels(int dtype);
void build_tree_segments(Tree &t, const void *cols, const void *mass_dev, const uint32_t *ids, int lseg);

struct KernelParams {
    int kernel_id;
    int k;
    double wendland_self;   // (21/16)*(1-0.0294*pow(k*0.01,-0.977)), sph/kernels.hpp:85-86
};
KernelParams make_kernel_params(int kernel_id, int k);

// kNN pass: writes t.smooth_t (and t.rho_t when fuse_rho); optional neighbour lists (file order rows)
// keep_ids: also record every particle's neighbour list in t.nn_ids (and the distances in t.nn_d2 if keep_d2)
void run_knn(Tree &t, int k, double period /* <=0: open */, bool fuse_rho, const KernelParams &kp,
             bool keep_ids, bool keep_d2);
bool neighbour_lists_fit(const Tree &t, int k, bool with_d2);
extern int g_opt_neighbour_lists, g_opt_gather_walk, g_opt_tree_build, g_opt_streams, g_opt_knn_blocks_per_sm;     // pnb_set_option (api.cu)
// reductions over the resident neighbour lists (gather_list.cu); same arguments as run_gather
void run_gather_lists(Tree &t, int propid, int k, double period, const KernelParams &kp,
                      const void *smooth_tree, const void *rho_tree, const void *qty_tree, int qdtype,
                      void *out_tree);
// nCnt of the reference's gather for every particle (tree order), through the tree walk (gather.cu)
void run_gather_counts(Tree &t, double period, const void *smooth_tree, int32_t *counts_tree);
// gather pass for property ids 2..8.  All inputs are dense device arrays in TREE order.
// smooth/rho: T[n]; qty: Tq [ncols][n]; out: T[n] (rho) or Tq[ncols_out][n].
// Returns true if some particle found more than k+500 neighbours (smooth.h:253-267).
bool run_gather(Tree &t, int propid, int k, double period, const KernelParams &kp,
                const void *smooth_tree, const void *rho_tree, const void *qty_tree, int qdtype,
                void *out_tree);
// reference-layout KDNode records (kd.h:29-40) for the user's leafsize (export.cu)
void export_reference_tree(Tree &t, void *kdnodes_out, int64_t n_nodes, int64_t *offsets_out);
void record_stage(int stage, double ms, int64_t launches);
void set_kernel_ms(int which, double ms);
// flags (uint8, TREE order, pre-zeroed) of the particles inside any of nq balls; centres are dense
// device columns T[3][nq], radii T[nq]
void run_mark_in_balls(Tree &t, const void *centres_cols, const void *radii, int64_t nq, double period, uint8_t *flags_tree);
int64_t run_ball_query(Tree &t, double cx, double cy, double cz, double r, double period,
                       int64_t *out_host, int64_t cap);

}  // namespace pnb
