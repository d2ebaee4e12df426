// api.cu -- the extern "C" boundary of libpnb_b200.so (include/pnb_b200.h).
//
// Host-side orchestration of what kdmain.cpp does around its kernels: argument checks with the
// reference's error classes (kdmain.cpp:136-182, 200-208, 367-372, 511-579), the periodicity check of
// smInit (smooth.h:71-124), array registration, and moving results back to the caller's borrowed
// arrays in file order.
#include "common.cuh"
#include <algorithm>
#include <map>
#include <mutex>

namespace pnb {

static thread_local std::string g_error;
thread_local int64_t g_launches = 0;
static thread_local double g_stage_ms[4] = {0, 0, 0, 0};
static thread_local int64_t g_stage_launches[4] = {0, 0, 0, 0};

void set_last_error(const std::string &m) { g_error = m; }
void record_stage(int stage, double ms, int64_t launches)
{
    g_stage_ms[stage] = ms;
    g_stage_launches[stage] = launches;
}

static thread_local double g_kernel_ms[4] = {0, 0, 0, 0};
void set_kernel_ms(int which, double ms) { g_kernel_ms[which] = ms; }
KernelTimer::KernelTimer(int w, cudaStream_t s) : which(w), stream(s)
{
    PNB_CUDA(cudaEventCreate(&a));
    PNB_CUDA(cudaEventCreate(&b));
    PNB_CUDA(cudaEventRecord(a, stream));
}
void KernelTimer::stop()
{
    if (!a) return;
    PNB_CUDA(cudaEventRecord(b, stream));
    PNB_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    PNB_CUDA(cudaEventElapsedTime(&ms, a, b));
    g_kernel_ms[which] = ms;
    cudaEventDestroy(a); cudaEventDestroy(b);
    a = b = nullptr;
}
KernelTimer::~KernelTimer()
{
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
}

StageTimer::StageTimer(int st, cudaStream_t s) : stage(st), stream(s), launches0(g_launches)
{
    PNB_CUDA(cudaEventCreate(&a));
    PNB_CUDA(cudaEventCreate(&b));
    PNB_CUDA(cudaEventRecord(a, stream));
}
void StageTimer::stop()
{
    if (!a) return;
    PNB_CUDA(cudaEventRecord(b, stream));
    PNB_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    PNB_CUDA(cudaEventElapsedTime(&ms, a, b));
    g_stage_ms[stage] = ms;
    g_stage_launches[stage] = g_launches - launches0;
    cudaEventDestroy(a); cudaEventDestroy(b);
    a = b = nullptr;
}
StageTimer::~StageTimer()
{
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
}

// ---- caching device allocator ---------------------------------------------------------------------
namespace {
struct BlockCache {
    std::mutex mu;
    std::map<std::pair<int, size_t>, std::vector<void *>> free_lists;   // (device, bytes) -> blocks
    size_t cached_bytes = 0;
    size_t limit = 0;           // idle blocks kept at most: a quarter of the device's memory (set on first use)
    void trim_locked()
    {
        for (auto &kv : free_lists)
            for (void *q : kv.second) cudaFree(q);
        free_lists.clear();
        cached_bytes = 0;
    }
};
BlockCache &block_cache() { static BlockCache c; return c; }
}  // namespace

void *cache_alloc(size_t bytes)
{
    int dev = 0;
    PNB_CUDA(cudaGetDevice(&dev));
    BlockCache &c = block_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.free_lists.find({dev, bytes});
    if (it != c.free_lists.end() && !it->second.empty()) {
        void *p = it->second.back();
        it->second.pop_back();
        c.cached_bytes -= bytes;
        return p;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {                     // give cached blocks back to the driver and retry once
        cudaGetLastError();
        cudaDeviceSynchronize();
        c.trim_locked();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw Error(PNB_ERR_MEMORY, "device allocation of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
    }
    return p;
}
void cache_free(void *p, size_t bytes)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
    BlockCache &c = block_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.limit == 0) {
        // idle blocks kept at most: PNB_CACHE_LIMIT_GB if set, else HALF of the device's memory -- the working set of one
        // step must fit or every step pays for cudaMalloc / cudaFree again (10^9 particles on 8 GPUs: 34 GB of neighbour
        // lists + 15 GB of build and search buffers per rank; with a limit of a quarter the step took 719 instead of
        // 497 ms).  torch's own caching allocator in the same process cannot reclaim what sits here;
        // pnb_release_cached_memory() returns it.
        size_t free_b = 0, total_b = 0;
        c.limit = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess ? total_b / 2 : (size_t)32 << 30;
        cudaGetLastError();
        if (const char *e = getenv("PNB_CACHE_LIMIT_GB")) {
            const double gb = atof(e);
            if (gb >= 0) c.limit = (size_t)(gb * 1073741824.0) + 1;
        }
    }
    // Past the limit the block that is being released goes back to the driver; the blocks already cached stay (dropping
    // the whole cache, as round 1 did, makes the NEXT step re-allocate everything).
    if (c.cached_bytes + bytes > c.limit) {
        cudaDeviceSynchronize();
        cudaFree(p);
        return;
    }
    c.free_lists[{dev, bytes}].push_back(p);
    c.cached_bytes += bytes;
}
size_t cached_bytes()
{
    BlockCache &c = block_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    return c.cached_bytes;
}
void cache_trim()
{
    BlockCache &c = block_cache();
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lock(c.mu);
    c.trim_locked();
}

// ---- streams ---------------------------------------------------------------------------------------
int g_opt_knn_blocks_per_sm = 0;    // pnb_set_option("knn_blocks_per_sm"): > 0 caps the resident blocks per SM of the search kernel
int g_opt_streams = 1;              // pnb_set_option("streams"): 1 = one non-blocking stream per calling thread, 0 = default stream
cudaStream_t lib_stream()
{
    if (!g_opt_streams) return nullptr;
    static thread_local cudaStream_t mine[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    cudaStream_t &s = mine[dev & 63];
    if (!s && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); s = nullptr; }
    return s;
}
ApiScope::ApiScope()
{
    cudaStream_t s = lib_stream();
    if (!s) return;
    static thread_local cudaEvent_t ev[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
    cudaEvent_t &e = ev[dev & 63];
    if (!e && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); e = nullptr; return; }
    // everything the caller enqueued on the default stream so far (torch's tensors) precedes our work
    if (cudaEventRecord(e, cudaStreamLegacy) == cudaSuccess) cudaStreamWaitEvent(s, e, 0);
    cudaGetLastError();
}
ApiScope::~ApiScope()
{
    cudaStream_t s = lib_stream();
    if (s) { cudaStreamSynchronize(s); cudaGetLastError(); }
}

// a second, non-blocking stream per device for copies that run beside the work on the library's stream
static cudaStream_t copy_stream()
{
    static cudaStream_t sides[64] = {nullptr};
    int dev = 0;
    PNB_CUDA(cudaGetDevice(&dev));
    cudaStream_t &side = sides[dev & 63];
    if (!side) PNB_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    return side;
}
// an event that is waited for (and destroyed) when it goes out of scope, so that a buffer declared BEFORE it
// is never released while the copy it marks is still in flight
struct PendingCopy {
    cudaEvent_t &ev;
    ~PendingCopy() { if (ev) { cudaEventSynchronize(ev); cudaEventDestroy(ev); ev = nullptr; } }
};

static void require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        throw Error(PNB_ERR_CUDA, "no CUDA device available: pnb_b200 has no CPU fallback");
    }
}

// smooth.h:71-102: periodicity is honoured only if the particles fit inside one period
static double effective_period(const Tree &t, double period, int *ignored)
{
    if (ignored) *ignored = 0;
    if (!(period > 0)) return 0;
    for (int d = 0; d < 3; ++d)
        if (t.root_max[d] - t.root_min[d] > period) {
            if (ignored) *ignored = 1;
            return 0;
        }
    return period;
}

template <typename F>
static int guarded(F &&f)
{
    try {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
            ApiScope scope;
            f();
        } else {
            cudaGetLastError();
            f();
        }
        return PNB_OK;
    } catch (const Error &e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::bad_alloc &) {
        set_last_error("out of host memory");
        return PNB_ERR_MEMORY;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return PNB_ERR_RUNTIME;
    }
}

static const ArrRef &need(const Tree &t, int id, const char *name)
{
    if (!t.arr[id].set()) throw Error(PNB_ERR_VALUE, std::string(name) + " array has not been set");
    return t.arr[id];
}

// fetch a registered 1-D / N x ncols array into dense device columns in TREE order
static DevBuf fetch_tree_order(Tree &t, const ArrRef &r)
{
    const size_t elem = tsize(r.dtype);
    DevBuf file_cols(elem * r.ncols * t.n), tree_cols(elem * r.ncols * t.n);
    fetch_columns(r.ptr, r.stride0, r.stride1, r.ncols, r.dtype, r.mem, t.n, file_cols.p, t.stream);
    permute_to_tree(t, file_cols.p, tree_cols.p, r.ncols, (int)elem, t.stream);
    PNB_CUDA(cudaStreamSynchronize(t.stream));
    return tree_cols;
}
static void store_from_tree_order(Tree &t, const ArrRef &r, const void *tree_cols)
{
    const size_t elem = tsize(r.dtype);
    DevBuf file_cols(elem * r.ncols * t.n);
    permute_to_file(t, tree_cols, file_cols.p, r.ncols, (int)elem, t.stream);
    store_columns(r.ptr, r.stride0, r.stride1, r.ncols, r.dtype, r.mem, t.n, file_cols.p, t.stream);
    PNB_CUDA(cudaStreamSynchronize(t.stream));
}

// bitwise comparison of two arrays of 4-byte words; with a mask (one byte per `words_per_elem` words)
// only the selected elements are compared
__global__ void k_all_equal(const unsigned char *a, const unsigned char *b, size_t bytes, const uint8_t *mask,
                            int words_per_elem, int *differs)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint32_t *a4 = reinterpret_cast<const uint32_t *>(a), *b4 = reinterpret_cast<const uint32_t *>(b);
    bool d = false;
    size_t done = 0;
    if (!mask && (((uintptr_t)a | (uintptr_t)b) & 15) == 0) {       // the usual case: 16-byte loads, two in flight per array
        const uint4 *av = reinterpret_cast<const uint4 *>(a), *bv = reinterpret_cast<const uint4 *>(b);
        const size_t nv = bytes / 16;
        size_t j = i;
        for (; j + stride < nv; j += 2 * stride) {
            const uint4 x0 = av[j], y0 = bv[j], x1 = av[j + stride], y1 = bv[j + stride];
            d |= (x0.x != y0.x) | (x0.y != y0.y) | (x0.z != y0.z) | (x0.w != y0.w) | (x1.x != y1.x) | (x1.y != y1.y) | (x1.z != y1.z)
                 | (x1.w != y1.w);
        }
        for (; j < nv; j += stride) {
            const uint4 x0 = av[j], y0 = bv[j];
            d |= (x0.x != y0.x) | (x0.y != y0.y) | (x0.z != y0.z) | (x0.w != y0.w);
        }
        done = nv * 4;
    }
    for (size_t j = done + i; j < bytes / 4; j += stride)
        if (!mask || mask[j / words_per_elem]) d |= a4[j] != b4[j];
    if (d) *differs = 1;
}
static bool device_equal(Tree &t, const void *a, const void *b, size_t bytes, bool masked = false)
{
    DevBuf flag(sizeof(int));
    PNB_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), t.stream));
    const uint8_t *mask = (masked && t.query_mask.p) ? t.query_mask.as<uint8_t>() : nullptr;
    PNB_LAUNCH(k_all_equal, 1184, 256, 0, t.stream, (const unsigned char *)a, (const unsigned char *)b, bytes, mask,
               (int)(tsize(t.dtype) / 4), flag.as<int>());
    int h = 0;
    PNB_CUDA(cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, t.stream));
    PNB_CUDA(cudaStreamSynchronize(t.stream));
    return h == 0;
}

template <typename T>
__global__ void k_any_differs_from_first(const T *m, int64_t n, int *differs)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const T m0 = m[0];
    bool d = false;
    for (int64_t j = i; j < n; j += stride) d |= m[j] != m0;
    if (d) *differs = 1;
}
static void update_mass_uniformity(Tree &t)
{
    t.mass_is_uniform = false;
    if (!t.has_mass) return;
    DevBuf flag(sizeof(int));
    PNB_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), t.stream));
    if (t.dtype == PNB_F64) PNB_LAUNCH((k_any_differs_from_first<double>), 1184, 256, 0, t.stream, t.mass.as<double>(), t.n, flag.as<int>());
    else PNB_LAUNCH((k_any_differs_from_first<float>), 1184, 256, 0, t.stream, t.mass.as<float>(), t.n, flag.as<int>());
    int h = 0;
    double m0 = 0;
    PNB_CUDA(cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, t.stream));
    if (t.dtype == PNB_F64) {
        PNB_CUDA(cudaMemcpyAsync(&m0, t.mass.p, sizeof(double), cudaMemcpyDeviceToHost, t.stream));
        PNB_CUDA(cudaStreamSynchronize(t.stream));
    } else {
        float f = 0;
        PNB_CUDA(cudaMemcpyAsync(&f, t.mass.p, sizeof(float), cudaMemcpyDeviceToHost, t.stream));
        PNB_CUDA(cudaStreamSynchronize(t.stream));
        m0 = f;
    }
    t.mass_is_uniform = h == 0;
    t.uniform_mass = m0;
}

int g_opt_neighbour_lists = -1;     // pnb_set_option("neighbour_lists"): -1 auto (keep them if they fit), 0 never, 1 always
// pnb_set_option("tree_build"): how the device tree is built (results never depend on it).
//   2 (default)  median-split kd-tree from three per-axis sorted lists: one fused global pass per level (stable partition with
//                decoupled look-back), the lowest 6 levels per 2048-particle sub-tree in shared memory (tree_build.cu)
//   3            the same tree by mark + CUB prefix sum + partition per level (the round-1 formulation)
//   4            the same tree with the fused global pass for every level
//   1            the same, but the lowest 6-7 levels per sub-tree in shared memory (tree_morton.cu: k_build_segment)
//   0            one Morton-key sort for the upper levels + shared-memory kd levels below
// Measured at 16.7 M float64 particles (profiles/r2_tree_build_variants.md): build 9.2 (2) / 15.3 (3) / 10.8 (4) / 16.1 (1) /
// 6.2 ms (0); the Morton tree is the cheapest to build but its buckets cost the neighbour search +25 %.
int g_opt_tree_build = 2;
int g_opt_gather_walk = 0;          // pnb_set_option("gather_walk"): 1 = reductions always walk the tree
static struct OptionsFromEnv {
    OptionsFromEnv()
    {
        if (const char *e = getenv("PNB_NEIGHBOUR_LISTS")) g_opt_neighbour_lists = atoi(e);
        if (const char *e = getenv("PNB_GATHER_WALK")) g_opt_gather_walk = atoi(e);
        if (const char *e = getenv("PNB_TREE_BUILD")) g_opt_tree_build = atoi(e);
        if (const char *e = getenv("PNB_STREAMS")) g_opt_streams = atoi(e);
    }
} g_options_from_env;

// The resident neighbour lists serve a reduction iff they come from this tree's own kNN pass with the same k
// (the caller has verified that the smooth array is that pass's result and that the period matches).
// PNB_GATHER=walk forces the tree-walking gather (tests compare the two paths).
static bool lists_usable(const Tree &t, int k)
{
    if (g_opt_gather_walk) return false;
    return t.nn_valid && t.nn_k == k && t.nn_ids.p != nullptr;
}

static void populate(Tree &t, int propid, int k, int kernel_id, double period_in, int *period_ignored)
{
    require_device();
    if (kernel_id != PNB_KERNEL_CUBIC_SPLINE && kernel_id != PNB_KERNEL_WENDLAND_C2)
        throw Error(PNB_ERR_VALUE, "unknown SPH kernel id");
    if (k > t.n) throw Error(PNB_ERR_VALUE, "Number of smoothing particles exceeds number of particles in tree");
    if (k < 1) throw Error(PNB_ERR_VALUE, "Number of smoothing particles must be positive");
    const double period = effective_period(t, period_in, period_ignored);
    const KernelParams kp = make_kernel_params(kernel_id, k);
    const size_t ts = tsize(t.dtype);
    StageTimer timer(1, t.stream);

    if (propid == PNB_PROP_HSM) {
        const ArrRef &sm = need(t, PNB_ARR_SMOOTH, "smooth");       // kdmain.cpp:633-639
        run_knn(t, k, period, /*fuse_rho=*/t.has_mass, kp, /*keep_ids=*/neighbour_lists_fit(t, k, false), false);
        store_from_tree_order(t, sm, t.smooth_t.p);
        timer.stop();
        return;
    }
    if (propid < PNB_PROP_RHO || propid > PNB_PROP_QTYCURL) throw Error(PNB_ERR_VALUE, "unknown smoothing property id");

    // sim['smooth'] -> sim['rho'] with pinned host arrays: the density that rode along the kNN pass starts its way
    // back to the host while the caller's smooth array travels the other way to be checked against the tree's
    // own (PCIe is full duplex); if the check fails the gather pass below overwrites it
    const bool period_matches = t.smooth_period == (period > 0 ? period : 0);
    DevBuf early_rho;
    cudaEvent_t early_done = nullptr;
    PendingCopy early_guard{early_done};
    if (propid == PNB_PROP_RHO && t.smooth_valid && t.smooth_k == k && period_matches && t.rho_valid
        && t.rho_kernel == kernel_id && t.arr[PNB_ARR_RHO].set()) {
        const ArrRef &o = t.arr[PNB_ARR_RHO];
        if (o.mem == PNB_HOST && o.ncols == 1 && o.dtype == t.dtype && o.stride0 == (int64_t)ts
            && ts * (size_t)t.n >= ((size_t)4 << 20) && host_is_pinned(o.ptr)) {
            early_rho.alloc(ts * t.n);
            permute_to_file(t, t.rho_t.p, early_rho.p, 1, (int)ts, t.stream);
            cudaEvent_t ready = nullptr;
            PNB_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
            PNB_CUDA(cudaEventRecord(ready, t.stream));
            cudaStream_t side = copy_stream();
            PNB_CUDA(cudaStreamWaitEvent(side, ready, 0));
            PNB_CUDA(cudaEventDestroy(ready));
            PNB_CUDA(cudaEventCreateWithFlags(&early_done, cudaEventDisableTiming));
            PNB_CUDA(cudaMemcpyAsync(o.ptr, early_rho.p, ts * (size_t)t.n, cudaMemcpyDeviceToHost, side));
            PNB_CUDA(cudaEventRecord(early_done, side));
        }
    }

    const ArrRef &smr = need(t, PNB_ARR_SMOOTH, "smooth");
    DevBuf smooth_tree = fetch_tree_order(t, smr);
    // is the caller's smooth array the one this tree produced (the usual sim['smooth'] -> sim['rho'] flow)?
    const bool own_smooth = t.smooth_valid && t.smooth_k == k && period_matches
                            && device_equal(t, smooth_tree.p, t.smooth_t.p, ts * t.n, /*masked=*/true);
    if (early_done) {                                               // the early copy has to land either way
        PNB_CUDA(cudaEventSynchronize(early_done));
        PNB_CUDA(cudaEventDestroy(early_done));
        early_done = nullptr;
        if (own_smooth) { timer.stop(); return; }
    }

    if (propid == PNB_PROP_RHO) {
        const ArrRef &out = need(t, PNB_ARR_RHO, "rho");
        if (own_smooth && t.rho_valid && t.rho_kernel == kernel_id) {
            store_from_tree_order(t, out, t.rho_t.p);                // density fused with the kNN pass
        } else {
            DevBuf rho_tree(ts * t.n);
            if (own_smooth && lists_usable(t, k))
                run_gather_lists(t, propid, k, period, kp, smooth_tree.p, nullptr, nullptr, t.dtype, rho_tree.p);
            else if (run_gather(t, propid, k, period, kp, smooth_tree.p, nullptr, nullptr, t.dtype, rho_tree.p))
                throw Error(PNB_ERR_RUNTIME, "Buffer overflow in smoothing operation. This probably means that your smoothing lengths are too large compared to the number of neighbours you specified.");
            store_from_tree_order(t, out, rho_tree.p);
        }
        timer.stop();
        return;
    }

    const ArrRef &rhor = need(t, PNB_ARR_RHO, "rho");
    const ArrRef &qty = need(t, PNB_ARR_QTY, "qty");
    const ArrRef &out = need(t, PNB_ARR_QTY_SM, "qty_sm");
    const bool nd = propid == PNB_PROP_QTYMEAN_ND || propid == PNB_PROP_QTYDISP_ND || propid == PNB_PROP_QTYDIV
                    || propid == PNB_PROP_QTYCURL;
    const bool out_nd = propid == PNB_PROP_QTYMEAN_ND || propid == PNB_PROP_QTYCURL;
    if (qty.ncols != (nd ? 3 : 1)) throw Error(PNB_ERR_VALUE, "qty array has the wrong shape for this operation");
    if (out.ncols != (out_nd ? 3 : 1)) throw Error(PNB_ERR_VALUE, "qty_sm array has the wrong shape for this operation");
    if (qty.dtype != out.dtype) throw Error(PNB_ERR_TYPE, "qty and qty_sm must have the same dtype");
    DevBuf rho_tree = fetch_tree_order(t, rhor);
    DevBuf qty_tree = fetch_tree_order(t, qty);
    DevBuf out_tree(tsize(out.dtype) * out.ncols * t.n);
    if (own_smooth && lists_usable(t, k))
        run_gather_lists(t, propid, k, period, kp, smooth_tree.p, rho_tree.p, qty_tree.p, qty.dtype, out_tree.p);
    else if (run_gather(t, propid, k, period, kp, smooth_tree.p, rho_tree.p, qty_tree.p, qty.dtype, out_tree.p))
        throw Error(PNB_ERR_RUNTIME, "Buffer overflow in smoothing operation. This probably means that your smoothing lengths are too large compared to the number of neighbours you specified.");
    store_from_tree_order(t, out, out_tree.p);
    timer.stop();
}

// rows of the resident neighbour lists for the particles with FILE index first .. first+count-1
template <typename T>
__global__ void k_nn_rows(int64_t first, int64_t count, int k, int K2, const uint32_t *__restrict__ inv,
                          const uint32_t *__restrict__ order, const uint32_t *__restrict__ ids,
                          const T *__restrict__ d2, const T *__restrict__ smooth_t, int64_t *__restrict__ idx_out,
                          T *__restrict__ d2_out, T *__restrict__ h_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= count) return;
    const uint32_t p = inv[first + r];
    const size_t base = (size_t)(p >> 5) * K2 * 32 + (p & 31u);
    for (int e = lane; e < k; e += 32) {
        idx_out[r * k + e] = (int64_t)order[ids[base + (size_t)e * 32]];
        d2_out[r * k + e] = d2[base + (size_t)e * 32];
    }
    if (lane == 0 && h_out) h_out[r] = smooth_t[p];
}
__global__ void k_invert_order(const uint32_t *__restrict__ order, uint32_t *__restrict__ inv, int64_t n)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < n) inv[order[p]] = (uint32_t)p;
}

static void knn_start(Tree &t, int k, double period_in, int *period_ignored)
{
    require_device();
    if (k > t.n) throw Error(PNB_ERR_VALUE, "Number of smoothing particles exceeds number of particles in tree");
    if (k < 1) throw Error(PNB_ERR_VALUE, "Number of smoothing particles must be positive");
    const double period = effective_period(t, period_in, period_ignored);
    StageTimer timer(1, t.stream);
    run_knn(t, k, period, false, make_kernel_params(0, k), /*keep_ids=*/true, /*keep_d2=*/true);
    if (!t.inv_order.p) {
        t.inv_order.alloc(sizeof(uint32_t) * (size_t)t.n);
        PNB_LAUNCH(k_invert_order, (unsigned)((t.n + 255) / 256), 256, 0, t.stream, t.order.as<uint32_t>(),
                   t.inv_order.as<uint32_t>(), t.n);
    }
    if (t.arr[PNB_ARR_SMOOTH].set()) store_from_tree_order(t, t.arr[PNB_ARR_SMOOTH], t.smooth_t.p);   // smooth.h:429
    PNB_CUDA(cudaStreamSynchronize(t.stream));
    timer.stop();
}

static void knn_rows(Tree &t, int64_t first, int64_t count, int64_t *idx_out, void *d2_out, void *h_out)
{
    require_device();
    if (!t.nn_valid || !t.nn_d2.p || !t.inv_order.p)
        throw Error(PNB_ERR_VALUE, "pnb_knn_rows: no neighbour lists; call pnb_knn_start first");
    if (first < 0 || count < 0 || first + count > t.n) throw Error(PNB_ERR_VALUE, "pnb_knn_rows: row range outside the particle set");
    const int k = t.nn_k;
    const size_t ts = tsize(t.dtype);
    // staged through a bounded device buffer (<= 256 MiB of rows at a time)
    const int64_t chunk = std::max<int64_t>(1, ((int64_t)256 << 20) / ((int64_t)k * (8 + (int64_t)ts)));
    DevBuf idx(sizeof(int64_t) * (size_t)std::min(chunk, std::max<int64_t>(count, 1)) * k),
           d2(ts * (size_t)std::min(chunk, std::max<int64_t>(count, 1)) * k), hh(ts * (size_t)std::min(chunk, std::max<int64_t>(count, 1)));
    for (int64_t done = 0; done < count; done += chunk) {
        const int64_t m = std::min(chunk, count - done);
        const unsigned grid = (unsigned)((m + 7) / 8);
        if (t.dtype == PNB_F64)
            PNB_LAUNCH((k_nn_rows<double>), grid, 256, 0, t.stream, first + done, m, k, t.nn_K2, t.inv_order.as<uint32_t>(),
                       t.order.as<uint32_t>(), t.nn_ids.as<uint32_t>(), t.nn_d2.as<double>(), t.smooth_t.as<double>(),
                       idx.as<int64_t>(), d2.as<double>(), hh.as<double>());
        else
            PNB_LAUNCH((k_nn_rows<float>), grid, 256, 0, t.stream, first + done, m, k, t.nn_K2, t.inv_order.as<uint32_t>(),
                       t.order.as<uint32_t>(), t.nn_ids.as<uint32_t>(), t.nn_d2.as<float>(), t.smooth_t.as<float>(),
                       idx.as<int64_t>(), d2.as<float>(), hh.as<float>());
        PNB_CUDA(cudaMemcpyAsync(idx_out + done * k, idx.p, sizeof(int64_t) * (size_t)m * k, cudaMemcpyDeviceToHost, t.stream));
        PNB_CUDA(cudaMemcpyAsync((char *)d2_out + ts * (size_t)done * k, d2.p, ts * (size_t)m * k, cudaMemcpyDeviceToHost, t.stream));
        if (h_out) PNB_CUDA(cudaMemcpyAsync((char *)h_out + ts * (size_t)done, hh.p, ts * (size_t)m, cudaMemcpyDeviceToHost, t.stream));
        PNB_CUDA(cudaStreamSynchronize(t.stream));
    }
}

// one thread per bucket: could the k-neighbour ball of any of its particles reach a face of the cuboid [lo, hi)?
// The smallest ancestor of the bucket that holds >= k particles contains the query and k particles, so the ball
// radius is at most that ancestor's diagonal; a bucket further than that from every face is safely interior.
template <typename T>
__global__ void k_near_faces(int64_t nsplit, int levels, int k, const int32_t *__restrict__ bucket_start,
                             const T *__restrict__ box, double lx, double ly, double lz, double hx, double hy, double hz,
                             uint8_t *__restrict__ flags_tree)
{
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= nsplit) return;
    const int start = bucket_start[b], cnt = bucket_start[b + 1] - start;
    if (cnt <= 0) return;
    int j = 0;
    for (; j < levels; ++j) {
        const int64_t a = b >> j;
        const int64_t first = a << j, last = min(nsplit, (a + 1) << j);
        if (bucket_start[last] - bucket_start[first] >= k) break;
    }
    const T *ab = box + ((nsplit + b) >> j) * 6;
    const double dx = (double)ab[3] - (double)ab[0], dy = (double)ab[4] - (double)ab[1], dz = (double)ab[5] - (double)ab[2];
    const double reach = sqrt(dx * dx + dy * dy + dz * dz) * (1.0 + 1e-6);
    const T *bb = box + (nsplit + b) * 6;
    const double gap = fmin(fmin(fmin((double)bb[0] - lx, hx - (double)bb[3]), fmin((double)bb[1] - ly, hy - (double)bb[4])),
                            fmin((double)bb[2] - lz, hz - (double)bb[5]));
    const uint8_t near = !(gap >= reach) ? 1 : 0;          // (also set when the whole tree holds fewer than k particles)
    for (int i = 0; i < cnt; ++i) flags_tree[start + i] = near;
}

}  // namespace pnb

using namespace pnb;

extern "C" {

const char *pnb_last_error(void) { return g_error.c_str(); }
const char *pnb_version(void) { return "pnb_b200 0.1.0 (sm_100a)"; }

int pnb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pnb_set_option(const char *name, int value)
{
    return guarded([&] {
        if (!name) throw Error(PNB_ERR_VALUE, "option name is null");
        if (!strcmp(name, "neighbour_lists")) g_opt_neighbour_lists = value;
        else if (!strcmp(name, "gather_walk")) g_opt_gather_walk = value;
        else if (!strcmp(name, "tree_build")) g_opt_tree_build = value;
        else if (!strcmp(name, "streams")) g_opt_streams = value;
        else if (!strcmp(name, "knn_blocks_per_sm")) g_opt_knn_blocks_per_sm = value;
        else throw Error(PNB_ERR_VALUE, std::string("unknown option ") + name);
    });
}

int pnb_set_device(int device)
{
    return guarded([&] { PNB_CUDA(cudaSetDevice(device)); });
}

pnb_tree *pnb_tree_create(const void *pos, int64_t pos_stride0, int64_t pos_stride1, const void *mass,
                          int64_t mass_stride, int64_t n, int dtype, int leafsize, int mem)
{
    Tree *t = nullptr;
    int rc = guarded([&] {
        require_device();
        if (dtype != PNB_F32 && dtype != PNB_F64) throw Error(PNB_ERR_VALUE, "Unsupported array dtype for kdtree");
        if (!pos) throw Error(PNB_ERR_VALUE, "pos is null");
        if (n < 1) throw Error(PNB_ERR_VALUE, "kd-tree needs at least one particle");
        if (leafsize < 1) throw Error(PNB_ERR_VALUE, "leafsize must be positive");
        t = new Tree();
        t->dtype = dtype; t->n = n; t->user_leafsize = leafsize;
        PNB_CUDA(cudaGetDevice(&t->device));
        StageTimer timer(0, t->stream);
        const size_t ts = tsize(dtype);
        DevBuf cols(ts * 3 * n), mass_dev;
        PendingCopy pending{t->mass_ready};  // declared after mass_dev: the copy has landed before the buffer is released
        fetch_columns(pos, pos_stride0, pos_stride1, 3, dtype, mem, n, cols.p, t->stream);
        if (mass) {
            mass_dev.alloc(ts * n);
            if (mem == PNB_HOST && mass_stride == (int64_t)ts && ts * (size_t)n >= ((size_t)4 << 20) && host_is_pinned(mass)) {
                // pinned host masses: the copy rides on a second stream behind the sorts of the build and is
                // only waited for where the tree-ordered arrays are gathered
                cudaStream_t side = copy_stream();
                PNB_CUDA(cudaEventCreateWithFlags(&t->mass_ready, cudaEventDisableTiming));
                PNB_CUDA(cudaMemcpyAsync(mass_dev.p, mass, ts * (size_t)n, cudaMemcpyHostToDevice, side));
                PNB_CUDA(cudaEventRecord(t->mass_ready, side));
            } else {
                fetch_columns(mass, mass_stride, 0, 1, dtype, mem, n, mass_dev.p, t->stream);
            }
        }
        if (g_opt_tree_build == 0) build_tree_morton(*t, cols, mass ? mass_dev.p : nullptr);
        else build_tree(*t, cols, mass ? mass_dev.p : nullptr, packed_shape(n), true);
        update_mass_uniformity(*t);
        timer.stop();
    });
    if (rc != PNB_OK) { delete t; return nullptr; }
    return reinterpret_cast<pnb_tree *>(t);
}

void pnb_tree_destroy(pnb_tree *h) { delete reinterpret_cast<Tree *>(h); }

// a null handle is a caller error, not a crash
#define PNB_NEED_TREE(h)                                                     \
    do {                                                                     \
        if (!(h)) { set_last_error("tree handle is null"); return PNB_ERR_VALUE; } \
        int dev_ = -1;                                                       \
        const int tdev_ = reinterpret_cast<const Tree *>(h)->device;         \
        if (tdev_ >= 0 && cudaGetDevice(&dev_) == cudaSuccess && dev_ != tdev_) {                                     \
            set_last_error("the tree lives on CUDA device " + std::to_string(tdev_) + " but the calling thread's current " \
                           "device is " + std::to_string(dev_) + ": call pnb_set_device first");                     \
            return PNB_ERR_VALUE;                                            \
        }                                                                    \
    } while (0)

int64_t pnb_tree_num_particles(const pnb_tree *h) { PNB_NEED_TREE(h); return reinterpret_cast<const Tree *>(h)->n; }

int pnb_tree_bounds(const pnb_tree *h, double *out6)
{
    PNB_NEED_TREE(h);
    const Tree *t = reinterpret_cast<const Tree *>(h);
    for (int d = 0; d < 3; ++d) { out6[d] = t->root_min[d]; out6[3 + d] = t->root_max[d]; }
    return PNB_OK;
}

int64_t pnb_tree_node_count(const pnb_tree *h)
{
    PNB_NEED_TREE(h);
    const Tree *t = reinterpret_cast<const Tree *>(h);
    int64_t n = t->n, l = 1;                      // kd.cpp:98-111
    while (n > t->user_leafsize) { n >>= 1; l <<= 1; }
    return l << 1;
}

int pnb_tree_order(pnb_tree *h, int64_t *out)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] {
        std::vector<uint32_t> o((size_t)t->n);
        PNB_CUDA(cudaStreamSynchronize(t->stream));
        PNB_CUDA(cudaMemcpy(o.data(), t->order.p, sizeof(uint32_t) * (size_t)t->n, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < t->n; ++i) out[i] = (int64_t)o[(size_t)i];
    });
}

int64_t pnb_tree_num_buckets(const pnb_tree *h) { PNB_NEED_TREE(h); return reinterpret_cast<const Tree *>(h)->nsplit; }

int pnb_tree_buckets(pnb_tree *h, int64_t *start, double *bounds)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] {
        std::vector<int32_t> s((size_t)t->nsplit + 1);
        PNB_CUDA(cudaMemcpy(s.data(), t->bucket_start.p, sizeof(int32_t) * s.size(), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < s.size(); ++i) start[i] = s[i];
        const size_t nb = (size_t)t->nsplit * 6;
        if (t->dtype == PNB_F64) {
            PNB_CUDA(cudaMemcpy(bounds, t->box.as<double>() + nb, sizeof(double) * nb, cudaMemcpyDeviceToHost));
        } else {
            std::vector<float> b(nb);
            PNB_CUDA(cudaMemcpy(b.data(), t->box.as<float>() + nb, sizeof(float) * nb, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < nb; ++i) bounds[i] = (double)b[i];
        }
    });
}

int pnb_tree_node_bounds(pnb_tree *h, double *bounds)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] {
        const size_t nb = (size_t)t->nsplit * 2 * 6;
        if (t->dtype == PNB_F64) {
            PNB_CUDA(cudaMemcpy(bounds, t->box.p, sizeof(double) * nb, cudaMemcpyDeviceToHost));
        } else {
            std::vector<float> b(nb);
            PNB_CUDA(cudaMemcpy(b.data(), t->box.p, sizeof(float) * nb, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < nb; ++i) bounds[i] = (double)b[i];
        }
    });
}

int pnb_tree_export(pnb_tree *h, void *kdnodes_out, int64_t n_nodes, int64_t *offsets_out)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] { require_device(); export_reference_tree(*t, kdnodes_out, n_nodes, offsets_out); });
}

int pnb_set_array(pnb_tree *h, int id, void *ptr, int64_t stride0, int64_t stride1, int ncols, int dtype, int mem)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] {
        if (id < 0 || id > 4) throw Error(PNB_ERR_VALUE, "invalid array id");
        if (dtype != PNB_F32 && dtype != PNB_F64) throw Error(PNB_ERR_VALUE, "Unsupported array dtype for kdtree");
        if (id <= PNB_ARR_MASS) {
            if (dtype != t->dtype) throw Error(PNB_ERR_TYPE, "array dtype must match the dtype of the positions");
            if (ncols != 1) throw Error(PNB_ERR_VALUE, "array must be one-dimensional");
        } else if (ncols != 1 && ncols != 3) {
            throw Error(PNB_ERR_VALUE, "qty array must be one-dimensional or N x 3");
        }
        if (!ptr) throw Error(PNB_ERR_VALUE, "array pointer is null");
        ArrRef r;
        r.ptr = ptr; r.stride0 = stride0; r.stride1 = stride1; r.ncols = ncols; r.dtype = dtype; r.mem = mem;
        if (id == PNB_ARR_MASS) {
            // mass lives on the device in tree order; refresh it (the reference re-reads the borrowed buffer)
            require_device();
            DevBuf m = fetch_tree_order(*t, r);
            const bool same = t->has_mass && device_equal(*t, m.p, t->mass.p, tsize(dtype) * (size_t)t->n);
            if (!same) {
                t->mass = std::move(m);
                t->pos4.release();
                t->has_mass = true;
                t->rho_valid = false;
                update_mass_uniformity(*t);
            }
        }
        t->arr[id] = r;
    });
}

__global__ void k_u8_to_tree(const uint32_t *order, const uint8_t *in_file, uint8_t *out_tree, int64_t n, int invert)
{
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < n) out_tree[p] = ((in_file[order[p]] != 0) != (invert != 0)) ? 1 : 0;
}
__global__ void k_u8_to_file(const uint32_t *order, const uint8_t *in_tree, uint8_t *out_file, int64_t n)
{
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < n) out_file[order[p]] = in_tree[p];
}

static void set_query_mask(Tree *t, const uint8_t *mask, int mem, int invert, int keep_results)
{
    if (!keep_results) { t->smooth_valid = false; t->rho_valid = false; t->nn_valid = false; }
    if (!mask) {
        t->query_mask.release();
        return;
    }
    require_device();
    DevBuf file_mask;
    const uint8_t *src = mask;
    if (mem == PNB_HOST) {
        file_mask.alloc((size_t)t->n);
        PNB_CUDA(cudaMemcpyAsync(file_mask.p, mask, (size_t)t->n, cudaMemcpyHostToDevice, t->stream));
        src = file_mask.as<uint8_t>();
    }
    if (!t->query_mask.p) t->query_mask.alloc((size_t)t->n);
    PNB_LAUNCH(k_u8_to_tree, (unsigned)((t->n + 255) / 256), 256, 0, t->stream, t->order.as<uint32_t>(), src,
               t->query_mask.as<uint8_t>(), t->n, invert);
    PNB_CUDA(cudaStreamSynchronize(t->stream));
}

int pnb_set_query_mask(pnb_tree *h, const uint8_t *mask, int mem)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] {
        if (!mask && !t->query_mask.p) return;           // nothing changes: results stay valid
        set_query_mask(t, mask, mem, 0, 0);
    });
}

int pnb_set_query_mask2(pnb_tree *h, const uint8_t *mask, int mem, int invert, int keep_results)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] { set_query_mask(t, mask, mem, invert, keep_results); });
}

int pnb_mark_in_balls(pnb_tree *h, const void *centres, int64_t stride0, int64_t stride1, const void *radii,
                      int64_t rstride, int64_t nq, double period_in, uint8_t *flags_out, int mem)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] {
        require_device();
        if (nq < 0) throw Error(PNB_ERR_VALUE, "negative ball count");
        if (!flags_out) throw Error(PNB_ERR_VALUE, "flags pointer is null");
        const double period = effective_period(*t, period_in, nullptr);
        const size_t ts = tsize(t->dtype);
        DevBuf flags_tree((size_t)t->n), flags_file((size_t)t->n);
        PNB_CUDA(cudaMemsetAsync(flags_tree.p, 0, (size_t)t->n, t->stream));
        if (nq > 0) {
            DevBuf c(ts * 3 * nq), r(ts * nq);
            fetch_columns(centres, stride0, stride1, 3, t->dtype, mem, nq, c.p, t->stream);
            fetch_columns(radii, rstride, 0, 1, t->dtype, mem, nq, r.p, t->stream);
            run_mark_in_balls(*t, c.p, r.p, nq, period, flags_tree.as<uint8_t>());
            PNB_CUDA(cudaStreamSynchronize(t->stream));
        }
        uint8_t *dst = mem == PNB_DEVICE ? flags_out : flags_file.as<uint8_t>();
        PNB_LAUNCH(k_u8_to_file, (unsigned)((t->n + 255) / 256), 256, 0, t->stream, t->order.as<uint32_t>(),
                   flags_tree.as<uint8_t>(), dst, t->n);
        if (mem != PNB_DEVICE)
            PNB_CUDA(cudaMemcpyAsync(flags_out, flags_file.p, (size_t)t->n, cudaMemcpyDeviceToHost, t->stream));
        PNB_CUDA(cudaStreamSynchronize(t->stream));
    });
}

int pnb_mark_near_faces(pnb_tree *h, int k, const double *lo3, const double *hi3, uint8_t *flags_out, int mem)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] {
        require_device();
        if (!lo3 || !hi3 || !flags_out) throw Error(PNB_ERR_VALUE, "pointer is null");
        if (k < 1) throw Error(PNB_ERR_VALUE, "Number of smoothing particles must be positive");
        DevBuf flags_tree((size_t)t->n), flags_file((size_t)t->n);
        PNB_CUDA(cudaMemsetAsync(flags_tree.p, 0, (size_t)t->n, t->stream));
        const unsigned grid = (unsigned)((t->nsplit + 127) / 128);
        if (t->dtype == PNB_F64)
            PNB_LAUNCH((k_near_faces<double>), grid, 128, 0, t->stream, t->nsplit, t->levels, k, t->bucket_start.as<int32_t>(),
                       t->box.as<double>(), lo3[0], lo3[1], lo3[2], hi3[0], hi3[1], hi3[2], flags_tree.as<uint8_t>());
        else
            PNB_LAUNCH((k_near_faces<float>), grid, 128, 0, t->stream, t->nsplit, t->levels, k, t->bucket_start.as<int32_t>(),
                       t->box.as<float>(), lo3[0], lo3[1], lo3[2], hi3[0], hi3[1], hi3[2], flags_tree.as<uint8_t>());
        uint8_t *dst = mem == PNB_DEVICE ? flags_out : flags_file.as<uint8_t>();
        PNB_LAUNCH(k_u8_to_file, (unsigned)((t->n + 255) / 256), 256, 0, t->stream, t->order.as<uint32_t>(),
                   flags_tree.as<uint8_t>(), dst, t->n);
        if (mem != PNB_DEVICE)
            PNB_CUDA(cudaMemcpyAsync(flags_out, flags_file.p, (size_t)t->n, cudaMemcpyDeviceToHost, t->stream));
        PNB_CUDA(cudaStreamSynchronize(t->stream));
    });
}

int pnb_populate(pnb_tree *h, int propid, int k, int kernel_id, double period, int *period_ignored)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] { populate(*t, propid, k, kernel_id, period, period_ignored); });
}

int pnb_knn_start(pnb_tree *h, int k, double period, int *period_ignored)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] { knn_start(*t, k, period, period_ignored); });
}

int pnb_knn_rows(pnb_tree *h, int64_t first, int64_t count, int64_t *idx_out, void *d2_out, void *h_out)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] { knn_rows(*t, first, count, idx_out, d2_out, h_out); });
}

int pnb_knn(pnb_tree *h, int k, double period_in, int64_t *idx_out, void *d2_out, void *h_out, int *period_ignored)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] {
        knn_start(*t, k, period_in, period_ignored);
        knn_rows(*t, 0, t->n, idx_out, d2_out, h_out);
    });
}

int pnb_gather_counts(pnb_tree *h, double period_in, int32_t *counts_out, int mem)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] {
        require_device();
        if (!counts_out) throw Error(PNB_ERR_VALUE, "counts pointer is null");
        const double period = effective_period(*t, period_in, nullptr);
        const ArrRef &smr = need(*t, PNB_ARR_SMOOTH, "smooth");
        DevBuf smooth_tree = fetch_tree_order(*t, smr);
        DevBuf counts_tree(sizeof(int32_t) * (size_t)t->n), counts_file(sizeof(int32_t) * (size_t)t->n);
        PNB_CUDA(cudaMemsetAsync(counts_tree.p, 0, counts_tree.bytes, t->stream));
        run_gather_counts(*t, period, smooth_tree.p, counts_tree.as<int32_t>());
        permute_to_file(*t, counts_tree.p, counts_file.p, 1, 4, t->stream);
        PNB_CUDA(cudaMemcpyAsync(counts_out, counts_file.p, counts_file.bytes,
                                 mem == PNB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, t->stream));
        PNB_CUDA(cudaStreamSynchronize(t->stream));
    });
}

int pnb_ball_query(pnb_tree *h, double cx, double cy, double cz, double r, double period_in, int64_t *out,
                   int64_t cap, int64_t *n_found)
{
    PNB_NEED_TREE(h);
    Tree *t = reinterpret_cast<Tree *>(h);
    return guarded([&] {
        require_device();
        const double period = effective_period(*t, period_in, nullptr);
        int64_t f = run_ball_query(*t, cx, cy, cz, r, period, out, out ? cap : 0);
        if (n_found) *n_found = f;
    });
}

void pnb_release_cached_memory(void) { cache_trim(); }

double pnb_last_device_ms(int stage) { return (stage >= 0 && stage < 4) ? g_stage_ms[stage] : 0.0; }
int64_t pnb_total_launches(void) { return g_launches; }
double pnb_last_kernel_ms(int which) { return (which >= 0 && which < 4) ? g_kernel_ms[which] : 0.0; }
int64_t pnb_last_launch_count(int stage) { return (stage >= 0 && stage < 4) ? g_stage_launches[stage] : 0; }

}  // extern "C"
