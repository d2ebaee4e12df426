// io.cu -- movement of the borrowed, possibly strided, host or device arrays of the reference ABI
// (numpy buffers accessed through PyArray_GETPTR in kd.h:163-190) into dense device columns and back,
// and the tree-order <-> file-order permutations (particleOffsets, kd.h:205-206).
#include "common.cuh"

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>

namespace pnb {

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// ---- host <-> device copies of ordinary (pageable) host memory ----------------------------------------
// numpy arrays are pageable.  A plain cudaMemcpy from pageable memory is staged by the driver through
// one thread (~13 GB/s measured here); instead the copy is chunked through two pinned buffers, each chunk
// moved host-side by several threads while the previous chunk's DMA is in flight.  Memory that is already
// pinned (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) goes straight to cudaMemcpyAsync.
namespace {
constexpr size_t STAGE_BYTES = (size_t)32 << 20;
struct Staging {
    void *buf[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool ok = false;
    Staging()
    {
        ok = cudaHostAlloc(&buf[0], STAGE_BYTES, cudaHostAllocDefault) == cudaSuccess
             && cudaHostAlloc(&buf[1], STAGE_BYTES, cudaHostAllocDefault) == cudaSuccess
             && cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming) == cudaSuccess
             && cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
};
// one set of pinned buffers and events per (thread, device): events are bound to the device that was current
// when they were created (a thread that calls pnb_set_device() in between must not reuse another device's)
Staging &staging()
{
    static thread_local Staging *per_dev[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    Staging *&s = per_dev[dev & 63];
    if (!s) s = new Staging();
    return *s;
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

}  // namespace
bool host_is_pinned(const void *p) { return is_pinned(p); }
namespace {
// A small pool of copy threads that lives as long as the library: spawning threads per 32 MB chunk cost more
// than the copies they made.  One job at a time (callers hold the job mutex); the calling thread copies a share too.
class CopyPool {
public:
    static CopyPool &get() { static CopyPool p; return p; }
    void copy(void *dst, const void *src, size_t bytes)
    {
        const unsigned nt = (unsigned)std::min<size_t>(workers_.size() + 1, std::max<size_t>(1, bytes >> 21));
        if (nt <= 1) { memcpy(dst, src, bytes); return; }
        std::lock_guard<std::mutex> job(job_mu_);
        const size_t per = ((bytes + nt - 1) / nt + 63) & ~(size_t)63;
        {
            std::lock_guard<std::mutex> l(mu_);
            dst_ = (char *)dst; src_ = (const char *)src; bytes_ = bytes; per_ = per; parts_ = nt - 1; next_ = 0; left_ = nt - 1;
            ++generation_;
        }
        cv_.notify_all();
        const size_t a = (size_t)(nt - 1) * per;                     // the caller takes the last share
        if (a < bytes) memcpy((char *)dst + a, (const char *)src + a, bytes - a);
        std::unique_lock<std::mutex> l(mu_);
        done_cv_.wait(l, [&] { return left_ == 0; });
    }

private:
    CopyPool()
    {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        unsigned want = 12;                                           // PNB_COPY_THREADS overrides (8: +2 %, 16: no gain, measured)
        if (const char *e = getenv("PNB_COPY_THREADS")) want = (unsigned)std::max(1, atoi(e));
        const unsigned n = std::min(want, hw) - 1;
        for (unsigned i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> l(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void run()
    {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> l(mu_);
        for (;;) {
            cv_.wait(l, [&] { return stop_ || (generation_ != seen && next_ < parts_); });
            if (stop_) return;
            while (next_ < parts_) {
                const unsigned part = next_++;
                const size_t a = (size_t)part * per_;
                const size_t len = a < bytes_ ? std::min(per_, bytes_ - a) : 0;
                char *d = dst_ + a;
                const char *s = src_ + a;
                l.unlock();
                if (len) memcpy(d, s, len);
                l.lock();
                if (--left_ == 0) done_cv_.notify_all();
            }
            seen = generation_;
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, job_mu_;
    std::condition_variable cv_, done_cv_;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0, per_ = 0;
    unsigned parts_ = 0, next_ = 0, left_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

void parallel_memcpy(void *dst, const void *src, size_t bytes) { CopyPool::get().copy(dst, src, bytes); }
}  // namespace

void host_to_device(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream)
{
    Staging &st = staging();
    if (bytes < ((size_t)4 << 20) || !st.ok || is_pinned(src_host)) {
        PNB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream));
        PNB_CUDA(cudaStreamSynchronize(stream));
        return;
    }
    int c = 0;
    for (size_t off = 0; off < bytes; off += STAGE_BYTES, ++c) {
        const size_t len = std::min(STAGE_BYTES, bytes - off);
        const int b = c & 1;
        if (c >= 2) PNB_CUDA(cudaEventSynchronize(st.done[b]));        // this buffer's previous DMA has finished
        parallel_memcpy(st.buf[b], (const char *)src_host + off, len);
        PNB_CUDA(cudaMemcpyAsync((char *)dst_dev + off, st.buf[b], len, cudaMemcpyHostToDevice, stream));
        PNB_CUDA(cudaEventRecord(st.done[b], stream));
    }
    PNB_CUDA(cudaStreamSynchronize(stream));
}

void device_to_host(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t stream)
{
    Staging &st = staging();
    if (bytes < ((size_t)4 << 20) || !st.ok || is_pinned(dst_host)) {
        PNB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
        PNB_CUDA(cudaStreamSynchronize(stream));
        return;
    }
    const size_t nchunks = (bytes + STAGE_BYTES - 1) / STAGE_BYTES;
    auto issue = [&](size_t c) {
        const size_t off = c * STAGE_BYTES, len = std::min(STAGE_BYTES, bytes - off);
        PNB_CUDA(cudaMemcpyAsync(st.buf[c & 1], (const char *)src_dev + off, len, cudaMemcpyDeviceToHost, stream));
        PNB_CUDA(cudaEventRecord(st.done[c & 1], stream));
    };
    issue(0);
    for (size_t c = 0; c < nchunks; ++c) {
        PNB_CUDA(cudaEventSynchronize(st.done[c & 1]));
        if (c + 1 < nchunks) issue(c + 1);                              // next DMA overlaps this chunk's host copy
        const size_t off = c * STAGE_BYTES, len = std::min(STAGE_BYTES, bytes - off);
        parallel_memcpy((char *)dst_host + off, st.buf[c & 1], len);
    }
}

template <typename E>
__global__ void k_strided_to_columns(const unsigned char *__restrict__ src, int64_t s0, int64_t s1, int ncols,
                                     int64_t n, E *__restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = 0; c < ncols; ++c)
        out[(int64_t)c * n + i] = *reinterpret_cast<const E *>(src + i * s0 + c * s1);
}
template <typename E>
__global__ void k_columns_to_strided(const E *__restrict__ in, int64_t s0, int64_t s1, int ncols, int64_t n,
                                     unsigned char *__restrict__ dst)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = 0; c < ncols; ++c)
        *reinterpret_cast<E *>(dst + i * s0 + c * s1) = in[(int64_t)c * n + i];
}
template <typename E>
__global__ void k_to_tree(const uint32_t *__restrict__ order, const E *__restrict__ in, E *__restrict__ out,
                          int ncols, int64_t n)
{
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t id = order[p];
    for (int c = 0; c < ncols; ++c) out[(int64_t)c * n + p] = in[(int64_t)c * n + id];
}
template <typename E>
__global__ void k_to_file(const uint32_t *__restrict__ order, const E *__restrict__ in, E *__restrict__ out,
                          int ncols, int64_t n)
{
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t id = order[p];
    for (int c = 0; c < ncols; ++c) out[(int64_t)c * n + id] = in[(int64_t)c * n + p];
}

static bool dense_rows(int64_t s0, int64_t s1, int ncols, int64_t elem)
{
    return (ncols == 1 || s1 == elem) && s0 == (int64_t)ncols * elem;
}

void fetch_columns(const void *ptr, int64_t s0, int64_t s1, int ncols, int dtype, int mem, int64_t n, void *dev_out,
                   cudaStream_t stream)
{
    const int64_t elem = (int64_t)tsize(dtype);
    if (s0 % elem || (ncols > 1 && s1 % elem) || ((uintptr_t)ptr % elem))
        throw Error(PNB_ERR_VALUE, "array is not aligned to its element size");
    const int BS = 256;
    if (mem == PNB_DEVICE) {
        if (elem == 8) PNB_LAUNCH((k_strided_to_columns<uint64_t>), nblk(n, BS), BS, 0, stream, (const unsigned char *)ptr, s0, s1, ncols, n, (uint64_t *)dev_out);
        else PNB_LAUNCH((k_strided_to_columns<uint32_t>), nblk(n, BS), BS, 0, stream, (const unsigned char *)ptr, s0, s1, ncols, n, (uint32_t *)dev_out);
        return;
    }
    if (ncols == 1 && s0 == elem) {                                    // dense 1-D: straight copy
        host_to_device(dev_out, ptr, (size_t)(n * elem), stream);
        return;
    }
    if (ncols == 1 && s0 > elem && s0 <= 8 * elem && n > 4096) {
        // a column view of a wider host array (e.g. pos[:, 1], stride 3 elements): picking the elements out on the
        // host is a scalar loop at well under 1 GB/s; ship the enclosing span and pick on the device instead
        const size_t span = (size_t)((n - 1) * s0 + elem);
        DevBuf wide(span);
        host_to_device(wide.p, ptr, span, stream);
        if (elem == 8) PNB_LAUNCH((k_strided_to_columns<uint64_t>), nblk(n, BS), BS, 0, stream, wide.as<unsigned char>(), s0, 0, 1, n, (uint64_t *)dev_out);
        else PNB_LAUNCH((k_strided_to_columns<uint32_t>), nblk(n, BS), BS, 0, stream, wide.as<unsigned char>(), s0, 0, 1, n, (uint32_t *)dev_out);
        PNB_CUDA(cudaStreamSynchronize(stream));
        return;
    }
    DevBuf stage((size_t)(n * ncols * elem));
    if (dense_rows(s0, s1, ncols, elem)) {
        host_to_device(stage.p, ptr, stage.bytes, stream);
    } else {                                                           // arbitrary strides: pack on the host
        std::vector<unsigned char> packed((size_t)(n * ncols * elem));
        const unsigned char *src = (const unsigned char *)ptr;
        for (int64_t i = 0; i < n; ++i)
            for (int c = 0; c < ncols; ++c)
                memcpy(&packed[(size_t)((i * ncols + c) * elem)], src + i * s0 + c * s1, (size_t)elem);
        PNB_CUDA(cudaMemcpyAsync(stage.p, packed.data(), stage.bytes, cudaMemcpyHostToDevice, stream));
        PNB_CUDA(cudaStreamSynchronize(stream));
    }
    if (elem == 8) PNB_LAUNCH((k_strided_to_columns<uint64_t>), nblk(n, BS), BS, 0, stream, stage.as<unsigned char>(), ncols * elem, elem, ncols, n, (uint64_t *)dev_out);
    else PNB_LAUNCH((k_strided_to_columns<uint32_t>), nblk(n, BS), BS, 0, stream, stage.as<unsigned char>(), ncols * elem, elem, ncols, n, (uint32_t *)dev_out);
    PNB_CUDA(cudaStreamSynchronize(stream));
}

void store_columns(void *ptr, int64_t s0, int64_t s1, int ncols, int dtype, int mem, int64_t n, const void *dev_in,
                   cudaStream_t stream)
{
    const int64_t elem = (int64_t)tsize(dtype);
    if (s0 % elem || (ncols > 1 && s1 % elem) || ((uintptr_t)ptr % elem))
        throw Error(PNB_ERR_VALUE, "array is not aligned to its element size");
    const int BS = 256;
    if (mem == PNB_DEVICE) {
        if (elem == 8) PNB_LAUNCH((k_columns_to_strided<uint64_t>), nblk(n, BS), BS, 0, stream, (const uint64_t *)dev_in, s0, s1, ncols, n, (unsigned char *)ptr);
        else PNB_LAUNCH((k_columns_to_strided<uint32_t>), nblk(n, BS), BS, 0, stream, (const uint32_t *)dev_in, s0, s1, ncols, n, (unsigned char *)ptr);
        return;
    }
    if (ncols == 1 && s0 == elem) {
        PNB_CUDA(cudaStreamSynchronize(stream));
        device_to_host(ptr, dev_in, (size_t)(n * elem), stream);
        return;
    }
    DevBuf stage((size_t)(n * ncols * elem));
    if (elem == 8) PNB_LAUNCH((k_columns_to_strided<uint64_t>), nblk(n, BS), BS, 0, stream, (const uint64_t *)dev_in, ncols * elem, elem, ncols, n, stage.as<unsigned char>());
    else PNB_LAUNCH((k_columns_to_strided<uint32_t>), nblk(n, BS), BS, 0, stream, (const uint32_t *)dev_in, ncols * elem, elem, ncols, n, stage.as<unsigned char>());
    if (dense_rows(s0, s1, ncols, elem)) {
        PNB_CUDA(cudaStreamSynchronize(stream));
        device_to_host(ptr, stage.p, stage.bytes, stream);
    } else {
        std::vector<unsigned char> packed(stage.bytes);
        PNB_CUDA(cudaMemcpyAsync(packed.data(), stage.p, stage.bytes, cudaMemcpyDeviceToHost, stream));
        PNB_CUDA(cudaStreamSynchronize(stream));
        unsigned char *dst = (unsigned char *)ptr;
        for (int64_t i = 0; i < n; ++i)
            for (int c = 0; c < ncols; ++c)
                memcpy(dst + i * s0 + c * s1, &packed[(size_t)((i * ncols + c) * elem)], (size_t)elem);
    }
}

void permute_to_tree(const Tree &t, const void *file_in, void *tree_out, int ncols, int elem, cudaStream_t s)
{
    const int BS = 256;
    if (elem == 8) PNB_LAUNCH((k_to_tree<uint64_t>), nblk(t.n, BS), BS, 0, s, t.order.as<uint32_t>(), (const uint64_t *)file_in, (uint64_t *)tree_out, ncols, t.n);
    else PNB_LAUNCH((k_to_tree<uint32_t>), nblk(t.n, BS), BS, 0, s, t.order.as<uint32_t>(), (const uint32_t *)file_in, (uint32_t *)tree_out, ncols, t.n);
}
void permute_to_file(const Tree &t, const void *tree_in, void *file_out, int ncols, int elem, cudaStream_t s)
{
    const int BS = 256;
    if (elem == 8) PNB_LAUNCH((k_to_file<uint64_t>), nblk(t.n, BS), BS, 0, s, t.order.as<uint32_t>(), (const uint64_t *)tree_in, (uint64_t *)file_out, ncols, t.n);
    else PNB_LAUNCH((k_to_file<uint32_t>), nblk(t.n, BS), BS, 0, s, t.order.as<uint32_t>(), (const uint32_t *)tree_in, (uint32_t *)file_out, ncols, t.n);
}

}  // namespace pnb
