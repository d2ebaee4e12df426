// route.cu -- multi-GPU particle routing: which rank owns a particle, and the packing of per-destination send
// buffers.  No reference counterpart (the reference is single-process); this is the device
// side of pynbody_b200/distributed.py's step 2 ("particles move to the rank that owns their domain").
//
//   pnb_route_plan   every particle walks the recorded coordinate bisection (one comparison per level, the split
//                    planes live in shared memory) -> owner rank (one byte) + particles per destination.
//   pnb_route_pack   stable multi-split: rows {x, y, z, m} are written destination-major -- rank 0's particles
//                    first, in their original order, then rank 1's ... -- together with the original index of
//                    every row, so that one all-to-all of contiguous ranges moves them and results can be
//                    returned to where they came from.  Three passes: per-block histogram, scan over blocks,
//                    scatter with ranks from warp match/ballot.  Deterministic (no atomics in the scatter).
//
// Both replace a chain of eager tensor ops (per-level gathers, a stable sort by owner, fancy-index gathers and
// concatenations) that cost ~9 of the 10.4 ms routing took at 8 GPUs x 16.7 M particles (profiles/r1_scaling.md).
#include "common.cuh"

namespace pnb {

namespace {
constexpr int RT_THREADS = 256;
constexpr int RT_PER_THREAD = 4;
constexpr int RT_BLOCK = RT_THREADS * RT_PER_THREAD;     // particles per block
constexpr int RT_MAX_SPLITS = 255;

struct Splits {
    int n;
    int32_t axis[RT_MAX_SPLITS];
    int32_t child[RT_MAX_SPLITS][2];                      // >= 0: another split; < 0: -(rank + 1)
    double plane[RT_MAX_SPLITS];
};

template <typename T>
__global__ void __launch_bounds__(RT_THREADS) k_route_owner(const unsigned char *__restrict__ pos, int64_t s0, int64_t s1,
                                                            int64_t n, const Splits *__restrict__ sp_g, int n_ranks,
                                                            uint8_t *__restrict__ owner, unsigned long long *__restrict__ counts)
{
    __shared__ Splits sp;
    __shared__ unsigned int hist[256];
    for (int i = threadIdx.x; i < (int)(sizeof(Splits) / 4); i += RT_THREADS)
        reinterpret_cast<uint32_t *>(&sp)[i] = reinterpret_cast<const uint32_t *>(sp_g)[i];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RT_BLOCK;
    for (int r = 0; r < RT_PER_THREAD; ++r) {
        const int64_t i = base + r * RT_THREADS + threadIdx.x;
        if (i >= n) break;
        const unsigned char *row = pos + i * s0;
        const T x[3] = {*reinterpret_cast<const T *>(row), *reinterpret_cast<const T *>(row + s1),
                        *reinterpret_cast<const T *>(row + 2 * s1)};
        int node = sp.n > 0 ? 0 : -1;
        while (node >= 0) {                               // (coordinate >= plane) goes to the upper side
            const T c = x[sp.axis[node]];
            node = sp.child[node][c >= (T)sp.plane[node] ? 1 : 0];
        }
        const int o = -node - 1;
        owner[i] = (uint8_t)o;
        atomicAdd(&hist[o], 1u);
    }
    __syncthreads();
    if (threadIdx.x < n_ranks && hist[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)hist[threadIdx.x]);
}

__global__ void __launch_bounds__(RT_THREADS) k_route_block_hist(const uint8_t *__restrict__ owner, int64_t n, int n_ranks,
                                                                 uint32_t *__restrict__ block_hist)
{
    __shared__ unsigned int hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RT_BLOCK;
    for (int r = 0; r < RT_PER_THREAD; ++r) {
        const int64_t i = base + r * RT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&hist[owner[i]], 1u);
    }
    __syncthreads();
    if (threadIdx.x < n_ranks) block_hist[(int64_t)blockIdx.x * n_ranks + threadIdx.x] = hist[threadIdx.x];
}

// exclusive scan, per destination, over the blocks (+ the destination's global start): one block, one warp per rank
__global__ void k_route_scan(uint32_t *__restrict__ block_hist, int64_t n_blocks, int n_ranks, uint64_t *__restrict__ block_base)
{
    __shared__ unsigned long long total[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    for (int r = warp; r < n_ranks; r += warps) {
        unsigned long long run = 0;
        for (int64_t b0 = 0; b0 < n_blocks; b0 += 32) {
            const int64_t b = b0 + lane;
            const unsigned long long v = b < n_blocks ? block_hist[b * n_ranks + r] : 0ull;
            unsigned long long inc = v;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long up = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += up;
            }
            if (b < n_blocks) block_base[b * n_ranks + r] = run + inc - v;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) total[r] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                               // destination starts
        unsigned long long acc = 0;
        for (int r = 0; r < n_ranks; ++r) { const unsigned long long t = total[r]; total[r] = acc; acc += t; }
    }
    __syncthreads();
    for (int64_t i = threadIdx.x; i < n_blocks * n_ranks; i += blockDim.x) block_base[i] += total[i % n_ranks];
}

template <typename T>
__global__ void __launch_bounds__(RT_THREADS) k_route_scatter(const unsigned char *__restrict__ pos, int64_t s0, int64_t s1,
                                                              const unsigned char *__restrict__ mass, int64_t ms, int64_t n,
                                                              const uint8_t *__restrict__ owner, int n_ranks,
                                                              const uint64_t *__restrict__ block_base, T *__restrict__ packed,
                                                              int64_t *__restrict__ src_index)
{
    __shared__ unsigned int wcount[RT_THREADS / 32][256];
    __shared__ unsigned long long cursor[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < n_ranks) cursor[threadIdx.x] = block_base[(int64_t)blockIdx.x * n_ranks + threadIdx.x];
    const int64_t base = (int64_t)blockIdx.x * RT_BLOCK;
    for (int r = 0; r < RT_PER_THREAD; ++r) {
        for (int i = threadIdx.x; i < (RT_THREADS / 32) * 256; i += RT_THREADS) (&wcount[0][0])[i] = 0;
        __syncthreads();
        const int64_t i = base + r * RT_THREADS + threadIdx.x;
        const bool valid = i < n;
        const int o = valid ? owner[i] : 0;
        const unsigned active = __ballot_sync(0xffffffffu, valid);
        unsigned peers = 0;
        if (valid) {
            peers = __match_any_sync(active, o);
            if (lane == __ffs(peers) - 1) wcount[warp][o] = __popc(peers);
        }
        __syncthreads();
        if (valid) {
            unsigned long long dst = cursor[o] + __popc(peers & ((1u << lane) - 1u));
            for (int w = 0; w < warp; ++w) dst += wcount[w][o];
            const unsigned char *row = pos + i * s0;
            T *out = packed + dst * 4;
            out[0] = *reinterpret_cast<const T *>(row);
            out[1] = *reinterpret_cast<const T *>(row + s1);
            out[2] = *reinterpret_cast<const T *>(row + 2 * s1);
            out[3] = mass ? *reinterpret_cast<const T *>(mass + i * ms) : (T)0;
            src_index[dst] = i;
        }
        __syncthreads();
        if (threadIdx.x < n_ranks) {
            unsigned int tot = 0;
            for (int w = 0; w < RT_THREADS / 32; ++w) tot += wcount[w][threadIdx.x];
            cursor[threadIdx.x] += tot;
        }
        __syncthreads();
    }
}

void check_route_args(const void *pos, int64_t n, int dtype, int n_ranks)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw Error(PNB_ERR_CUDA, "no CUDA device available: pnb_b200 has no CPU fallback");
    }
    if (dtype != PNB_F32 && dtype != PNB_F64) throw Error(PNB_ERR_VALUE, "Unsupported array dtype");
    if (n < 0 || (n > 0 && !pos)) throw Error(PNB_ERR_VALUE, "bad particle array");
    if (n_ranks < 1 || n_ranks > 256) throw Error(PNB_ERR_VALUE, "between 1 and 256 ranks are supported");
}

template <typename F>
int guarded_route(F &&f)
{
    try {
        ApiScope scope;
        f();
        return PNB_OK;
    } catch (const Error &e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return PNB_ERR_RUNTIME;
    }
}
}  // namespace
}  // namespace pnb

using namespace pnb;

extern "C" {

int pnb_route_plan(const void *pos, int64_t stride0, int64_t stride1, int64_t n, int dtype, const int32_t *split_axis,
                   const double *split_plane, const int32_t *split_child, int n_splits, int n_ranks, uint8_t *owner_out,
                   int64_t *counts_out)
{
    return guarded_route([&] {
        check_route_args(pos, n, dtype, n_ranks);
        if (n_splits < 0 || n_splits > RT_MAX_SPLITS) throw Error(PNB_ERR_VALUE, "too many split planes");
        if (!owner_out || !counts_out) throw Error(PNB_ERR_VALUE, "output pointer is null");
        cudaStream_t s = lib_stream();
        Splits h;
        memset(&h, 0, sizeof(h));
        h.n = n_splits;
        for (int i = 0; i < n_splits; ++i) {
            h.axis[i] = split_axis[i];
            h.plane[i] = split_plane[i];
            h.child[i][0] = split_child[2 * i];
            h.child[i][1] = split_child[2 * i + 1];
            if (h.axis[i] < 0 || h.axis[i] > 2) throw Error(PNB_ERR_VALUE, "split axis out of range");
            for (int c = 0; c < 2; ++c)
                if (h.child[i][c] >= n_splits || -h.child[i][c] - 1 >= n_ranks) throw Error(PNB_ERR_VALUE, "bad split tree");
        }
        DevBuf sp(sizeof(Splits)), counts(sizeof(unsigned long long) * 256);
        PNB_CUDA(cudaMemcpyAsync(sp.p, &h, sizeof(h), cudaMemcpyHostToDevice, s));
        PNB_CUDA(cudaMemsetAsync(counts.p, 0, counts.bytes, s));
        if (n > 0) {
            const unsigned grid = (unsigned)((n + RT_BLOCK - 1) / RT_BLOCK);
            if (dtype == PNB_F64)
                PNB_LAUNCH((k_route_owner<double>), grid, RT_THREADS, 0, s, (const unsigned char *)pos, stride0, stride1, n,
                           sp.as<Splits>(), n_ranks, owner_out, counts.as<unsigned long long>());
            else
                PNB_LAUNCH((k_route_owner<float>), grid, RT_THREADS, 0, s, (const unsigned char *)pos, stride0, stride1, n,
                           sp.as<Splits>(), n_ranks, owner_out, counts.as<unsigned long long>());
        }
        unsigned long long hc[256];
        PNB_CUDA(cudaMemcpyAsync(hc, counts.p, sizeof(unsigned long long) * n_ranks, cudaMemcpyDeviceToHost, s));
        PNB_CUDA(cudaStreamSynchronize(s));
        for (int r = 0; r < n_ranks; ++r) counts_out[r] = (int64_t)hc[r];
    });
}

int pnb_route_pack(const void *pos, int64_t stride0, int64_t stride1, const void *mass, int64_t mass_stride, int64_t n,
                   int dtype, const uint8_t *owner, int n_ranks, void *packed_out, int64_t *src_index_out)
{
    return guarded_route([&] {
        check_route_args(pos, n, dtype, n_ranks);
        if (!owner || !packed_out || !src_index_out) throw Error(PNB_ERR_VALUE, "pointer is null");
        if (n == 0) return;
        cudaStream_t s = lib_stream();
        const int64_t nb = (n + RT_BLOCK - 1) / RT_BLOCK;
        DevBuf hist(sizeof(uint32_t) * (size_t)nb * n_ranks), base(sizeof(uint64_t) * (size_t)nb * n_ranks);
        PNB_LAUNCH(k_route_block_hist, (unsigned)nb, RT_THREADS, 0, s, owner, n, n_ranks, hist.as<uint32_t>());
        PNB_LAUNCH(k_route_scan, 1, 1024, 0, s, hist.as<uint32_t>(), nb, n_ranks, base.as<uint64_t>());
        if (dtype == PNB_F64)
            PNB_LAUNCH((k_route_scatter<double>), (unsigned)nb, RT_THREADS, 0, s, (const unsigned char *)pos, stride0, stride1,
                       (const unsigned char *)mass, mass_stride, n, owner, n_ranks, base.as<uint64_t>(), (double *)packed_out,
                       src_index_out);
        else
            PNB_LAUNCH((k_route_scatter<float>), (unsigned)nb, RT_THREADS, 0, s, (const unsigned char *)pos, stride0, stride1,
                       (const unsigned char *)mass, mass_stride, n, owner, n_ranks, base.as<uint64_t>(), (float *)packed_out,
                       src_index_out);
        PNB_CUDA(cudaStreamSynchronize(s));
    });
}

}  // extern "C"
