// select.cu -- tree-less geometric selections (sphere / cuboid, optionally periodic).
//
// Replaces pynbody.filt.geometry_selection.selection (pynbody/filt/geometry_selection.pyx:31-103,
// geometry_selection.hpp:3-119): one pass over a C-contiguous N x 3 position array writing one uint8
// flag per particle.  The arithmetic is the reference's, in the array's own dtype T and without FMA
// contraction (the library is built with --fmad=false), so the flags are bit-identical:
//   d = p - centre; with wrap > 0: d -= wrap if d > wrap/2, d += wrap if d < -wrap/2   (hpp:11-28)
//   sphere:  (dx*dx + dy*dy) + dz*dz < r_max*r_max                                      (hpp:47-62)
//   cuboid:  |dx| < xsize && |dy| < ysize && |dz| < zsize, centre/half-size in T        (hpp:64-77)
//
// HBM-bound: 3*sizeof(T) bytes read + 1 byte written per particle.  A CTA stages 12 KB of contiguous values
// through shared memory with 16-byte loads, then each thread tests four (f32) or two (f64) particles.
#include "common.cuh"

namespace pnb {
void host_to_device(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream);
void device_to_host(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t stream);

namespace {

constexpr int SEL_THREADS = 256;
// particles per thread and per CTA iteration: 12 KB of shared memory for either dtype
template <typename T> constexpr int sel_per_thread() { return 16 / (int)sizeof(T); }

template <typename T>
struct SelParams {
    T cx, cy, cz;       // centre
    T a, b, c;          // sphere: a = r_max^2; cuboid: half sizes
    T wrap, half_wrap;
};

// One CTA iteration: 3*SEL_TILE contiguous values are staged through shared memory with 16-byte loads
// (scalar loads for the ragged last tile or an unaligned base), then thread t tests particles
// t, t+256, ... (shared-memory stride of 3 words: conflict-free) and writes their flags (coalesced bytes).
template <typename T, bool CUBE, bool WRAP>
__global__ void __launch_bounds__(SEL_THREADS) k_select(const T *__restrict__ pos, int64_t n, SelParams<T> p,
                                                         uint8_t *__restrict__ out, int aligned16)
{
    constexpr int SEL_PER_THREAD = sel_per_thread<T>(), SEL_TILE = SEL_THREADS * SEL_PER_THREAD;
    __shared__ __align__(16) T tile[3 * SEL_TILE];
    constexpr int VEC = 16 / (int)sizeof(T);
    const int64_t ntiles = (n + SEL_TILE - 1) / SEL_TILE;
    for (int64_t blk = blockIdx.x; blk < ntiles; blk += gridDim.x) {
        const int64_t first = blk * SEL_TILE;
        const int count = (int)min((int64_t)SEL_TILE, n - first);
        const T *src = pos + 3 * first;
        if (count == SEL_TILE && aligned16) {
            const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
            uint4 *dst4 = reinterpret_cast<uint4 *>(tile);
#pragma unroll
            for (int j = threadIdx.x; j < 3 * SEL_TILE / VEC; j += SEL_THREADS) dst4[j] = __ldg(src4 + j);
        } else {
            for (int j = threadIdx.x; j < 3 * count; j += SEL_THREADS) tile[j] = src[j];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SEL_PER_THREAD; ++r) {
            const int i = r * SEL_THREADS + threadIdx.x;
            if (i < count) {
                T dx = tile[3 * i] - p.cx, dy = tile[3 * i + 1] - p.cy, dz = tile[3 * i + 2] - p.cz;
                if (WRAP) {
                    if (dx > p.half_wrap) dx -= p.wrap;
                    if (dy > p.half_wrap) dy -= p.wrap;
                    if (dz > p.half_wrap) dz -= p.wrap;
                    if (dx < -p.half_wrap) dx += p.wrap;
                    if (dy < -p.half_wrap) dy += p.wrap;
                    if (dz < -p.half_wrap) dz += p.wrap;
                }
                bool in;
                if (CUBE) in = fabs(dx) < p.a && fabs(dy) < p.b && fabs(dz) < p.c;
                else in = (dx * dx + dy * dy) + dz * dz < p.a;
                out[first + i] = in ? 1 : 0;
            }
        }
        __syncthreads();
    }
}

template <typename T>
void select_typed(const void *pos_dev, int64_t n, int region, const double *prm, double wrap_d, uint8_t *out_dev,
                  cudaStream_t s)
{
    SelParams<T> p;
    const T wrap = (T)wrap_d;
    p.wrap = wrap;
    p.half_wrap = wrap / 2;
    if (region == PNB_REGION_SPHERE) {
        p.cx = (T)prm[0]; p.cy = (T)prm[1]; p.cz = (T)prm[2];
        const T r = (T)prm[3];
        p.a = r * r; p.b = p.c = 0;
    } else {
        const T x0 = (T)prm[0], y0 = (T)prm[1], z0 = (T)prm[2], x1 = (T)prm[3], y1 = (T)prm[4], z1 = (T)prm[5];
        p.cx = (x0 + x1) / 2; p.cy = (y0 + y1) / 2; p.cz = (z0 + z1) / 2;
        p.a = (x1 - x0) / 2; p.b = (y1 - y0) / 2; p.c = (z1 - z0) / 2;
    }
    constexpr int SEL_TILE = SEL_THREADS * sel_per_thread<T>();
    const int64_t ntiles = (n + SEL_TILE - 1) / SEL_TILE;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)148 * 8);    // 8 resident CTAs per SM, grid-stride
    const int al = (reinterpret_cast<uintptr_t>(pos_dev) & 15) == 0;
    const bool w = wrap > 0;
    const T *pp = static_cast<const T *>(pos_dev);
    if (region == PNB_REGION_SPHERE) {
        if (w) PNB_LAUNCH((k_select<T, false, true>), grid, SEL_THREADS, 0, s, pp, n, p, out_dev, al);
        else PNB_LAUNCH((k_select<T, false, false>), grid, SEL_THREADS, 0, s, pp, n, p, out_dev, al);
    } else {
        if (w) PNB_LAUNCH((k_select<T, true, true>), grid, SEL_THREADS, 0, s, pp, n, p, out_dev, al);
        else PNB_LAUNCH((k_select<T, true, false>), grid, SEL_THREADS, 0, s, pp, n, p, out_dev, al);
    }
}

void select_region(const void *pos, int64_t n, int dtype, int mem, int region, const double *prm, int n_prm,
                   double wrap, uint8_t *out, int out_mem)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw Error(PNB_ERR_CUDA, "no CUDA device available: pnb_b200 has no CPU fallback");
    }
    if (dtype != PNB_F32 && dtype != PNB_F64) throw Error(PNB_ERR_TYPE, "positions must be float32 or float64");
    if (region == PNB_REGION_SPHERE) {
        if (n_prm != 4) throw Error(PNB_ERR_VALUE, "Sphere selection requires 4 parameters: (x0, y0, z0, r_max)");
    } else if (region == PNB_REGION_CUBE) {
        if (n_prm != 6) throw Error(PNB_ERR_VALUE, "Cube selection requires 6 parameters: (x0, y0, z0, x1, y1, z1)");
    } else {
        throw Error(PNB_ERR_VALUE, "Unknown region type");
    }
    if (n < 0) throw Error(PNB_ERR_VALUE, "negative particle count");
    if (n == 0) return;
    if (!pos || !prm || !out) throw Error(PNB_ERR_VALUE, "null pointer");
    cudaStream_t s = lib_stream();
    StageTimer stage(3, s);
    const size_t bytes = (size_t)n * 3 * tsize(dtype);
    DevBuf staged, flags;
    const void *pos_dev = pos;
    if (mem != PNB_DEVICE) {
        staged.alloc(bytes);
        host_to_device(staged.p, pos, bytes, s);
        pos_dev = staged.p;
    }
    uint8_t *out_dev = out;
    if (out_mem != PNB_DEVICE) {
        flags.alloc((size_t)n);
        out_dev = flags.as<uint8_t>();
    }
    {
        KernelTimer timer(3, s);
        if (dtype == PNB_F64) select_typed<double>(pos_dev, n, region, prm, wrap, out_dev, s);
        else select_typed<float>(pos_dev, n, region, prm, wrap, out_dev, s);
        timer.stop();
    }
    if (out_mem != PNB_DEVICE) device_to_host(out, out_dev, (size_t)n, s);
    PNB_CUDA(cudaStreamSynchronize(s));
    stage.stop();
}

}  // namespace
}  // namespace pnb

using namespace pnb;

extern "C" int pnb_select_region(const void *pos, int64_t n, int dtype, int mem, int region, const double *params,
                                 int n_params, double wrap, uint8_t *out, int out_mem)
{
    try {
        ApiScope scope;
        select_region(pos, n, dtype, mem, region, params, n_params, wrap, out, out_mem);
        return PNB_OK;
    } catch (const Error &e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return PNB_ERR_RUNTIME;
    }
}
