// render.cu -- K6/K7: SPH image and 3-D grid rendering.
//
// Replaces _render.render_image / _render.to_3d_grid (pynbody/sph/_render.pyx:209-365, 373-497) and
// their kernel lookup get_kernel (:186-193).  The reference is a per-particle SCATTER: every particle
// loops over the pixels its 2h footprint covers and does `image[y,x] += qty_i * LUT(d2) / h^dim`.
//
// On the GPU the same sums are evaluated as a tile-binned GATHER:
//   1. footprint pass: per particle (and per periodic image) the cull tests and the inclusive-exclusive
//      pixel rectangle [xa,xb) x [ya,yb) (x [za,zb)) are computed with the reference's double
//      arithmetic and C integer truncation (_render.pyx:297-346, 436-480); the number of 32x16-pixel
//      (4x8x8-voxel) tiles it overlaps is counted;
//   2. a prefix sum and a second footprint pass emit (tile, particle) pairs, which one radix sort
//      groups by tile;
//   3. a record pass writes one 64-byte record per pair (float position relative to the tile origin,
//      1/kernel support, weight, the rectangle clipped to the tile -- for images as a COLUMN MASK (one bit per
//      lane) and a ROW MASK (one bit per tile row), so that the render kernel's tests are single bit operations;
//      the doubles of the exact path) in sorted order, so that a tile's particles are contiguous;
//   4. the render kernel gives every CTA a fixed-length slice of the sorted pair list (perfect load
//      balance for clustered data), streams the records through shared memory with TMA bulk copies
//      (cp.async.bulk + mbarrier, double buffered), and its 256 threads accumulate the pixels of the current
//      tile in registers: image tiles are 32 x 16 pixels, a warp is one 32-pixel row segment and owns two
//      adjacent rows, so a record's row test is warp-uniform; grid tiles are 4 x 8 x 8 voxels, one per thread.
//      Pixels are flushed with one double-precision atomicAdd per (pixel, slice).
//      (C3 image: 54.1 ms in round 1, 35.6 ms now -- the table read through an opaque shared-window address, the exact
//      path's double arithmetic pinned inside its branch, bit-mask rectangle tests, edge rows by zero weight:
//      profiles/r2b_render_summary.txt.)
//      Three re-designs of the image kernel were built, verified against the goldens and measured SLOWER than
//      the round-1 kernel (54 ms then; profiles/r2_render_variants.md): 8 fixed pixels per thread on 64 x 32
//      tiles (84 ms: same 58 G warp instructions, issue efficiency 87 -> 60 %); one record per warp adding into
//      a shared-memory tile with float atomics (71 ms: those are CAS loops, ATOMS.CAST.SPIN); warps owning
//      row bands of a shared-memory tile (73 ms: 46 G instructions but the warps of a CTA wait for the busiest
//      band at every chunk barrier).
//
// LUT index parity: the reference indexes its 201-entry table with (unsigned)(ns * (d2 / kmax2))
// evaluated in double.  The inner loop evaluates the same quantity in float from tile-relative
// coordinates (absolute error < 9e-4 of an index step) and falls back to the reference's exact double
// expression whenever the float value lies within 2e-3 of an integer, so the selected table entry is
// the reference's in every case.  (A per-record bound -- 2.5e-4 for a typical 10-pixel kernel -- sends five
// times fewer evaluations through the exact path but costs a square root per record and thread: 57.6 vs 54.1 ms.)
//
// Perspective images (z_camera != 0) rescale the pixel grid per particle (_render.pyx:281-295); they go
// through an exact per-particle scatter kernel with double atomics instead of the tile path.
#include "common.cuh"

#include <algorithm>
#include <climits>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace pnb {

namespace {

constexpr int TILE_X = 32, TILE_Y = 16;          // image tile: 32 x 16 pixels, two pixels (rows 2w and 2w + 1 of warp w) per thread
constexpr int GTX = 4, GTY = 8, GTZ = 8;        // grid tile: 4 x 8 x 8 voxels (z fastest in memory)
constexpr int SEG = 1024;                       // (tile, particle) pairs per CTA
constexpr int CHUNK = 64;                       // records per TMA transfer
constexpr float BORDER = 2e-3f;
constexpr int64_t MAX_PAIRS = (int64_t)96 << 20;  // pairs handled per sort/render batch (6 GiB of records)

struct Geom {
    int nx, ny, nz, grid;
    double x1, x2, y1, y2, z1, z2;
    double pdx, pdy, pdz, xs, ys, zs;
    double z_camera, z0, smooth_lo, smooth_hi, z_lo, z_hi, min_smooth, max_d;
    float per_z_dx, per_z_dy, mid_x, mid_y;
    int ns, h_power, use_z;
    int tiles_x, tiles_y, tiles_z;
};

struct Cols {
    const void *x, *y, *z;
    int pos_f64;
    const double *h, *w;
    int64_t n;
};

struct Footprint {
    double xi, yi, zi, h, w;
    int xa, xb, ya, yb, za, zb;
};

// the periodic images handled by one pass of the tile pipeline (_render.pyx:237-243, :418): a pair's payload keeps
// the image number in its top 5 bits, the particle (relative to the range start) in the lower 27
constexpr int MAX_IMAGES = 32;
constexpr int IMAGE_SHIFT = 27;
struct Images {
    int n;
    float ox[MAX_IMAGES], oy[MAX_IMAGES], oz[MAX_IMAGES];
};

struct alignas(16) Rec {
    double xi, yi, zd, kmax2;     // exact path only.  zd: image -> dz = (z - z0) * use_z ; grid -> z
    float xr, yr, zr, inv;        // tile-relative position; image: zr = dz^2; inv = ns / kmax2 (<0: exact path only)
    float wq;                     // weight / h^dim
    uint32_t box;                 // grid: lxa | wx << 8 | lya << 16 | wy << 24 (rectangle clipped to the tile); image: column mask
    uint32_t boxz;                // grid: lza | wz << 8;  image: row mask (16 rows)
    uint32_t tile;
};
static_assert(sizeof(Rec) == 64, "record must be 64 bytes");

__device__ __forceinline__ int c_int(double v)
{
    // C's (int) on x86-64 (cvttsd2si): out-of-range and NaN give INT_MIN
    return (v >= 2147483648.0 || v < -2147483648.0 || v != v) ? INT_MIN : (int)v;
}
__device__ __forceinline__ double ld_pos(const void *p, int f64, int64_t i, float off)
{
    // x[i] + (float)offset evaluated in the element type of x (_render.pyx:270-272)
    if (f64) return ((const double *)p)[i] + (double)off;
    return (double)(((const float *)p)[i] + off);
}

// cull tests and pixel rectangle of particle i for one periodic image; false if it contributes nothing
__device__ __forceinline__ bool footprint(const Geom &g, const Cols &c, int64_t i, float ox, float oy, float oz,
                                          Footprint &f)
{
    const double xi = ld_pos(c.x, c.pos_f64, i, ox), yi = ld_pos(c.y, c.pos_f64, i, oy);
    const double zi = ld_pos(c.z, c.pos_f64, i, g.grid ? oz : 0.0f);
    double h = c.h[i];
    const double w = c.w[i];
    if (!g.grid) {
        if (zi < g.z_lo || zi > g.z_hi) return false;                      // _render.pyx:275
        if (w != w) return false;                                          // :278
        if (h < g.min_smooth) h = g.min_smooth;                            // :297
    }
    if (h < g.pdx * g.smooth_lo || h > g.pdx * g.smooth_hi) return false;  // :299, :438
    if (!g.grid) {
        if (!((g.use_z * fabs(zi - g.z0) < g.max_d * h) && xi > g.x1 - 2 * h && xi < g.x2 + 2 * h
              && yi > g.y1 - 2 * h && yi < g.y2 + 2 * h))
            return false;                                                  // :303-306
    } else {
        if (!(zi > g.z1 - 2 * h && zi < g.z2 + 2 * h && xi > g.x1 - 2 * h && xi < g.x2 + 2 * h
              && yi > g.y1 - 2 * h && yi < g.y2 + 2 * h))
            return false;                                                  // :441-444
    }
    f.xi = xi; f.yi = yi; f.zi = zi; f.h = h; f.w = w;
    f.za = 0; f.zb = 1;
    if (g.max_d * h / g.pdx < 1 && g.max_d * h / g.pdy < 1) {              // single pixel (:316, :455)
        const int px = c_int((xi - g.x1) / g.pdx), py = c_int((yi - g.y1) / g.pdy);
        if (!(px >= 0 && px < g.nx && py >= 0 && py < g.ny)) return false;
        f.xa = px; f.xb = px + 1; f.ya = py; f.yb = py + 1;
        if (g.grid) {
            const int pz = c_int((zi - g.z1) / g.pdz);
            if (!(pz >= 0 && pz < g.nz)) return false;
            f.za = pz; f.zb = pz + 1;
        }
        return true;
    }
    int xa = c_int((xi - g.max_d * h - g.x1) / g.pdx), xb = c_int((xi + g.max_d * h - g.x1) / g.pdx);
    int ya = c_int((yi - g.max_d * h - g.y1) / g.pdy), yb = c_int((yi + g.max_d * h - g.y1) / g.pdy);
    if (xa < 0) xa = 0; if (xb > g.nx) xb = g.nx;
    if (ya < 0) ya = 0; if (yb > g.ny) yb = g.ny;
    f.xa = xa; f.xb = xb; f.ya = ya; f.yb = yb;
    if (g.grid) {
        int za = c_int((zi - g.max_d * h - g.z1) / g.pdz), zb = c_int((zi + g.max_d * h - g.z1) / g.pdz);
        if (za < 0) za = 0; if (zb > g.nz) zb = g.nz;
        f.za = za; f.zb = zb;
    }
    return xb > xa && yb > ya && f.zb > f.za;
}

__device__ __forceinline__ void tile_span(const Geom &g, const Footprint &f, int &ta, int &tb, int &ua, int &ub,
                                          int &va, int &vb)
{
    if (!g.grid) {
        ta = f.xa / TILE_X; tb = (f.xb - 1) / TILE_X;
        ua = f.ya / TILE_Y; ub = (f.yb - 1) / TILE_Y;
        va = 0; vb = 0;
    } else {
        ta = f.xa / GTX; tb = (f.xb - 1) / GTX;
        ua = f.ya / GTY; ub = (f.yb - 1) / GTY;
        va = f.za / GTZ; vb = (f.zb - 1) / GTZ;
    }
}

// ---- per-particle weight and smoothing length as double (C promotion rules of _render.pyx:276) -------
__global__ void k_prepare(int64_t n, const void *sm, int sm64, const void *qty, int q64, const void *mass, int m64,
                          const void *rho, int r64, double *h_out, double *w_out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    h_out[i] = sm64 ? ((const double *)sm)[i] : (double)((const float *)sm)[i];
    double w;
    if (!q64 && !m64) {
        float qm = ((const float *)qty)[i] * ((const float *)mass)[i];
        if (!r64) w = (double)(qm / ((const float *)rho)[i]);
        else w = (double)qm / ((const double *)rho)[i];
    } else {
        double q = q64 ? ((const double *)qty)[i] : (double)((const float *)qty)[i];
        double m = m64 ? ((const double *)mass)[i] : (double)((const float *)mass)[i];
        double r = r64 ? ((const double *)rho)[i] : (double)((const float *)rho)[i];
        w = q * m / r;
    }
    w_out[i] = w;
}

__global__ void k_count_tiles(Geom g, Cols c, int64_t i0, int64_t i1, Images im,
                              unsigned long long *__restrict__ counts)
{
    int64_t i = i0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= i1) return;
    unsigned long long cnt = 0;
    for (int o = 0; o < im.n; ++o) {
        Footprint f;
        if (footprint(g, c, i, im.ox[o], im.oy[o], im.oz[o], f)) {
            int ta, tb, ua, ub, va, vb;
            tile_span(g, f, ta, tb, ua, ub, va, vb);
            cnt += (unsigned long long)(tb - ta + 1) * (unsigned long long)(ub - ua + 1) * (unsigned long long)(vb - va + 1);
        }
    }
    counts[i - i0] = cnt;
}

__global__ void k_emit_pairs(Geom g, Cols c, int64_t i0, int64_t i1, Images im,
                             const unsigned long long *__restrict__ offsets, uint32_t *__restrict__ keys,
                             uint32_t *__restrict__ vals)
{
    int64_t i = i0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= i1) return;
    unsigned long long o = offsets[i - i0];
    for (int img = 0; img < im.n; ++img) {
        Footprint f;
        if (!footprint(g, c, i, im.ox[img], im.oy[img], im.oz[img], f)) continue;
        int ta, tb, ua, ub, va, vb;
        tile_span(g, f, ta, tb, ua, ub, va, vb);
        const uint32_t payload = (uint32_t)(i - i0) | ((uint32_t)img << IMAGE_SHIFT);
        for (int t = ta; t <= tb; ++t)
            for (int u = ua; u <= ub; ++u)
                for (int v = va; v <= vb; ++v) {
                    uint32_t tile = g.grid ? (uint32_t)((t * g.tiles_y + u) * g.tiles_z + v) : (uint32_t)(u * g.tiles_x + t);
                    keys[o] = tile;
                    vals[o] = payload;
                    ++o;
                }
    }
}

__global__ void k_make_records(Geom g, Cols c, int64_t i0, Images im, int64_t npairs,
                               const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                               Rec *__restrict__ recs)
{
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= npairs) return;
    const uint32_t tile = keys[j];
    const uint32_t payload = vals[j];
    const int img = (int)(payload >> IMAGE_SHIFT);
    Footprint f;
    footprint(g, c, i0 + (payload & ((1u << IMAGE_SHIFT) - 1u)), im.ox[img], im.oy[img], im.oz[img], f);
    int tx, ty, tz = 0, px0, py0, pz0 = 0, sx, sy, sz = 1;
    if (!g.grid) {
        ty = tile / g.tiles_x; tx = tile - ty * g.tiles_x;
        px0 = tx * TILE_X; py0 = ty * TILE_Y; sx = TILE_X; sy = TILE_Y;
    } else {
        tz = tile % g.tiles_z; int r = tile / g.tiles_z; ty = r % g.tiles_y; tx = r / g.tiles_y;
        px0 = tx * GTX; py0 = ty * GTY; pz0 = tz * GTZ; sx = GTX; sy = GTY; sz = GTZ;
    }
    Rec R;
    const double h = f.h;
    const float hk = (g.h_power == 2) ? (float)(h * h) : (float)(h * h * h);        // :309-312
    const double kmax2 = (h * h) * (g.max_d * g.max_d);                             // :313
    R.xi = f.xi; R.yi = f.yi; R.kmax2 = kmax2;
    R.xr = (float)(f.xi - (g.x1 + g.pdx * px0));
    R.yr = (float)(f.yi - (g.y1 + g.pdy * py0));
    if (!g.grid) {
        const double dz = (f.zi - g.z0) * g.use_z;                                  // :314
        R.zd = dz;
        R.zr = (float)(dz * dz);
    } else {
        R.zd = f.zi;
        R.zr = (float)(f.zi - (g.z1 + g.pdz * pz0));
    }
    const bool single = g.max_d * h / g.pdx < 1 && g.max_d * h / g.pdy < 1;
    const float inv = (float)((double)g.ns / kmax2);
    R.inv = (single || !(inv > 0.0f) || inv > 1e37f) ? -1.0f : inv;                 // tiny kernels: exact path only
    R.wq = (float)(f.w / (double)hk);
    const int lxa = max(f.xa - px0, 0), lxb = min(f.xb - px0, sx);
    const int lya = max(f.ya - py0, 0), lyb = min(f.yb - py0, sy);
    const int lza = max(f.za - pz0, 0), lzb = min(f.zb - pz0, sz);
    if (!g.grid) {
        // image tiles: the rectangle as a lane mask (columns) and a row mask, so that the tile kernel's tests are single
        // bit operations -- the unpack-and-compare of a packed rectangle cost 14 of its ~67 instructions per record and warp
        const int wx = lxb - lxa, wy = lyb - lya;
        R.box = wx <= 0 ? 0u : ((wx >= 32 ? 0xffffffffu : ((1u << wx) - 1u)) << lxa);
        R.boxz = wy <= 0 ? 0u : ((((1u << wy) - 1u) << lya) & 0xffffu);
    } else {
        R.box = (uint32_t)lxa | ((uint32_t)(lxb - lxa) << 8) | ((uint32_t)lya << 16) | ((uint32_t)(lyb - lya) << 24);
        R.boxz = (uint32_t)lza | ((uint32_t)(lzb - lza) << 8);
    }
    R.tile = tile;
    recs[j] = R;
}

// ---- TMA bulk copy + mbarrier (PTX) --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_bulk(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// the reference's table lookup index (_render.pyx:186-193) in exact double arithmetic
template <bool GRID>
__device__ __forceinline__ unsigned exact_index(const Geom &g, const Rec &R, int px, int py, int pz)
{
    const double ddx = R.xi - (g.pdx * (double)px + g.xs);
    const double ddy = R.yi - (g.pdy * (double)py + g.ys);
    double d2;
    if (GRID) {
        const double ddz = R.zd - (g.pdz * (double)pz + g.zs);
        d2 = ddx * ddx + ddy * ddy + ddz * ddz;
    } else {
        d2 = ddx * ddx + ddy * ddy + R.zd * R.zd;
    }
    const double t = g.ns * (d2 / R.kmax2);
    return (t >= 4294967296.0 || t != t) ? 0xffffffffu : (unsigned)t;
}

template <bool GRID>
__global__ void __launch_bounds__(256) k_render_tiles(const Rec *__restrict__ recs, int64_t npairs, Geom g,
                                                      const float *__restrict__ lut, double *__restrict__ acc)
{
    extern __shared__ __align__(128) unsigned char smem[];
    Rec *buf = reinterpret_cast<Rec *>(smem);                                  // [2][CHUNK]
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 2 * CHUNK * sizeof(Rec));
    float *slut = reinterpret_cast<float *>(smem + 2 * CHUNK * sizeof(Rec) + 16);
    const int tid = threadIdx.x;
    const int64_t first = (int64_t)blockIdx.x * SEG;
    const int64_t last = min(npairs, first + (int64_t)SEG);
    const int nchunks = (int)((last - first + CHUNK - 1) / CHUNK);

    for (int i = tid; i < g.ns + 3; i += 256) slut[i] = i < g.ns ? lut[i] : 0.0f;
    // the table is read through its 32-bit shared-window address: with a generic pointer nvcc rebuilds the window base
    // (S2UR SR_CgaCtaId + ULEA) in front of EVERY lookup of the record loop -- 6 % of the kernel's stall samples (ncu)
    uint32_t slut_s = (uint32_t)__cvta_generic_to_shared(slut);
    asm volatile("mov.u32 %0, %0;" : "+r"(slut_s));           // opaque: computed here once, kept in a register
    auto table = [&](unsigned idx) {
        float v;
        asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(slut_s + idx * 4u));
        return v;
    };
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int c) {
        const int64_t b = first + (int64_t)c * CHUNK;
        const uint32_t bytes = (uint32_t)(min((int64_t)CHUNK, last - b) * sizeof(Rec));
        mbar_expect_tx(&bar[c & 1], bytes);
        tma_load_bulk(buf + (c & 1) * CHUNK, recs + b, bytes, &bar[c & 1]);
    };
    if (tid == 0) issue(0);

    // Image tiles are 32 x 16 pixels: a warp is one row of 32 pixels (lane = x) and takes rows 2w and 2w + 1, so the
    // row test of a record's rectangle is warp-uniform -- rows outside cost the warp three instructions instead of
    // idling its lanes through the pixel arithmetic.  Grid tiles are 4 x 8 x 8 voxels, one per thread.
    constexpr int NPIX = GRID ? 1 : 2;
    int lx, ly, lz = 0;
    if (GRID) { lz = tid & (GTZ - 1); ly = (tid >> 3) & (GTY - 1); lx = tid >> 6; }
    else { lx = tid & 31; ly = tid >> 5; }
    const float cx = (float)(g.pdx * (lx + 0.5)), cz = (float)(g.pdz * (lz + 0.5));
    float cy[NPIX];
#pragma unroll
    for (int q = 0; q < NPIX; ++q) cy[q] = (float)(g.pdy * (NPIX * ly + q + 0.5));
    const float nsf = (float)g.ns;
    uint32_t lanebit = 1u << (lx & 31);
    asm volatile("mov.u32 %0, %0;" : "+r"(lanebit));          // kept in a register (nvcc otherwise rebuilds lane and shift per record)

    // a pixel's partial sum over this CTA's slice of one tile (at most SEG terms) is kept in float, as the
    // reference keeps its whole sum; slices are combined in double
    float a[NPIX];
#pragma unroll
    for (int q = 0; q < NPIX; ++q) a[q] = 0.0f;
    uint32_t cur = 0xffffffffu;
    int px = 0, py = 0, pz = 0;                        // first pixel of this thread in the current tile
    auto flush = [&]() {
#pragma unroll
        for (int q = 0; q < NPIX; ++q) {
            if (a[q] != 0.0f) {
                const int64_t pix = GRID ? ((int64_t)px * g.ny + py) * g.nz + pz : (int64_t)(py + q) * g.nx + px;
                atomicAdd(&acc[pix], (double)a[q]);
            }
            a[q] = 0.0f;
        }
    };

    for (int c = 0; c < nchunks; ++c) {
        if (tid == 0 && c + 1 < nchunks) issue(c + 1);       // buffer (c+1)&1 was released by the barrier below
        mbar_wait(&bar[c & 1], (uint32_t)((c >> 1) & 1));
        const Rec *chunk = buf + (c & 1) * CHUNK;
        const int cnt = (int)min((int64_t)CHUNK, last - (first + (int64_t)c * CHUNK));
        for (int r = 0; r < cnt; ++r) {
            const Rec &R = chunk[r];
            if (R.tile != cur) {                               // block-uniform
                flush();
                cur = R.tile;
                if (GRID) {
                    const int tz = cur % g.tiles_z, q = cur / g.tiles_z;
                    pz = tz * GTZ + lz; py = (q % g.tiles_y) * GTY + ly; px = (q / g.tiles_y) * GTX + lx;
                } else {
                    const int ty = cur / g.tiles_x;
                    py = ty * TILE_Y + NPIX * ly; px = (cur - ty * g.tiles_x) * TILE_X + lx;
                }
            }
            const uint32_t box = R.box;
            // rows of this thread inside the record's rectangle (image: the same for the whole warp, one bit per row)
            const int row0 = GRID ? ly : NPIX * ly;
            unsigned rel = 0, wy = 0;
            uint32_t rows = 0;
            if (GRID) {
                rel = (unsigned)(row0 - (int)((box >> 16) & 0xff)); wy = box >> 24;
                if (!(rel < wy)) continue;
                bool in_xz = (unsigned)(lx - (int)(box & 0xff)) < ((box >> 8) & 0xff);
                in_xz &= (unsigned)(lz - (int)(R.boxz & 0xff)) < ((R.boxz >> 8) & 0xff);
                if (!in_xz) continue;
            } else {
                rows = R.boxz >> (NPIX * ly);               // bit q: this warp's row q lies inside the rectangle
                if (!(rows & (NPIX > 1 ? 3u : 1u))) continue;
                if (!(box & lanebit)) continue;
            }
            // (fused multiply-adds here: the float index is only trusted away from the bin borders -- BORDER below -- and
            // re-evaluated exactly otherwise, so contraction cannot change a table index)
            const float ddx = R.xr - cx;
            float d2xz;
            if (GRID) { const float ddz = R.zr - cz; d2xz = fmaf(ddz, ddz, ddx * ddx); }
            else d2xz = fmaf(ddx, ddx, R.zr);
            // every pixel of the rectangle receives w * (LUT or 0) as in the reference, so that a non-finite
            // weight poisons the same pixels (w * 0 = NaN); the table in shared memory is followed by zeros, so
            // indices past its end need no test
            const float inv = R.inv, wq = R.wq;
            const unsigned ns = (unsigned)g.ns;
            if (inv > 0.0f) {                                  // (a property of the record: uniform)
#pragma unroll
                for (int q = 0; q < NPIX; ++q) {
                    // image: a row of the warp's pair outside the rectangle (its top or bottom edge: ~5 % of the records of
                    // a warp) is evaluated with weight 0 instead of branched around -- cheaper than a test and a branch per pixel
                    if (GRID && !(rel + (unsigned)q < wy)) continue;
                    const float wqq = GRID ? wq : (((rows >> q) & 1u) ? wq : 0.0f);
                    const float ddy = R.yr - cy[q];
                    const float d2 = fmaf(ddy, ddy, d2xz);
                    // (clamped to a value whose fraction is 0.5, so that clamped indices never look like a bin border)
                    const float t = fminf(d2 * inv, nsf + 2.5f);
                    unsigned idx = (unsigned)t;
                    const float fr = t - (float)idx;
                    if (fabsf(fr - 0.5f) > 0.5f - BORDER) {
                        int pxo = px;
                        asm volatile("" : "+r"(pxo));          // keeps the double-precision pixel arithmetic of the exact
                        idx = min(exact_index<GRID>(g, R, pxo, py + q, pz), ns);   // path inside this rare branch (nvcc hoists it)
                    }
                    a[q] = fmaf(table(idx), wqq, a[q]);
                }
            } else {                                           // tiny kernels: exact index only
#pragma unroll
                for (int q = 0; q < NPIX; ++q) {
                    if (GRID ? !(rel + (unsigned)q < wy) : !(rows & (1u << q))) continue;
                    a[q] = fmaf(table(min(exact_index<GRID>(g, R, px, py + q, pz), ns)), wq, a[q]);
                }
            }
        }
        __syncthreads();
    }
    flush();
}

// ---- exact per-particle scatter (perspective images; also the cross-check path) ------------------------
__global__ void k_render_scatter(Geom g, Cols c, float ox, float oy, const float *__restrict__ lut,
                                 double *__restrict__ acc)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= c.n) return;
    Geom l = g;                                               // per-particle pixel grid in perspective mode
    const double zi = ld_pos(c.z, c.pos_f64, i, 0.0f);
    if (g.z_camera != 0.0) {
        if (zi < g.z_lo || zi > g.z_hi) return;
        if ((zi > g.z_camera && g.z_camera > 0) || (zi < g.z_camera && g.z_camera < 0)) return;   // :281-283
        const float dzc = (float)(g.z_camera - zi);
        l.x1 = g.mid_x - g.per_z_dx * dzc; l.x2 = g.mid_x + g.per_z_dx * dzc;                     // float arithmetic, :284-290
        l.y1 = g.mid_y - g.per_z_dy * dzc; l.y2 = g.mid_y + g.per_z_dy * dzc;
        l.pdx = (l.x2 - l.x1) / g.nx; l.pdy = (l.y2 - l.y1) / g.ny;
        l.xs = l.x1 + l.pdx / 2; l.ys = l.y1 + l.pdy / 2;
    }
    Footprint f;
    if (!footprint(l, c, i, ox, oy, 0.0f, f)) return;
    const double h = f.h;
    const float hk = (g.h_power == 2) ? (float)(h * h) : (float)(h * h * h);
    const double kmax2 = (h * h) * (g.max_d * g.max_d);
    const double dz = (f.zi - g.z0) * g.use_z;
    const int wy = f.yb - f.ya, npx = (f.xb - f.xa) * wy;
    for (int t = lane; t < npx; t += 32) {
        const int px = f.xa + t / wy, py = f.ya + t % wy;
        const double ddx = f.xi - (l.pdx * (double)px + l.xs), ddy = f.yi - (l.pdy * (double)py + l.ys);
        const double d2 = ddx * ddx + ddy * ddy + dz * dz;
        const double tt = g.ns * (d2 / kmax2);
        const unsigned idx = (tt >= 4294967296.0 || tt != tt) ? 0xffffffffu : (unsigned)tt;
        if (idx < (unsigned)g.ns) atomicAdd(&acc[(int64_t)py * g.nx + px], f.w * (double)(lut[idx] / hk));
    }
}

__global__ void k_finalize(const double *__restrict__ acc, int64_t n, float *__restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)acc[i];
}

inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

struct Work {
    cudaStream_t s = lib_stream();
    DevBuf counts, offsets, keys, vals, keys2, vals2, recs, scan_tmp, sort_tmp;
    size_t scan_bytes = 0, sort_bytes = 0;
};

// render particles [i0, i1) with all periodic images of `im` through the tile path in one pass
void render_range(Work &wk, const Geom &g, const Cols &c, int64_t i0, int64_t i1, const Images &im,
                  const float *lut_dev, double *acc, double &kernel_ms)
{
    typedef unsigned long long u64;
    const int64_t m = i1 - i0;
    if (m <= 0) return;
    cudaStream_t s = wk.s;
    const int BS = 256;
    wk.counts.alloc(sizeof(u64) * (m + 1));
    wk.offsets.alloc(sizeof(u64) * (m + 1));
    PNB_CUDA(cudaMemsetAsync(wk.counts.as<u64>() + m, 0, sizeof(u64), s));
    PNB_LAUNCH(k_count_tiles, nblk(m, BS), BS, 0, s, g, c, i0, i1, im, wk.counts.as<u64>());
    size_t need = 0;
    PNB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, wk.counts.as<u64>(), wk.offsets.as<u64>(), m + 1, s));
    if (need > wk.scan_tmp.bytes) wk.scan_tmp.alloc(need);
    PNB_CUDA(cub::DeviceScan::ExclusiveSum(wk.scan_tmp.p, need, wk.counts.as<u64>(), wk.offsets.as<u64>(), m + 1, s));
    g_launches += 2;
    u64 total = 0;
    PNB_CUDA(cudaMemcpyAsync(&total, wk.offsets.as<u64>() + m, sizeof(u64), cudaMemcpyDeviceToHost, s));
    PNB_CUDA(cudaStreamSynchronize(s));
    if (total == 0) return;
    if (((int64_t)total > MAX_PAIRS && m > 1) || m > ((int64_t)1 << IMAGE_SHIFT)) {
        // too many pairs (or particles) for one batch: footprints are independent and the sums linear, so split
        const int64_t mid = i0 + m / 2;
        render_range(wk, g, c, i0, mid, im, lut_dev, acc, kernel_ms);
        render_range(wk, g, c, mid, i1, im, lut_dev, acc, kernel_ms);
        return;
    }
    const int64_t P = (int64_t)total;
    wk.keys.alloc(sizeof(uint32_t) * P); wk.vals.alloc(sizeof(uint32_t) * P);
    wk.keys2.alloc(sizeof(uint32_t) * P); wk.vals2.alloc(sizeof(uint32_t) * P);
    PNB_LAUNCH(k_emit_pairs, nblk(m, BS), BS, 0, s, g, c, i0, i1, im, wk.offsets.as<u64>(),
               wk.keys.as<uint32_t>(), wk.vals.as<uint32_t>());
    const int64_t ntiles = (int64_t)g.tiles_x * g.tiles_y * g.tiles_z;
    int bits = 1;
    while (((int64_t)1 << bits) < ntiles) ++bits;
    need = 0;
    PNB_CUDA((cub::DeviceRadixSort::SortPairs<uint32_t, uint32_t, int64_t>(
        nullptr, need, wk.keys.as<uint32_t>(), wk.keys2.as<uint32_t>(), wk.vals.as<uint32_t>(), wk.vals2.as<uint32_t>(), P, 0, bits, s)));
    if (need > wk.sort_tmp.bytes) wk.sort_tmp.alloc(need);
    PNB_CUDA((cub::DeviceRadixSort::SortPairs<uint32_t, uint32_t, int64_t>(
        wk.sort_tmp.p, need, wk.keys.as<uint32_t>(), wk.keys2.as<uint32_t>(), wk.vals.as<uint32_t>(), wk.vals2.as<uint32_t>(), P, 0, bits, s)));
    g_launches += 3;
    wk.recs.alloc(sizeof(Rec) * P);
    PNB_LAUNCH(k_make_records, nblk(P, BS), BS, 0, s, g, c, i0, im, P, wk.keys2.as<uint32_t>(),
               wk.vals2.as<uint32_t>(), wk.recs.as<Rec>());
    const size_t smem = 2 * CHUNK * sizeof(Rec) + 16 + sizeof(float) * ((size_t)g.ns + 3);
    KernelTimer timer(2, s);
    if (g.grid) {
        PNB_CUDA(cudaFuncSetAttribute(k_render_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PNB_LAUNCH((k_render_tiles<true>), nblk(P, SEG), 256, smem, s, wk.recs.as<Rec>(), P, g, lut_dev, acc);
    } else {
        PNB_CUDA(cudaFuncSetAttribute(k_render_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PNB_LAUNCH((k_render_tiles<false>), nblk(P, SEG), 256, smem, s, wk.recs.as<Rec>(), P, g, lut_dev, acc);
    }
    timer.stop();
    kernel_ms += pnb_last_kernel_ms(2);
}

// x, y, z as three dense device columns inside `dpos` (ps * 3 * n bytes).  The usual caller passes the columns of one
// C-contiguous N x 3 position array (renderers.py:660-665: sim['x'], sim['y'], sim['z']); from host memory those come
// over in ONE contiguous copy and are separated on the device.
void fetch_positions(const pnb_render_arrays *a, DevBuf &dpos, void *&dx, void *&dy, void *&dz, cudaStream_t s)
{
    const int64_t n = a->n;
    const int64_t ps = (int64_t)tsize(a->pos_dtype);
    dx = dpos.p; dy = (char *)dpos.p + ps * n; dz = (char *)dpos.p + 2 * ps * n;
    const char *x = (const char *)a->x, *y = (const char *)a->y, *z = (const char *)a->z;
    if (y == x + ps && z == x + 2 * ps && a->pos_stride == 3 * ps) {
        fetch_columns(a->x, 3 * ps, ps, 3, a->pos_dtype, a->mem, n, dpos.p, s);
        return;
    }
    fetch_columns(a->x, a->pos_stride, 0, 1, a->pos_dtype, a->mem, n, dx, s);
    fetch_columns(a->y, a->pos_stride, 0, 1, a->pos_dtype, a->mem, n, dy, s);
    fetch_columns(a->z, a->pos_stride, 0, 1, a->pos_dtype, a->mem, n, dz, s);
}

void check_arrays(const pnb_render_arrays *a)
{
    if (!a) throw Error(PNB_ERR_VALUE, "render arrays are null");
    if (a->n < 0) throw Error(PNB_ERR_VALUE, "negative particle count");
    const int dts[5] = {a->pos_dtype, a->sm_dtype, a->qty_dtype, a->mass_dtype, a->rho_dtype};
    for (int d : dts)
        if (d != PNB_F32 && d != PNB_F64) throw Error(PNB_ERR_VALUE, "render arrays must be float32 or float64");
    if (a->n > 0 && (!a->x || !a->y || !a->z || !a->sm || !a->qty || !a->mass || !a->rho))
        throw Error(PNB_ERR_VALUE, "render array pointer is null");
    if (a->n >= ((int64_t)1 << 31)) throw Error(PNB_ERR_VALUE, "at most 2^31-1 particles per render call");   // _render.pyx:229
}

void render(bool grid, int nx, int ny, int nz, const pnb_render_arrays *a, double x1, double x2, double y1, double y2,
            double z1, double z2, double z_camera, double z0, double smooth_lo, double smooth_hi, double z_lo,
            double z_hi, double min_smooth, const float *lut, int n_lut, int h_power, double max_d,
            const double *wrap_x, int nwx, const double *wrap_y, int nwy, const double *wrap_z, int nwz, float *out,
            int out_mem)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw Error(PNB_ERR_CUDA, "no CUDA device available: pnb_b200 has no CPU fallback");
    }
    check_arrays(a);
    if (nx < 1 || ny < 1 || nz < 1) throw Error(PNB_ERR_VALUE, "image dimensions must be positive");
    if (!lut || n_lut < 1 || n_lut > 8192) throw Error(PNB_ERR_VALUE, "kernel table must hold 1..8192 samples");
    if (grid && h_power < 3) throw Error(PNB_ERR_VALUE, "Cannot render to 3D grid without 3-dimensional kernel or greater");
    if (!out) throw Error(PNB_ERR_VALUE, "output pointer is null");
    cudaStream_t s = lib_stream();
    StageTimer stage(2, s);

    Geom g;
    memset(&g, 0, sizeof(g));
    g.nx = nx; g.ny = ny; g.nz = nz; g.grid = grid ? 1 : 0;
    g.x1 = x1; g.x2 = x2; g.y1 = y1; g.y2 = y2; g.z1 = z1; g.z2 = z2;
    g.pdx = (x2 - x1) / nx; g.pdy = (y2 - y1) / ny;
    g.pdz = grid ? (z2 - z1) / ny : 1.0;                     // sic: _render.pyx:391 divides by ny
    g.xs = x1 + g.pdx / 2; g.ys = y1 + g.pdy / 2; g.zs = z1 + g.pdz / 2;
    g.z_camera = grid ? 0.0 : z_camera; g.z0 = z0;
    g.smooth_lo = smooth_lo; g.smooth_hi = smooth_hi; g.z_lo = z_lo; g.z_hi = z_hi; g.min_smooth = min_smooth;
    g.max_d = max_d; g.ns = n_lut; g.h_power = h_power; g.use_z = h_power >= 3 ? 1 : 0;
    g.per_z_dx = (float)((x2 - x1) / (2 * z_camera)); g.per_z_dy = (float)((y2 - y1) / (2 * z_camera));
    g.mid_x = (float)((x2 + x1) / 2); g.mid_y = (float)((y2 + y1) / 2);
    if (grid) { g.tiles_x = (nx + GTX - 1) / GTX; g.tiles_y = (ny + GTY - 1) / GTY; g.tiles_z = (nz + GTZ - 1) / GTZ; }
    else { g.tiles_x = (nx + TILE_X - 1) / TILE_X; g.tiles_y = (ny + TILE_Y - 1) / TILE_Y; g.tiles_z = 1; }
    if ((int64_t)g.tiles_x * g.tiles_y * g.tiles_z >= ((int64_t)1 << 31)) throw Error(PNB_ERR_VALUE, "image too large");

    const int64_t n = a->n, npix = (int64_t)nx * ny * nz;
    DevBuf acc(sizeof(double) * npix);
    PNB_CUDA(cudaMemsetAsync(acc.p, 0, sizeof(double) * npix, s));
    DevBuf lut_dev(sizeof(float) * n_lut);
    PNB_CUDA(cudaMemcpyAsync(lut_dev.p, lut, sizeof(float) * n_lut, cudaMemcpyHostToDevice, s));
    double kernel_ms = 0;
    if (n > 0) {
        const size_t ps = tsize(a->pos_dtype);
        DevBuf dpos(ps * 3 * n), dsm(tsize(a->sm_dtype) * n), dq, dm(tsize(a->mass_dtype) * n), dr(tsize(a->rho_dtype) * n),
            hd(sizeof(double) * n), wd(sizeof(double) * n);
        void *dxp, *dyp, *dzp;
        fetch_positions(a, dpos, dxp, dyp, dzp, s);
        fetch_columns(a->sm, a->sm_stride, 0, 1, a->sm_dtype, a->mem, n, dsm.p, s);
        fetch_columns(a->mass, a->mass_stride, 0, 1, a->mass_dtype, a->mem, n, dm.p, s);
        fetch_columns(a->rho, a->rho_stride, 0, 1, a->rho_dtype, a->mem, n, dr.p, s);
        const void *dqp = dr.p;                               // rendering rho itself (qty is the rho array): one copy serves both
        if (!(a->qty == a->rho && a->qty_stride == a->rho_stride && a->qty_dtype == a->rho_dtype)) {
            dq.alloc(tsize(a->qty_dtype) * n);
            fetch_columns(a->qty, a->qty_stride, 0, 1, a->qty_dtype, a->mem, n, dq.p, s);
            dqp = dq.p;
        }
        PNB_LAUNCH(k_prepare, nblk(n, 256), 256, 0, s, n, dsm.p, a->sm_dtype, dqp, a->qty_dtype, dm.p, a->mass_dtype,
                   dr.p, a->rho_dtype, hd.as<double>(), wd.as<double>());
        Cols c;
        c.x = dxp; c.y = dyp; c.z = dzp; c.pos_f64 = a->pos_dtype; c.h = hd.as<double>(); c.w = wd.as<double>(); c.n = n;
        Work wk;
        wk.s = s;
        const double zero = 0.0;
        if (!wrap_x || nwx < 1) { wrap_x = &zero; nwx = 1; }
        if (!wrap_y || nwy < 1) { wrap_y = &zero; nwy = 1; }
        if (!grid || !wrap_z || nwz < 1) { wrap_z = &zero; nwz = 1; }
        Images im;
        im.n = 0;
        for (int i = 0; i < nwx; ++i)
            for (int j = 0; j < nwy; ++j)
                for (int k = 0; k < nwz; ++k) {
                    const float ox = (float)wrap_x[i], oy = (float)wrap_y[j], oz = (float)wrap_z[k];   // :237-243, :418
                    if (!grid && z_camera != 0.0) {
                        KernelTimer timer(2, s);
                        PNB_LAUNCH(k_render_scatter, nblk(n * 32, 256), 256, 0, s, g, c, ox, oy, lut_dev.as<float>(), acc.as<double>());
                        timer.stop();
                        kernel_ms += pnb_last_kernel_ms(2);
                        continue;
                    }
                    // all periodic images go through the tile pipeline together (up to MAX_IMAGES per pass)
                    im.ox[im.n] = ox; im.oy[im.n] = oy; im.oz[im.n] = oz;
                    if (++im.n == MAX_IMAGES) {
                        render_range(wk, g, c, 0, n, im, lut_dev.as<float>(), acc.as<double>(), kernel_ms);
                        im.n = 0;
                    }
                }
        if (im.n > 0) render_range(wk, g, c, 0, n, im, lut_dev.as<float>(), acc.as<double>(), kernel_ms);
        PNB_CUDA(cudaStreamSynchronize(s));
    }
    if (out_mem == PNB_DEVICE) {
        PNB_LAUNCH(k_finalize, nblk(npix, 256), 256, 0, s, acc.as<double>(), npix, out);
        PNB_CUDA(cudaStreamSynchronize(s));
    } else {
        DevBuf of(sizeof(float) * npix);
        PNB_LAUNCH(k_finalize, nblk(npix, 256), 256, 0, s, acc.as<double>(), npix, of.as<float>());
        PNB_CUDA(cudaMemcpyAsync(out, of.p, sizeof(float) * npix, cudaMemcpyDeviceToHost, s));
        PNB_CUDA(cudaStreamSynchronize(s));
    }
    stage.stop();
    set_kernel_ms(2, kernel_ms);
}

// ================================================================================================
// Healpix renderer.  Replaces _render.render_spherical_image_core (pynbody/sph/_render.pyx:51-179) and
// its disc query query_disc_c (pynbody/sph/healpix.c:94-204, RING scheme).  Like the reference this is
// a per-particle scatter over the pixels of the particle's angular disc -- the disc is found ring by
// ring, exactly as healpix.c does it -- but one WARP takes one particle: the ring loop is warp-uniform
// and the lanes share the pixels of a ring; contributions go to a double accumulator with atomicAdd.
// ================================================================================================
struct HpArgs {
    const void *x, *y, *z, *h, *qty, *mass, *rho;
    int pos_f64, h_f64, qty_f64, mass_f64, rho_f64;
    int64_t n;
    int nside, ns, projected;
    double max_d, shell_radius;
};

struct HpRing {                 // one ring of a particle's disc (healpix.c:120-200)
    double st, ct;              // sin / cos of the ring's colatitude
    unsigned ip_lo, n_in_ring, shift, iring;
    unsigned long long base;    // index of the ring's pixel "0" (= hp_pixel_index(nside, iring, 0)); pixel ip is base + ip
};

__device__ __forceinline__ double ldv(const void *p, int f64, int64_t i)
{
    return f64 ? ((const double *)p)[i] : (double)((const float *)p)[i];
}
__device__ __forceinline__ size_t hp_ring_num_from_z(size_t nside, double z)            // healpix.c:24-43
{
    const size_t nring = 4 * nside;
    size_t iring;
    const double nd = (double)nside;
    if (z > 2.0 / 3.0) iring = (size_t)(nd * sqrt(3.0 * (1.0 - z)) + 0.5);
    else if (z >= -2.0 / 3.0) iring = (size_t)(nd * (2.0 - 1.5 * z) + 0.5);
    else iring = (size_t)(nring - nd * sqrt(3.0 * (1.0 + z)) + 0.5);
    if (iring < 1) iring = 1;
    if (iring > nring) iring = nring;
    return iring;
}
__device__ __forceinline__ double hp_z_from_ring_num(size_t nside, size_t iring)        // healpix.c:46-64
{
    const size_t nl4 = 4 * nside;
    if (iring <= nside) { const double t = (double)iring; return 1.0 - (t * t) / (3.0 * nside * nside); }
    if (iring <= 3 * nside) return (2 * (double)nside - (double)iring) / (1.5 * nside);
    const double t = (double)(nl4 - iring);
    return -1.0 + (t * t) / (3.0 * nside * nside);
}
__device__ __forceinline__ size_t hp_pixel_index(size_t nside, size_t iring, size_t iphi)   // healpix.c:67-91
{
    const size_t npix = 12 * nside * nside, ncap = 2 * nside * (nside + 1);
    if (iring <= nside) return iring * (iring - 1) * 2 + iphi - 1;
    if (iring <= 3 * nside) return ncap + (iring - nside - 1) * 4 * nside + iphi - 1;
    const size_t nr = 4 * nside - iring;
    return npix - nr * (nr - 1) * 2 + iphi - 1 - 4 * nr;
}

// Unit vector of every pixel centre, (st cos(phi), st sin(phi), ct, -) exactly as the per-pixel code forms them
// (healpix.c pix2vec for the RING scheme).  Filled once per call so that a (particle, pixel) pair costs a 32-byte
// load instead of a double-precision sincos.
__global__ void k_hp_pixel_vectors(int nside_i, double4 *__restrict__ vec)
{
    const double PI = 3.14159265358979323846, twopi = 2.0 * PI;
    const size_t nside = (size_t)nside_i;
    const size_t iring = (size_t)blockIdx.y + 1;                 // 1 .. 4 nside - 1
    size_t n_in_ring, shift;
    if (iring <= nside) { n_in_ring = 4 * iring; shift = 1; }
    else if (iring <= 3 * nside) { n_in_ring = 4 * nside; shift = 2 - (iring - nside + 1) % 2; }
    else { n_in_ring = 4 * (4 * nside - iring); shift = 1; }
    const size_t ip = (size_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (ip > n_in_ring) return;
    const double z = hp_z_from_ring_num(nside, iring);
    const double theta = acos(z);
    double st, ct;
    sincos(theta, &st, &ct);
    double phi = (twopi * ((double)ip - 0.5 * ((double)shift)) / n_in_ring);
    if (phi >= twopi) phi -= twopi;
    double sphi, cphi;
    sincos(phi, &sphi, &cphi);
    vec[hp_pixel_index(nside, iring, ip)] = make_double4(st * cphi, st * sphi, ct, 0.0);
}

// (six resident CTAs per SM: the kernel waits on the gathers of the pixel-centre table, not on arithmetic -- 3 CTAs of 80 registers:
// 95.7 ms, 4 of 64: 76.9, 5 of 48: 70.5, 6 of 40: 67.8 ms for the 6.9 M-particle shell at nside 256)
__global__ void __launch_bounds__(256, 6) k_render_healpix(HpArgs a, const float *__restrict__ lut, double *__restrict__ acc,
                                                        const double4 *__restrict__ pixvec)
{
    const double PI = 3.14159265358979323846, twopi = 2.0 * PI;
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= a.n) return;
    const double qty_i = ldv(a.qty, a.qty_f64, i);
    if (qty_i != qty_i) return;                                                          // _render.pyx:128-129
    double v0 = ldv(a.x, a.pos_f64, i), v1 = ldv(a.y, a.pos_f64, i), v2 = ldv(a.z, a.pos_f64, i);
    const double hi = ldv(a.h, a.h_f64, i);
    const double distance2 = v0 * v0 + v1 * v1 + v2 * v2;
    const double distance = sqrt(distance2);
    // h[i]*h[i] is evaluated in the element type of h (_render.pyx:136)
    const double smooth_2 = a.h_f64 ? hi * hi : (double)(((const float *)a.h)[i] * ((const float *)a.h)[i]);
    const double kernel_max_2 = smooth_2 * (a.max_d * a.max_d);
    const double shell_2 = a.shell_radius * a.shell_radius;
    double radius, two_rs_d = 0.0;
    float h_to_kdim;
    if (a.projected) {
        radius = a.max_d * hi / distance;
        h_to_kdim = (float)(smooth_2 / distance2);
    } else {
        two_rs_d = 2.0 * a.shell_radius * distance;
        const double cos_alpha = (shell_2 + distance2 - kernel_max_2) / two_rs_d;
        if (cos_alpha >= 1.0) return;
        radius = cos_alpha <= -1.0 ? PI : acos(cos_alpha);
        h_to_kdim = (float)(smooth_2 * hi);
    }
    // qty_i * kern * mass / rho with C's left-to-right promotion over the element types (_render.pyx:176)
    const size_t nside = (size_t)a.nside;
    const double vecnorm = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
    v0 /= vecnorm; v1 /= vecnorm; v2 /= vecnorm;
    const double z0 = v2, sin_theta0 = sqrt(1 - z0 * z0), phi0 = atan2(v1, v0), theta0 = acos(z0);
    double thetamin = theta0 - radius, thetamax = theta0 + radius;
    if (thetamin < 0.0) thetamin = 0.0;
    if (thetamax > PI) thetamax = PI;
    const size_t irmin = hp_ring_num_from_z(nside, cos(thetamin)), irmax = hp_ring_num_from_z(nside, cos(thetamax));
    const double cos_radius = cos(radius);
    const float radiusf = (float)radius, distancef = (float)distance, nsf = (float)a.ns;
    const float radius2f = radiusf * radiusf, ns_over_radius2f = (float)((double)a.ns / (radius * radius));
    const bool small_disc = a.projected && radiusf <= 0.2f;             // (warp-uniform: one particle per warp)
    // the particle's mass and density, loaded once (they enter every pixel's weight chain in their own element type)
    const double mass_d = ldv(a.mass, a.mass_f64, i), rho_d = ldv(a.rho, a.rho_f64, i);
    const float mass_f = (float)mass_d, rho_f = (float)rho_d;
    const float inv_kmax2f = (float)(1.0 / kernel_max_2);
    const double ns_over_kmax2 = (double)a.ns / kernel_max_2;
    // The disc is found ring by ring as healpix.c does it, 32 rings at a time with one ring per lane; the
    // pixels of those rings are then numbered consecutively and dealt out to the lanes, so that all lanes
    // are busy whatever the width of a ring (a disc of the usual size has ~12 pixels per ring).
    __shared__ HpRing rings[8][32];
    HpRing *wr = rings[threadIdx.x >> 5];
    for (size_t rb = irmin; rb <= irmax; rb += 32) {
        const size_t iring = rb + lane;
        int count = 0;
        if (iring <= irmax) {
            const double z = hp_z_from_ring_num(nside, iring), sin_theta = sqrt(1.0 - z * z);
            const double cos_dphi = (cos_radius - z * z0) / (sin_theta * sin_theta0);
            const double dphi = cos_dphi >= 1.0 ? 0.0 : (cos_dphi <= -1.0 ? PI : acos(cos_dphi));
            size_t ip_lo, ip_hi, n_in_ring, shift;
            if (iring <= nside) { n_in_ring = 4 * iring; shift = 1; }
            else if (iring <= 3 * nside) { n_in_ring = 4 * nside; shift = 2 - (iring - nside + 1) % 2; }
            else { n_in_ring = 4 * (4 * nside - iring); shift = 1; }
            if (dphi > 0.5 * PI) { ip_lo = 1; ip_hi = n_in_ring; }
            else {
                double phi_low = phi0 - dphi, phi_high = phi0 + dphi;
                if (phi_low < 0.0) phi_low += twopi;
                if (phi_high >= twopi) phi_high -= twopi;
                const double phi_factor = n_in_ring / twopi;
                ip_lo = (size_t)(phi_low * phi_factor + 0.5 * shift);
                ip_hi = (size_t)(phi_high * phi_factor + 0.5 * shift + 1);
            }
            if (ip_hi < ip_lo) ip_hi += n_in_ring;
            if (ip_hi - ip_lo > n_in_ring) { ip_lo = 1; ip_hi = n_in_ring; }
            const double theta = acos(z);
            sincos(theta, &wr[lane].st, &wr[lane].ct);
            wr[lane].ip_lo = (unsigned)ip_lo; wr[lane].n_in_ring = (unsigned)n_in_ring;
            wr[lane].shift = (unsigned)shift; wr[lane].iring = (unsigned)iring;
            wr[lane].base = (unsigned long long)hp_pixel_index(nside, iring, 0);
            count = (int)(ip_hi - ip_lo + 1);
        }
        int incl = count;                               // inclusive prefix sum over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int excl = incl - count;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        for (int fb = 0; fb < total; fb += 32) {        // (warp-uniform trip count: the shuffles need every lane)
            const int f = fb + lane;
            int r = 0;                                  // the ring that holds pixel number f: largest r with excl[r] <= f
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                const int t = __shfl_sync(0xffffffffu, excl, (r + d) & 31);
                if (r + d < 32 && t <= f) r += d;
            }
            const int first = __shfl_sync(0xffffffffu, excl, r);
            if (f >= total) continue;
            const HpRing R = wr[r];
            const size_t n_in_ring = R.n_in_ring, shift = R.shift;
            unsigned ip = R.ip_lo + (unsigned)(f - first);             // (32-bit: at most two turns of a ring)
            if (ip > R.n_in_ring) ip -= R.n_in_ring;
            const size_t ipix = (size_t)(R.base + ip);
            // Fast path (the pixel's centre vector comes from the table): the angle between the two unit vectors is
            // taken from their chord -- c^2 = |v - p|^2 in double without cancellation, angle = 2 asin(c / 2) in float
            // (relative error < 5e-7) -- instead of a double-precision acos(v . p) per pixel, and the table index
            // t = ns d2 / kmax2 is formed in float.  The reference's double expressions decide whenever the float
            // result is not safe: angle within 1e-5 of the disc radius, t within 5e-4 of an integer, nearly antipodal
            // directions.  For the thin-shell kernel cos(angle) = 1 - c^2/2 is used directly (no transcendental at all).
            unsigned idx = 0;
            bool exact = !(pixvec && ip >= 1);
            double4 pv = make_double4(0.0, 0.0, 0.0, 0.0);
            if (!exact) {
                pv = pixvec[ipix];
                const double ex = v0 - pv.x, ey = v1 - pv.y, ez = v2 - pv.z;
                const double chord2 = ex * ex + ey * ey + ez * ez;
                if (a.projected) {
                    const float c2 = (float)chord2;
                    if (small_disc && c2 <= 0.09f) {
                        // discs of the usual size: the SQUARED angle straight from the squared chord,
                        // (2 asin(c/2))^2 = c^2 (1 + c^2/12 + c^4/90 + c^6/560 + ...) (truncation < 2.2e-8 for c <= 0.3) -- no
                        // square root, no arcsine; t = ns d2 / kmax2 = (ns / radius^2) angle^2 in this mode
                        const float ang2 = c2 * fmaf(c2, fmaf(c2, fmaf(c2, 1.0f / 560.0f, 1.0f / 90.0f), 1.0f / 12.0f), 1.0f);
                        if (fabsf(ang2 - radius2f) <= 2e-5f * radius2f) exact = true;
                        else {
                            if (ang2 > radius2f) continue;
                            const float tf = ns_over_radius2f * ang2;
                            idx = (unsigned)fminf(tf, nsf + 2.0f);
                            const float fr = tf - (float)idx;
                            if (fabsf(fr - 0.5f) > 0.5f - 5e-4f && tf < nsf + 1.0f) exact = true;
                        }
                    } else {
                        const float chord = sqrtf(c2);
                        const float distf = 2.0f * asinf(fminf(0.5f * chord, 1.0f));
                        if (chord > 1.9f || fabsf(distf - radiusf) <= 1e-5f * radiusf) exact = true;
                        else {
                            if (distf > radiusf) continue;
                            const float off = distancef * distf;
                            const float tf = nsf * (off * off) * inv_kmax2f;
                            idx = (unsigned)fminf(tf, nsf + 2.0f);
                            const float fr = tf - (float)idx;
                            if (fabsf(fr - 0.5f) > 0.5f - 5e-4f && tf < nsf + 1.0f) exact = true;
                        }
                    }
                } else {
                    const double cosd = 1.0 - 0.5 * chord2;
                    if (fabs(cosd - cos_radius) <= 1e-12) exact = true;
                    else {
                        if (cosd < cos_radius) continue;
                        const double d2 = shell_2 + distance2 - two_rs_d * cosd;
                        const double td = d2 * ns_over_kmax2;
                        const double tc = td < 0.0 ? 0.0 : (td > 4.0e9 ? 4.0e9 : td);
                        idx = (unsigned)tc;
                        const double fr = tc - (double)idx;
                        if (fabs(fr - 0.5) > 0.5 - 1e-7 && td < (double)a.ns + 1.0) exact = true;
                    }
                }
            }
            if (exact) {
                double dot;
                if (!(pixvec && ip >= 1)) {      // (ip = 0 happens at phi ~ 0: the reference then takes the direction
                                                 // of "pixel 0" of the ring but the index before the ring's first)
                    double phi = (twopi * ((double)ip - 0.5 * ((double)shift)) / n_in_ring);
                    if (phi >= twopi) phi -= twopi;
                    double sphi, cphi;
                    sincos(phi, &sphi, &cphi);
                    dot = v0 * (R.st * cphi) + v1 * (R.st * sphi) + v2 * R.ct;
                } else {
                    double px_ = pv.x;
                    asm volatile("" : "+d"(px_));                  // (the dot product belongs to this rare branch only)
                    dot = v0 * px_ + v1 * pv.y + v2 * pv.z;
                }
                if (dot > 1.0) dot = 1.0;
                if (dot < -1.0) dot = -1.0;
                const double dist = acos(dot);
                if (!(dist <= radius)) continue;
                double d2;
                if (a.projected) { const double off = distance * dist; d2 = off * off; }
                else d2 = shell_2 + distance2 - two_rs_d * cos(dist);
                const double t = a.ns * (d2 / kernel_max_2);
                idx = (t >= 4294967296.0 || t != t) ? 0xffffffffu : (unsigned)t;
            }
            const float kern = idx < (unsigned)a.ns ? lut[idx] / h_to_kdim : 0.0f;
            bool isd = a.qty_f64;
            double vd = isd ? qty_i * (double)kern : 0.0;
            float vf = isd ? 0.0f : (float)qty_i * kern;
            if (isd || a.mass_f64) { vd = (isd ? vd : (double)vf) * mass_d; isd = true; }
            else vf = vf * mass_f;
            if (isd || a.rho_f64) { vd = (isd ? vd : (double)vf) / rho_d; isd = true; }
            else vf = vf / rho_f;
            // The reference's ip == 0 case on the first ring gives pixel index (size_t)-1: it writes 4 bytes BEFORE its
            // image there (an out-of-bounds write on the host).  That contribution is dropped here.
            if (ipix < (size_t)12 * nside * nside) atomicAdd(&acc[ipix], isd ? vd : (double)vf);
        }
        __syncwarp();
    }
}

void render_healpix(const pnb_render_arrays *a, int nside, const float *lut, int n_lut, int h_power, double max_d,
                    double shell_radius, float *out, int out_mem)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw Error(PNB_ERR_CUDA, "no CUDA device available: pnb_b200 has no CPU fallback");
    }
    check_arrays(a);
    if (nside < 1 || (nside & (nside - 1))) throw Error(PNB_ERR_VALUE, "nside value must be a power of 2");
    if (h_power != 2 && h_power != 3)
        throw Error(PNB_ERR_VALUE, "Only kernels of dimension 2 (projected) or 3 (thin shell) are supported for healpix maps");
    if (h_power == 3 && !(shell_radius > 0.0))
        throw Error(PNB_ERR_VALUE, "A positive shell_radius is required to render a 3D (thin shell) healpix map");
    if (!lut || n_lut < 1 || n_lut > 8192) throw Error(PNB_ERR_VALUE, "kernel table must hold 1..8192 samples");
    if (!out) throw Error(PNB_ERR_VALUE, "output pointer is null");
    cudaStream_t s = lib_stream();
    StageTimer stage(2, s);
    const int64_t n = a->n, npix = (int64_t)12 * nside * nside;
    DevBuf acc(sizeof(double) * npix), lut_dev(sizeof(float) * n_lut);
    PNB_CUDA(cudaMemsetAsync(acc.p, 0, sizeof(double) * npix, s));
    PNB_CUDA(cudaMemcpyAsync(lut_dev.p, lut, sizeof(float) * n_lut, cudaMemcpyHostToDevice, s));
    double kernel_ms = 0;
    if (n > 0) {
        const size_t ps = tsize(a->pos_dtype);
        DevBuf dpos(ps * 3 * n), dh(tsize(a->sm_dtype) * n), dq(tsize(a->qty_dtype) * n),
            dm(tsize(a->mass_dtype) * n), dr(tsize(a->rho_dtype) * n);
        void *dxp, *dyp, *dzp;
        fetch_positions(a, dpos, dxp, dyp, dzp, s);
        fetch_columns(a->sm, a->sm_stride, 0, 1, a->sm_dtype, a->mem, n, dh.p, s);
        fetch_columns(a->qty, a->qty_stride, 0, 1, a->qty_dtype, a->mem, n, dq.p, s);
        fetch_columns(a->mass, a->mass_stride, 0, 1, a->mass_dtype, a->mem, n, dm.p, s);
        fetch_columns(a->rho, a->rho_stride, 0, 1, a->rho_dtype, a->mem, n, dr.p, s);
        HpArgs h;
        h.x = dxp; h.y = dyp; h.z = dzp; h.h = dh.p; h.qty = dq.p; h.mass = dm.p; h.rho = dr.p;
        h.pos_f64 = a->pos_dtype; h.h_f64 = a->sm_dtype; h.qty_f64 = a->qty_dtype; h.mass_f64 = a->mass_dtype;
        h.rho_f64 = a->rho_dtype; h.n = n; h.nside = nside; h.ns = n_lut; h.projected = h_power == 2;
        h.max_d = max_d; h.shell_radius = shell_radius;
        // pixel-centre vectors, if the table fits comfortably (32 B per pixel: nside 2048 -> 1.6 GB)
        DevBuf pixvec;
        if (nside <= 2048) {
            pixvec.alloc(sizeof(double4) * (size_t)npix);
            dim3 grid((unsigned)((4 * (size_t)nside + 255) / 256), (unsigned)(4 * nside - 1));
            PNB_LAUNCH(k_hp_pixel_vectors, grid, 256, 0, s, nside, pixvec.as<double4>());
        }
        KernelTimer timer(2, s);
        PNB_LAUNCH(k_render_healpix, nblk(n * 32, 256), 256, 0, s, h, lut_dev.as<float>(), acc.as<double>(),
                   pixvec.p ? pixvec.as<double4>() : (const double4 *)nullptr);
        timer.stop();
        kernel_ms = pnb_last_kernel_ms(2);
    }
    if (out_mem == PNB_DEVICE) {
        PNB_LAUNCH(k_finalize, nblk(npix, 256), 256, 0, s, acc.as<double>(), npix, out);
        PNB_CUDA(cudaStreamSynchronize(s));
    } else {
        DevBuf of(sizeof(float) * npix);
        PNB_LAUNCH(k_finalize, nblk(npix, 256), 256, 0, s, acc.as<double>(), npix, of.as<float>());
        PNB_CUDA(cudaMemcpyAsync(out, of.p, sizeof(float) * npix, cudaMemcpyDeviceToHost, s));
        PNB_CUDA(cudaStreamSynchronize(s));
    }
    stage.stop();
    set_kernel_ms(2, kernel_ms);
}

template <typename F>
int guarded_render(F &&f)
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

int pnb_render_image(int nx, int ny, const pnb_render_arrays *a, double x1, double x2, double y1, double y2,
                     double z_camera, double z0, double smooth_lo, double smooth_hi, double z_lo, double z_hi,
                     double min_smooth, const float *lut, int n_lut, int h_power, double max_d, const double *wrap_x,
                     int n_wrap_x, const double *wrap_y, int n_wrap_y, float *out, int out_mem)
{
    return guarded_render([&] {
        render(false, nx, ny, 1, a, x1, x2, y1, y2, 0.0, 1.0, z_camera, z0, smooth_lo, smooth_hi, z_lo, z_hi, min_smooth,
               lut, n_lut, h_power, max_d, wrap_x, n_wrap_x, wrap_y, n_wrap_y, nullptr, 0, out, out_mem);
    });
}

int pnb_render_healpix(const pnb_render_arrays *a, int nside, const float *lut, int n_lut, int h_power, double max_d,
                       double shell_radius, float *out, int out_mem)
{
    return guarded_render([&] { render_healpix(a, nside, lut, n_lut, h_power, max_d, shell_radius, out, out_mem); });
}

int pnb_render_grid(int nx, int ny, int nz, const pnb_render_arrays *a, double x1, double x2, double y1, double y2,
                    double z1, double z2, double smooth_lo, double smooth_hi, const float *lut, int n_lut, int h_power,
                    double max_d, const double *wrap_x, int n_wrap_x, const double *wrap_y, int n_wrap_y,
                    const double *wrap_z, int n_wrap_z, float *out, int out_mem)
{
    return guarded_render([&] {
        render(true, nx, ny, nz, a, x1, x2, y1, y2, z1, z2, 0.0, 0.0, smooth_lo, smooth_hi, 0.0, 0.0, 0.0, lut, n_lut,
               h_power, max_d, wrap_x, n_wrap_x, wrap_y, n_wrap_y, wrap_z, n_wrap_z, out, out_mem);
    });
}

}  // extern "C"
