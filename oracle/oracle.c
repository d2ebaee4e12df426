/* TEST INFRASTRUCTURE -- NOT part of the product path.
 *
 * liboracle.so: a plain-C, CPU restatement of pynbody's SPH smoothing hot path, used only
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the parity checker
 * for the CUDA implementation in pynbody_b200/csrc.  The product never links or loads it.
 *
 * Pinning: this restatement is checked (tests/test_oracle.py) against golden vectors
 * produced by the reference's own compiled modules (oracle/_ref, built from the sources in
 * /root/reference by oracle/Makefile; generator: tests/golden/make_golden.py) and against
 * the reference's known-answer kernel tables.
 *
 * What is restated (reference file:line):
 *   kd.cpp:31-238, kd.h:68-203, pq.h:57-204, smooth.h:71-480          -> oracle_kd.inc
 *   smooth.h:626-918, kdmain.cpp:681-806                               -> oracle_sph.inc
 *   sph/kernels.hpp:37-116                                             -> kern_* below
 *   kdmain.cpp:655-679 (particles_in_sphere)                           -> orc_ball
 *   sph/_render.pyx:186-193, 209-365, 373-497                          -> orc_render_image / orc_to_3d_grid
 */
#include <float.h>
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_1_PI 0.31830988618379067154   /* smooth.h:19 */

typedef struct {
    double split;
    double bmin[3], bmax[3];
    int32_t dim;
    int64_t lo, hi;
} kdnode_t;

typedef struct {
    int64_t n, nsplit, nnodes;
    int leafsize, is_f64;
    const float *pos_f32;
    const double *pos_f64;
    kdnode_t *nodes;
    int64_t *order;          /* tree order -> file order ("particleOffsets", kd.h:50) */
} kd_t;

/* smooth.h:71-102: periodicity is dropped when the particle extent exceeds the period. */
static void kd_effective_period(const kd_t *kd, double period, int *periodic)
{
    *periodic = 0;
    if (!(period > 0)) return;
    for (int j = 0; j < 3; ++j)
        if (kd->nodes[1].bmax[j] - kd->nodes[1].bmin[j] > period) return;
    *periodic = 1;
}

/* ---- sph/kernels.hpp:37-116.  The C++ templates mix double literals with T operands; the
 * float variants therefore evaluate in double and round once on assignment.  Both variants
 * below take and return double-converted values with the same intermediate roundings. ---- */
static inline double kern_w_f64(int id, int nsm, double q2)
{
    if (id == 0) {
        double q = sqrt(q2), t = 2.0 - q;
        if (t < 0) return 0;
        if (q2 < 1.0) return 1.0 - 0.75 * t * q2;
        return 0.25 * t * t * t;
    }
    if (q2 > 4.0) return 0;
    if (q2 == 0.0) return (21 / 16.) * (1 - 0.0294 * pow(nsm * 0.01, -0.977));
    double au = sqrt(q2 * 0.25);
    double w = pow(1 - au, 4);
    return (21 / 16.) * w * (1 + 4 * au);
}
static inline float kern_w_f32(int id, int nsm, float q2)
{
    if (id == 0) {
        float q = sqrtf(q2);
        float t = (float)(2.0 - q);
        if (t < 0) return 0;
        if (q2 < 1.0) return (float)(1.0 - 0.75 * t * q2);
        return (float)(0.25 * t * t * t);
    }
    if (q2 > 4.0) return 0;
    if (q2 == 0.0) return (float)((21 / 16.) * (1 - 0.0294 * pow(nsm * 0.01, -0.977)));
    float au = sqrtf((float)(q2 * 0.25));
    float w = (float)pow(1 - au, 4);
    return (float)((21 / 16.) * w * (1 + 4 * au));
}
static inline double kern_g_f64(int id, double q2, double r2)
{
    double q = sqrt(q2), r = sqrt(r2);
    if (id == 0) {
        if (q < 1e-10) return 0.0;
        double g = (q < 1.0) ? -3.0 * q + 2.25 * q2 : -0.75 * (2 - q) * (2 - q);
        return g / r;
    }
    if (r < 1e-24) r = 1e-24;
    if (q < 2.0) return -5.0 * q * (1.0 - 0.5 * q) * (1.0 - 0.5 * q) * (1.0 - 0.5 * q) / r;
    return 0.0;
}
static inline float kern_g_f32(int id, float q2, float r2)
{
    float q = sqrtf(q2), r = sqrtf(r2);
    if (id == 0) {
        if (q < 1e-10) return 0.0f;
        float g = (q < 1.0) ? (float)(-3.0 * q + 2.25 * q2) : (float)(-0.75 * (2 - q) * (2 - q));
        return g / r;
    }
    if (r < 1e-24) r = (float)1e-24;
    if (q < 2.0) return (float)(-5.0 * q * (1.0 - 0.5 * q) * (1.0 - 0.5 * q) * (1.0 - 0.5 * q) / r);
    return 0.0f;
}

/* ---- instantiate the tree-dtype layer ---- */
#define TF float
#define TF_MAX FLT_MAX
#define FN(x) x##_f32
#include "oracle_kd.inc"
#undef TF
#undef TF_MAX
#undef FN

#define TF double
#define TF_MAX DBL_MAX
#define FN(x) x##_f64
#include "oracle_kd.inc"
#undef TF
#undef TF_MAX
#undef FN

/* ---- instantiate the (Tf, Tq) reductions ---- */
#define TF float
#define TF_MAX FLT_MAX
#define FN(x) x##_f32
#define KERN_W(id, k, q2) kern_w_f32(id, k, q2)
#define KERN_G(id, q2, r2) kern_g_f32(id, q2, r2)
#define TQ float
#define FQ(x) x##_f32f32
#include "oracle_sph.inc"
#undef TQ
#undef FQ
#define TQ double
#define FQ(x) x##_f32f64
#include "oracle_sph.inc"
#undef TQ
#undef FQ
#undef TF
#undef TF_MAX
#undef FN
#undef KERN_W
#undef KERN_G

#define TF double
#define TF_MAX DBL_MAX
#define FN(x) x##_f64
#define KERN_W(id, k, q2) kern_w_f64(id, k, q2)
#define KERN_G(id, q2, r2) kern_g_f64(id, q2, r2)
#define TQ float
#define FQ(x) x##_f64f32
#include "oracle_sph.inc"
#undef TQ
#undef FQ
#define TQ double
#define FQ(x) x##_f64f64
#include "oracle_sph.inc"
#undef TQ
#undef FQ
#undef TF
#undef TF_MAX
#undef FN
#undef KERN_W
#undef KERN_G

/* ========================================================================================
 * Exported API (ctypes; see oracle/oracle.py)
 * ====================================================================================== */

/* kd.cpp:98-111 */
static void count_nodes(kd_t *kd)
{
    int64_t n = kd->n, l = 1;
    while (n > kd->leafsize) { n >>= 1; l <<= 1; }
    kd->nsplit = l;
    kd->nnodes = l << 1;
}

void *orc_tree_build(const void *pos, int64_t n, int is_f64, int leafsize)
{
    kd_t *kd = (kd_t *)calloc(1, sizeof(kd_t));
    kd->n = n; kd->leafsize = leafsize; kd->is_f64 = is_f64;
    if (is_f64) kd->pos_f64 = (const double *)pos; else kd->pos_f32 = (const float *)pos;
    count_nodes(kd);
    kd->nodes = (kdnode_t *)calloc((size_t)kd->nnodes, sizeof(kdnode_t));
    kd->order = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    if (is_f64) build_f64(kd); else build_f32(kd);
    return kd;
}
void orc_tree_free(void *h)
{
    kd_t *kd = (kd_t *)h;
    if (!kd) return;
    free(kd->nodes); free(kd->order); free(kd);
}
int64_t orc_tree_nnodes(void *h) { return ((kd_t *)h)->nnodes; }
void orc_tree_export(void *h, double *split, double *bmin, double *bmax, int32_t *dim,
                     int64_t *lo, int64_t *hi, int64_t *order)
{
    kd_t *kd = (kd_t *)h;
    for (int64_t i = 0; i < kd->nnodes; ++i) {
        split[i] = kd->nodes[i].split; dim[i] = kd->nodes[i].dim;
        lo[i] = kd->nodes[i].lo; hi[i] = kd->nodes[i].hi;
        for (int j = 0; j < 3; ++j) { bmin[i * 3 + j] = kd->nodes[i].bmin[j]; bmax[i * 3 + j] = kd->nodes[i].bmax[j]; }
    }
    memcpy(order, kd->order, sizeof(int64_t) * (size_t)kd->n);
}

/* populate('hsm'): returns 1 if periodicity was honoured, 0 if not (or disabled because the
 * particles span more than the period -> the reference raises a RuntimeWarning), -1 if k > N
 * (ValueError, kdmain.cpp:367-372). */
int orc_smooth(void *h, int k, double period, void *sm, int nthreads)
{
    kd_t *kd = (kd_t *)h;
    if (k > kd->n) return -1;
    return kd->is_f64 ? smooth_all_f64(kd, k, period, (double *)sm, nthreads, 0, 0, 0)
                      : smooth_all_f32(kd, k, period, (float *)sm, nthreads, 0, 0, 0);
}
/* nn(): neighbour lists [N,k] (file-order ids, d2) in the order particles are visited. */
int orc_nn(void *h, int k, double period, void *sm, int64_t *nn_idx, void *nn_d2, int64_t *visit)
{
    kd_t *kd = (kd_t *)h;
    if (k > kd->n) return -1;
    return kd->is_f64 ? smooth_all_f64(kd, k, period, (double *)sm, 1, nn_idx, (double *)nn_d2, visit)
                      : smooth_all_f32(kd, k, period, (float *)sm, 1, nn_idx, (float *)nn_d2, visit);
}
/* populate(propid 2..8).  0 ok, -2 gather overflow (RuntimeError), -1 bad arguments. */
int orc_populate(void *h, int propid, int k, double period, int kernel_id,
                 const void *sm, const void *mass, void *rho,
                 const void *qty, void *out, int q_is_f64, int nthreads)
{
    kd_t *kd = (kd_t *)h;
    if (propid < 2 || propid > 8) return -1;
    if (kd->is_f64) {
        if (q_is_f64) return populate_f64f64(kd, propid, k, period, kernel_id, sm, mass, rho, qty, out, nthreads);
        return populate_f64f32(kd, propid, k, period, kernel_id, sm, mass, rho, qty, out, nthreads);
    }
    if (q_is_f64) return populate_f32f64(kd, propid, k, period, kernel_id, sm, mass, rho, qty, out, nthreads);
    return populate_f32f32(kd, propid, k, period, kernel_id, sm, mass, rho, qty, out, nthreads);
}
/* nCnt of the gather pass (kdmain.cpp:779) for every particle, file order */
void orc_gather_counts(void *h, double period, const void *sm, int32_t *counts, int nthreads)
{
    kd_t *kd = (kd_t *)h;
    if (kd->is_f64) gather_counts_f64(kd, period, (const double *)sm, counts, nthreads);
    else gather_counts_f32(kd, period, (const float *)sm, counts, nthreads);
}
/* kdmain.cpp:655-679: centre and radius are parsed in the tree dtype. */
int64_t orc_ball(void *h, double period, double cx, double cy, double cz, double r,
                 int64_t *out, int64_t cap)
{
    kd_t *kd = (kd_t *)h;
    int periodic;
    kd_effective_period(kd, period, &periodic);
    int64_t n;
    if (kd->is_f64) {
        double per[3], c[3] = { cx, cy, cz };
        per[0] = per[1] = per[2] = periodic ? period : DBL_MAX;
        n = ball_gather_f64(kd, per, r * r, c, out, 0, cap);
    } else {
        float per[3], c[3] = { (float)cx, (float)cy, (float)cz }, rf = (float)r;
        per[0] = per[1] = per[2] = periodic ? (float)period : FLT_MAX;
        n = ball_gather_f32(kd, per, rf * rf, c, out, 0, cap);
    }
    for (int64_t i = 0; i < n; ++i) out[i] = kd->order[out[i]];
    return n;
}

/* ========================================================================================
 * Renderer.  _render.pyx:186-193 (LUT lookup), :209-365 (image), :373-497 (grid).
 * Inputs are 1-D strided arrays (strides in elements); x,y,z share a dtype, sm/qty/mass/rho
 * each have their own (the five fused types of _render.pyx:20-38).
 * ====================================================================================== */
typedef struct { const void *p; int64_t stride; int f64; } arr_t;
static inline double ld(const arr_t *a, int64_t i)
{
    return a->f64 ? ((const double *)a->p)[i * a->stride] : (double)((const float *)a->p)[i * a->stride];
}
/* qty*mass/rho with C's promotion rules for the three element types */
static inline double weight_of(const arr_t *q, const arr_t *m, const arr_t *r, int64_t i)
{
    if (!q->f64 && !m->f64) {
        float qm = ((const float *)q->p)[i * q->stride] * ((const float *)m->p)[i * m->stride];
        if (!r->f64) return (double)(qm / ((const float *)r->p)[i * r->stride]);
        return (double)qm / ld(r, i);
    }
    return ld(q, i) * ld(m, i) / ld(r, i);
}
/* x[i] + (float)offset in the element type of x */
static inline double shifted(const arr_t *a, int64_t i, float off)
{
    if (a->f64) return ((const double *)a->p)[i * a->stride] + (double)off;
    return (double)(((const float *)a->p)[i * a->stride] + off);
}
static inline float lut_lookup(double d2, double kmax2, float hk, int ns, const float *lut)
{
    unsigned int idx = (unsigned int)(ns * (d2 / kmax2));
    return idx < (unsigned int)ns ? lut[idx] / hk : 0.0f;
}

void orc_render_image(int nx, int ny, int64_t n,
                      const void *x, const void *y, const void *z, int64_t pos_stride, int pos_f64,
                      const void *sm, int sm_f64, const void *qty, int qty_f64,
                      const void *mass, int mass_f64, const void *rho, int rho_f64,
                      double x1, double x2, double y1, double y2, double z_camera, double z0,
                      double smooth_lo, double smooth_hi, double z_lo, double z_hi, double min_smooth,
                      const float *lut, int ns, int h_power, double max_d,
                      const double *wrap_x, int nwx, const double *wrap_y, int nwy, float *out)
{
    arr_t ax = { x, pos_stride, pos_f64 }, ay = { y, pos_stride, pos_f64 }, az = { z, pos_stride, pos_f64 };
    arr_t asm_ = { sm, 1, sm_f64 }, aq = { qty, 1, qty_f64 }, am = { mass, 1, mass_f64 }, ar = { rho, 1, rho_f64 };
    double pdx = (x2 - x1) / nx, pdy = (y2 - y1) / ny;
    double xs = x1 + pdx / 2, ys = y1 + pdy / 2;
    float per_z_dx = (float)((x2 - x1) / (2 * z_camera)), per_z_dy = (float)((y2 - y1) / (2 * z_camera));
    float mid_x = (float)((x2 + x1) / 2), mid_y = (float)((y2 + y1) / 2);
    int use_z = h_power >= 3;
    memset(out, 0, sizeof(float) * (size_t)nx * ny);

    for (int wi = 0; wi < nwx; ++wi) for (int wj = 0; wj < nwy; ++wj) {
        float ox = (float)wrap_x[wi], oy = (float)wrap_y[wj];
        for (int64_t i = 0; i < n; ++i) {
            double xi = shifted(&ax, i, ox), yi = shifted(&ay, i, oy), zi = ld(&az, i);
            double h = ld(&asm_, i), w = weight_of(&aq, &am, &ar, i);
            if (zi < z_lo || zi > z_hi) continue;
            if (w != w) continue;
            if (z_camera != 0.0) {
                if ((zi > z_camera && z_camera > 0) || (zi < z_camera && z_camera < 0)) continue;
                float dzc = (float)(z_camera - zi);
                x1 = mid_x - per_z_dx * dzc; x2 = mid_x + per_z_dx * dzc;
                y1 = mid_y - per_z_dy * dzc; y2 = mid_y + per_z_dy * dzc;
                pdx = (x2 - x1) / nx; pdy = (y2 - y1) / ny;
                xs = x1 + pdx / 2; ys = y1 + pdy / 2;
            }
            if (h < min_smooth) h = min_smooth;
            if (h < pdx * smooth_lo || h > pdx * smooth_hi) continue;
            if (!((use_z * fabs(zi - z0) < max_d * h) && xi > x1 - 2 * h && xi < x2 + 2 * h
                  && yi > y1 - 2 * h && yi < y2 + 2 * h)) continue;
            float hk = (h_power == 2) ? (float)(h * h) : (float)(h * h * h);
            double kmax2 = (h * h) * (max_d * max_d);
            double dz = (zi - z0) * use_z;
            if (max_d * h / pdx < 1 && max_d * h / pdy < 1) {
                int px = (int)((xi - x1) / pdx), py = (int)((yi - y1) / pdy);
                double cx = pdx * (double)px + xs, cy = pdy * (double)py + ys;
                if (px >= 0 && px < nx && py >= 0 && py < ny) {
                    double ddx = xi - cx, ddy = yi - cy;
                    out[(size_t)py * nx + px] += w * lut_lookup(ddx * ddx + ddy * ddy + dz * dz, kmax2, hk, ns, lut);
                }
            } else {
                int xa = (int)((xi - max_d * h - x1) / pdx), xb = (int)((xi + max_d * h - x1) / pdx);
                int ya = (int)((yi - max_d * h - y1) / pdy), yb = (int)((yi + max_d * h - y1) / pdy);
                if (xa < 0) xa = 0; if (xb > nx) xb = nx;
                if (ya < 0) ya = 0; if (yb > ny) yb = ny;
                for (int px = xa; px < xb; ++px) {
                    double ddx = xi - (pdx * (double)px + xs);
                    for (int py = ya; py < yb; ++py) {
                        double ddy = yi - (pdy * (double)py + ys);
                        out[(size_t)py * nx + px] += w * lut_lookup(ddx * ddx + ddy * ddy + dz * dz, kmax2, hk, ns, lut);
                    }
                }
            }
        }
    }
}

void orc_to_3d_grid(int nx, int ny, int nz, int64_t n,
                    const void *x, const void *y, const void *z, int64_t pos_stride, int pos_f64,
                    const void *sm, int sm_f64, const void *qty, int qty_f64,
                    const void *mass, int mass_f64, const void *rho, int rho_f64,
                    double x1, double x2, double y1, double y2, double z1, double z2,
                    double smooth_lo, double smooth_hi,
                    const float *lut, int ns, int h_power, double max_d,
                    const double *wrap_x, int nwx, const double *wrap_y, int nwy,
                    const double *wrap_z, int nwz, float *out)
{
    arr_t ax = { x, pos_stride, pos_f64 }, ay = { y, pos_stride, pos_f64 }, az = { z, pos_stride, pos_f64 };
    arr_t asm_ = { sm, 1, sm_f64 }, aq = { qty, 1, qty_f64 }, am = { mass, 1, mass_f64 }, ar = { rho, 1, rho_f64 };
    const double pdx = (x2 - x1) / nx, pdy = (y2 - y1) / ny;
    const double pdz = (z2 - z1) / ny;                         /* sic: _render.pyx:391 */
    const double xs = x1 + pdx / 2, ys = y1 + pdy / 2, zs = z1 + pdz / 2;
    memset(out, 0, sizeof(float) * (size_t)nx * ny * nz);

    for (int wi = 0; wi < nwx; ++wi) for (int wj = 0; wj < nwy; ++wj) for (int wk = 0; wk < nwz; ++wk) {
        float ox = (float)wrap_x[wi], oy = (float)wrap_y[wj], oz = (float)wrap_z[wk];
        for (int64_t i = 0; i < n; ++i) {
            double xi = shifted(&ax, i, ox), yi = shifted(&ay, i, oy), zi = shifted(&az, i, oz);
            double h = ld(&asm_, i), w = weight_of(&aq, &am, &ar, i);
            if (h < pdx * smooth_lo || h > pdx * smooth_hi) continue;
            if (!(zi > z1 - 2 * h && zi < z2 + 2 * h && xi > x1 - 2 * h && xi < x2 + 2 * h
                  && yi > y1 - 2 * h && yi < y2 + 2 * h)) continue;
            float hk = (h_power == 2) ? (float)(h * h) : (float)(h * h * h);
            double kmax2 = (h * h) * (max_d * max_d);
            if (max_d * h / pdx < 1 && max_d * h / pdy < 1) {
                int px = (int)((xi - x1) / pdx), py = (int)((yi - y1) / pdy), pz = (int)((zi - z1) / pdz);
                double cx = pdx * (double)px + xs, cy = pdy * (double)py + ys, cz = pdz * (double)pz + zs;
                if (px >= 0 && px < nx && py >= 0 && py < ny && pz >= 0 && pz < nz) {
                    double ddx = xi - cx, ddy = yi - cy, ddz = zi - cz;
                    out[((size_t)px * ny + py) * nz + pz] += w * lut_lookup(ddx * ddx + ddy * ddy + ddz * ddz, kmax2, hk, ns, lut);
                }
            } else {
                int xa = (int)((xi - max_d * h - x1) / pdx), xb = (int)((xi + max_d * h - x1) / pdx);
                int ya = (int)((yi - max_d * h - y1) / pdy), yb = (int)((yi + max_d * h - y1) / pdy);
                int za = (int)((zi - max_d * h - z1) / pdz), zb = (int)((zi + max_d * h - z1) / pdz);
                if (xa < 0) xa = 0; if (xb > nx) xb = nx;
                if (ya < 0) ya = 0; if (yb > ny) yb = ny;
                if (za < 0) za = 0; if (zb > nz) zb = nz;
                for (int px = xa; px < xb; ++px) {
                    double ddx = xi - (pdx * (double)px + xs);
                    for (int py = ya; py < yb; ++py) {
                        double ddy = yi - (pdy * (double)py + ys);
                        for (int pz = za; pz < zb; ++pz) {
                            double ddz = zi - (pdz * (double)pz + zs);
                            out[((size_t)px * ny + py) * nz + pz] += w * lut_lookup(ddx * ddx + ddy * ddy + ddz * ddz, kmax2, hk, ns, lut);
                        }
                    }
                }
            }
        }
    }
}

/* ========================================================================================
 * Healpix renderer.  _render.pyx:51-179 (render_spherical_image_core) with the disc query of
 * sph/healpix.c:24-204 (RING scheme).  Same operation order; libm sin/cos/acos/atan2/sqrt.
 * ====================================================================================== */
#define ORC_PI 3.14159265358979323846

static size_t hp_ring_num_from_z(size_t nside, double z)                 /* healpix.c:24-43 */
{
    size_t nring = 4 * nside, iring;
    double nd = (double)nside;
    if (z > 2.0 / 3.0) iring = (size_t)(nd * sqrt(3.0 * (1.0 - z)) + 0.5);
    else if (z >= -2.0 / 3.0) iring = (size_t)(nd * (2.0 - 1.5 * z) + 0.5);
    else iring = (size_t)(nring - nd * sqrt(3.0 * (1.0 + z)) + 0.5);
    if (iring < 1) iring = 1;
    if (iring > nring) iring = nring;
    return iring;
}
static double hp_z_from_ring_num(size_t nside, size_t iring)             /* healpix.c:46-64 */
{
    size_t nl4 = 4 * nside;
    if (iring <= nside) { double t = iring; return 1.0 - (t * t) / (3.0 * nside * nside); }
    if (iring <= 3 * nside) return (2 * (double)nside - (double)iring) / (1.5 * nside);
    { double t = nl4 - iring; return -1.0 + (t * t) / (3.0 * nside * nside); }
}
static size_t hp_pixel_index(size_t nside, size_t iring, size_t iphi)    /* healpix.c:67-91 */
{
    size_t npix = 12 * nside * nside, ncap = 2 * nside * (nside + 1), nr;
    if (iring <= nside) { nr = iring; return nr * (nr - 1) * 2 + iphi - 1; }
    if (iring <= 3 * nside) return ncap + (iring - nside - 1) * 4 * nside + iphi - 1;
    nr = 4 * nside - iring;
    return npix - nr * (nr - 1) * 2 + iphi - 1 - 4 * nr;
}
static size_t hp_query_disc(size_t nside, double *vec0, double radius, size_t *listpix, double *distpix)
{                                                                        /* healpix.c:94-204 */
    double vecnorm = sqrt(vec0[0] * vec0[0] + vec0[1] * vec0[1] + vec0[2] * vec0[2]);
    vec0[0] /= vecnorm; vec0[1] /= vecnorm; vec0[2] /= vecnorm;
    double z0 = vec0[2], sin_theta0 = sqrt(1 - z0 * z0), phi0 = atan2(vec0[1], vec0[0]), theta0 = acos(z0);
    double thetamin = theta0 - radius, thetamax = theta0 + radius;
    if (thetamin < 0.0) thetamin = 0.0;
    if (thetamax > ORC_PI) thetamax = ORC_PI;
    double zmax = cos(thetamin), zmin = cos(thetamax);
    size_t irmin = hp_ring_num_from_z(nside, zmax), irmax = hp_ring_num_from_z(nside, zmin), count = 0;
    const double twopi = 2.0 * ORC_PI;
    for (size_t iring = irmin; iring <= irmax; iring++) {
        double z = hp_z_from_ring_num(nside, iring), sin_theta = sqrt(1.0 - z * z);
        double cos_dphi = (cos(radius) - z * z0) / (sin_theta * sin_theta0), dphi;
        if (cos_dphi >= 1.0) dphi = 0.0; else if (cos_dphi <= -1.0) dphi = ORC_PI; else dphi = acos(cos_dphi);
        size_t ip_lo, ip_hi, n_in_ring, shift;
        if (iring <= nside) { n_in_ring = 4 * iring; shift = 1; }
        else if (iring <= 3 * nside) { n_in_ring = 4 * nside; shift = 2 - (iring - nside + 1) % 2; }
        else { n_in_ring = 4 * (4 * nside - iring); shift = 1; }
        if (dphi > 0.5 * ORC_PI) { ip_lo = 1; ip_hi = n_in_ring; }
        else {
            double phi_low = phi0 - dphi, phi_high = phi0 + dphi;
            if (phi_low < 0.0) phi_low += twopi;
            if (phi_high >= twopi) phi_high -= twopi;
            double phi_factor = n_in_ring / twopi;
            ip_lo = (size_t)(phi_low * phi_factor + 0.5 * shift);
            ip_hi = (size_t)(phi_high * phi_factor + 0.5 * shift + 1);
        }
        if (ip_hi < ip_lo) ip_hi += n_in_ring;
        if (ip_hi - ip_lo > n_in_ring) { ip_lo = 1; ip_hi = n_in_ring; }
        for (size_t iphi = ip_lo; iphi <= ip_hi; iphi++) {
            size_t ip = iphi;
            if (ip > n_in_ring) ip -= n_in_ring;
            size_t ipix = hp_pixel_index(nside, iring, ip);
            double theta = acos(z), phi = (twopi * ((double)ip - 0.5 * ((double)shift)) / n_in_ring);
            if (phi >= twopi) phi -= twopi;
            double vec[3] = { sin(theta) * cos(phi), sin(theta) * sin(phi), cos(theta) };
            double dot = vec0[0] * vec[0] + vec0[1] * vec[1] + vec0[2] * vec[2];
            if (dot > 1.0) dot = 1.0;
            if (dot < -1.0) dot = -1.0;
            double dist = acos(dot);
            if (dist <= radius) { distpix[count] = dist; listpix[count++] = ipix; }
        }
    }
    return count;
}

void orc_render_healpix(int64_t n, const void *rho, int rho_f64, const void *mass, int mass_f64,
                        const void *qty, int qty_f64, const void *x, const void *y, const void *z, int64_t pos_stride,
                        int pos_f64, const void *h, int h_f64, int nside, const float *lut, int ns, int h_power,
                        double max_d, double shell_radius, float *out)
{
    arr_t ar = { rho, 1, rho_f64 }, am = { mass, 1, mass_f64 }, aq = { qty, 1, qty_f64 }, ah = { h, 1, h_f64 };
    arr_t ax = { x, pos_stride, pos_f64 }, ay = { y, pos_stride, pos_f64 }, az = { z, pos_stride, pos_f64 };
    size_t npix = (size_t)12 * nside * nside;
    /* (a ring whose phi range wraps all the way round lists one pixel twice, healpix.c:166-171: leave room) */
    size_t *index = (size_t *)malloc(sizeof(size_t) * (npix + 8 * (size_t)nside + 8));
    double *angle = (double *)malloc(sizeof(double) * (npix + 8 * (size_t)nside + 8));
    const int projected = h_power == 2;
    const double max2 = max_d * max_d, shell_2 = shell_radius * shell_radius;
    memset(out, 0, sizeof(float) * npix);
    for (int64_t i = 0; i < n; ++i) {
        double qty_i = ld(&aq, i);
        if (qty_i != qty_i) continue;
        double pos_i[3] = { ld(&ax, i), ld(&ay, i), ld(&az, i) };
        double hi = ld(&ah, i);
        double distance2 = pos_i[0] * pos_i[0] + pos_i[1] * pos_i[1] + pos_i[2] * pos_i[2];
        /* smooth_2 = h[i]*h[i] is evaluated in the element type of h (_render.pyx:136) */
        double smooth_2 = ah.f64 ? hi * hi : (double)(((const float *)h)[i] * ((const float *)h)[i]);
        double distance = sqrt(distance2), kernel_max_2 = smooth_2 * max2;
        double angular_size, two_rs_d = 0;
        float h_to_kdim;
        if (projected) {
            angular_size = max_d * hi / distance;
            h_to_kdim = (float)(smooth_2 / distance2);
        } else {
            two_rs_d = 2.0 * shell_radius * distance;
            double cos_alpha = (shell_2 + distance2 - kernel_max_2) / two_rs_d;
            if (cos_alpha >= 1.0) continue;
            else if (cos_alpha <= -1.0) angular_size = ORC_PI;
            else angular_size = acos(cos_alpha);
            h_to_kdim = (float)(smooth_2 * hi);
        }
        size_t np_ = hp_query_disc((size_t)nside, pos_i, angular_size, index, angle);
        for (size_t j = 0; j < np_; ++j) {
            double d2;
            if (projected) { double off = distance * angle[j]; d2 = off * off; }
            else d2 = shell_2 + distance2 - two_rs_d * cos(angle[j]);
            float kern = lut_lookup(d2, kernel_max_2, h_to_kdim, ns, lut);
            /* im[...] += qty_i * kern * mass[i] / rho[i] with C's left-to-right promotion over the fused
             * element types (_render.pyx:176): float until the first double operand appears */
            int isd = aq.f64;
            double vd = isd ? qty_i * (double)kern : 0.0;
            float vf = isd ? 0.0f : (float)qty_i * kern;
            if (isd || am.f64) { vd = (isd ? vd : (double)vf) * ld(&am, i); isd = 1; }
            else vf = vf * ((const float *)mass)[i];
            if (isd || ar.f64) { vd = (isd ? vd : (double)vf) / ld(&ar, i); isd = 1; }
            else vf = vf / ((const float *)rho)[i];
            if (isd) out[index[j]] = (float)((double)out[index[j]] + vd);
            else out[index[j]] += vf;
        }
    }
    free(index); free(angle);
}
