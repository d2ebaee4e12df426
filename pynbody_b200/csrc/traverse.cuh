// traverse.cuh -- device helpers shared by the kNN, gather and ball-query kernels.
//
// Arithmetic contract (hard part (a)): the reference computes
//     d2 = ((sx-x)*(sx-x) + (sy-y)*(sy-y)) + (sz-z)*(sz-z)
// in the tree dtype with one rounding per operation (gcc -O2, x86-64 SSE2, no FMA;
// pynbody/kdtree/smooth.h:204-206,226-229) where (sx,sy,sz) is the query shifted by +-L per axis as
// chosen by the cell test (pynbody/kdtree/kd.h:68-154).  Everything that decides neighbour-set
// membership below uses the explicitly rounded intrinsics so nvcc can never contract to FMA.
#pragma once
#include "common.cuh"
#include <cfloat>

namespace pnb {

constexpr unsigned FULL = 0xffffffffu;

template <typename T> struct Ex;
template <> struct Ex<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float slack() { return 1.0f + 1e-5f; }
};
template <> struct Ex<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
    static __device__ __forceinline__ double slack() { return 1.0 + 1e-12; }
};

template <typename T>
__device__ __forceinline__ T dist2(T sx, T sy, T sz, T x, T y, T z)
{
    T dx = Ex<T>::sub(sx, x), dy = Ex<T>::sub(sy, y), dz = Ex<T>::sub(sz, z);
    return Ex<T>::add(Ex<T>::add(Ex<T>::mul(dx, dx), Ex<T>::mul(dy, dy)), Ex<T>::mul(dz, dz));
}

// Separation along one axis with the query image chosen PER PARTICLE: the nearest of q, fl(q+L), fl(q-L).
// Same arithmetic as the per-cell shift wherever the cell is small against the box; used for cells that
// span more than half the period (tiny particle sets) where no single shift serves the whole cell.
template <typename T>
__device__ __forceinline__ T nearest_image_delta(T q, T x, T L)
{
    T d = Ex<T>::sub(q, x);
    T dp = Ex<T>::sub(Ex<T>::add(q, L), x), dm = Ex<T>::sub(Ex<T>::sub(q, L), x);
    if (fabs(dp) < fabs(d)) d = dp;
    if (fabs(dm) < fabs(d)) d = dm;
    return d;
}
template <typename T>
__device__ __forceinline__ T dist2_nearest(T qx, T qy, T qz, T x, T y, T z, T L)
{
    T dx = nearest_image_delta<T>(qx, x, L), dy = nearest_image_delta<T>(qy, y, L), dz = nearest_image_delta<T>(qz, z, L);
    return Ex<T>::add(Ex<T>::add(Ex<T>::mul(dx, dx), Ex<T>::mul(dy, dy)), Ex<T>::mul(dz, dz));
}

// read-only view of the device tree
template <typename T>
struct TreeView {
    int64_t n, nsplit;
    const T *x, *y, *z, *mass;
    const T *box;                 // [2*nsplit][6]
    const int32_t *bucket_start;  // [nsplit+1]
    const uint32_t *order;        // tree -> file
    const uint8_t *query_mask;    // tree order; null = every particle is a query (pnb_set_query_mask)
    T period;                     // valid when periodic
    int periodic;
};

template <typename T>
inline TreeView<T> make_view(const Tree &t, double period)
{
    TreeView<T> v;
    v.n = t.n; v.nsplit = t.nsplit;
    v.x = t.x.as<T>(); v.y = t.y.as<T>(); v.z = t.z.as<T>();
    v.mass = t.has_mass ? t.mass.as<T>() : nullptr;
    v.box = t.box.as<T>();
    v.bucket_start = t.bucket_start.as<int32_t>();
    v.order = t.order.as<uint32_t>();
    v.query_mask = t.query_mask.p ? t.query_mask.as<uint8_t>() : nullptr;
    v.periodic = period > 0;
    v.period = v.periodic ? (T)period : (T)0;
    return v;
}

// bounds of one node; every lane of the warp reads the same address (one broadcast transaction)
template <typename T> struct Box6 { T mn[3], mx[3]; };
__device__ __forceinline__ Box6<double> load_box(const double *box, int64_t node)
{
    const double2 *p = reinterpret_cast<const double2 *>(box + node * 6);
    double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    Box6<double> r;
    r.mn[0] = a.x; r.mn[1] = a.y; r.mn[2] = b.x; r.mx[0] = b.y; r.mx[1] = c.x; r.mx[2] = c.y;
    return r;
}
__device__ __forceinline__ Box6<float> load_box(const float *box, int64_t node)
{
    const float2 *p = reinterpret_cast<const float2 *>(box + node * 6);
    float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    Box6<float> r;
    r.mn[0] = a.x; r.mn[1] = a.y; r.mn[2] = b.x; r.mx[0] = b.y; r.mx[1] = c.x; r.mx[2] = c.y;
    return r;
}

// One axis of the periodic cell test (kd.h:78-103): squared gap between query coordinate q and the
// interval [mn,mx], taking the nearer of the direct and the wrapped image; s receives the query
// coordinate shifted accordingly.
template <typename T, bool PERIODIC>
__device__ __forceinline__ T axis_gap2(T q, T mn, T mx, T L, T &s)
{
    T below = mn - q;      // > 0: cell wholly above the query
    T above = q - mx;      // > 0: cell wholly below the query
    T g = 0;
    s = q;
    if (below > 0) {
        g = below;
        if (PERIODIC) { T w = above + L; if (w < below) { g = w; s = q + L; } }
    } else if (above > 0) {
        g = above;
        if (PERIODIC) { T w = below + L; if (w < above) { g = w; s = q - L; } }
    }
    return g * g;
}

// lower bound on the squared distance from (qx,qy,qz) to any particle of the node, and the shifted
// query to use for its particles
template <typename T, bool PERIODIC>
__device__ __forceinline__ T box_gap2(const Box6<T> &b, T qx, T qy, T qz, T L, T &sx, T &sy, T &sz)
{
    T g = axis_gap2<T, PERIODIC>(qx, b.mn[0], b.mx[0], L, sx);
    g += axis_gap2<T, PERIODIC>(qy, b.mn[1], b.mx[1], L, sy);
    g += axis_gap2<T, PERIODIC>(qz, b.mn[2], b.mx[2], L, sz);
    return g;
}

// The per-cell shift (sx,sy,sz) is the nearest image for EVERY particle of the cell iff no particle of the
// cell is further than half a period from it along any axis.  Otherwise (tiny particle sets, or balls
// approaching half the box) the image has to be chosen per particle.
template <typename T, bool PERIODIC>
__device__ __forceinline__ bool shift_is_ambiguous(const Box6<T> &b, T sx, T sy, T sz, T L)
{
    if (!PERIODIC) return false;
    const T half = (T)0.5 * L;
    return fmax(sx - b.mn[0], b.mx[0] - sx) > half || fmax(sy - b.mn[1], b.mx[1] - sy) > half
           || fmax(sz - b.mn[2], b.mx[2] - sz) > half;
}

// in-order successor in the implicit tree when the sub-tree under `cp` is finished (kd.h:17-23)
__device__ __forceinline__ int64_t next_node(int64_t cp)
{
    while ((cp & 1) && cp != 1) cp >>= 1;
    if (cp != 1) ++cp;
    return cp;
}

// ---- SPH kernels (pynbody/sph/kernels.hpp:37-116).  The float instantiation of the reference
// evaluates in double and rounds on assignment; evaluating in double for both dtypes matches that
// to rounding.  Argument is q2 = r2/h2.
__device__ __forceinline__ double kern_w(int kernel_id, double wendland_self, double q2)
{
    if (kernel_id == 0) {
        double q = sqrt(q2), t = 2.0 - q;
        if (t < 0) return 0.0;
        if (q2 < 1.0) return 1.0 - 0.75 * t * q2;
        return 0.25 * t * t * t;
    }
    if (q2 > 4.0) return 0.0;
    if (q2 == 0.0) return wendland_self;
    double au = sqrt(q2 * 0.25);
    double o = 1.0 - au;
    double w = (o * o) * (o * o);
    return (21.0 / 16.0) * w * (1.0 + 4.0 * au);
}
// (dW/dq)/r  (kernels.hpp:54-67, :96-115)
__device__ __forceinline__ double kern_g(int kernel_id, double q2, double r2)
{
    double q = sqrt(q2), r = sqrt(r2);
    if (kernel_id == 0) {
        if (q < 1e-10) return 0.0;
        double g = (q < 1.0) ? -3.0 * q + 2.25 * q2 : -0.75 * (2.0 - q) * (2.0 - q);
        return g / r;
    }
    if (r < 1e-24) r = 1e-24;
    if (q < 2.0) { double o = 1.0 - 0.5 * q; return -5.0 * q * o * o * o / r; }
    return 0.0;
}

// The same kernels for float32 trees.  The reference's float instantiation takes its square roots in float and
// mixes float and double in the polynomials (kernels.hpp:44-67, 91-115, T = float); evaluating in float agrees
// with it to float rounding (bar 1e-4) at a third of the instructions of the double version.
__device__ __forceinline__ float kern_w(int kernel_id, double wendland_self, float q2)
{
    if (kernel_id == 0) {
        const float q = sqrtf(q2), t = 2.0f - q;
        if (t < 0.0f) return 0.0f;
        if (q2 < 1.0f) return 1.0f - 0.75f * t * q2;
        return 0.25f * t * t * t;
    }
    if (q2 > 4.0f) return 0.0f;
    if (q2 == 0.0f) return (float)wendland_self;
    const float au = sqrtf(q2 * 0.25f);
    const float o = 1.0f - au;
    const float w = (o * o) * (o * o);
    return (21.0f / 16.0f) * w * (1.0f + 4.0f * au);
}
__device__ __forceinline__ float kern_g(int kernel_id, float q2, float r2)
{
    const float q = sqrtf(q2);
    float r = sqrtf(r2);
    if (kernel_id == 0) {
        if (q < 1e-10f) return 0.0f;
        const float g = (q < 1.0f) ? -3.0f * q + 2.25f * q2 : -0.75f * (2.0f - q) * (2.0f - q);
        return g / r;
    }
    if (r < 1e-24f) r = 1e-24f;
    if (q < 2.0f) { const float o = 1.0f - 0.5f * q; return -5.0f * q * o * o * o / r; }
    return 0.0f;
}

}  // namespace pnb
