// gather_list.cu -- K3/K4 over the resident neighbour lists: the SPH reductions without a second tree walk.
//
// Replaces, for the usual sim['smooth'] -> sim['rho'] -> v_mean / v_disp / v_div / v_curl sequence, the
// gather loop of typed_populate (pynbody/kdtree/kdmain.cpp:764-790) + smBallGather (smooth.h:270-324) and the
// reductions smDensity / smMeanQty* / smDispQty* / smDivQty / smCurlQty (smooth.h:626-918).
//
// The reference finds the neighbours of every particle again for every quantity (a fixed-radius search of
// radius 2h from the root).  Here the kNN pass has left each particle's k nearest neighbours in HBM
// (Tree::nn_ids, knn.cu), and a reduction is a streaming pass: lane = query particle, one coalesced 128-byte
// line of neighbour ids per slot, one 16/32-byte gather of the neighbour's packed record {x, y, z, m} and, for
// the quantity operators, one of {m/rho, q0, q1, q2}.  The distance is recomputed with the reference's
// arithmetic (query shifted by +-L first, then subtracted; separate roundings) and the reference's acceptance
// test d2 <= fl(4 h h) (smooth.h:310, kdmain.cpp:779) is applied to it.
//
// The list holds the k nearest; the reference's gather set is {d2 <= fl(4 h h)}.  With h = 0.5 sqrt(d2_k) the
// two differ only by particles AT the kernel edge (the k-th neighbour itself when fl(4hh) < d2_k, exact ties
// with it, or a particle within one rounding of it): q = 2 up to 1e-16, where both kernels and both gradients
// vanish to ~1e-48.  Results agree with the tree-walking gather (gather.cu) to rounding; the exact neighbour
// COUNT of the reference's gather is served by the tree walk (pnb_gather_counts).
#include "traverse.cuh"

namespace pnb {

constexpr int GL_WARPS = 8;

enum { GL_RHO = 2, GL_MEAN1 = 3, GL_MEAN3 = 4, GL_DISP1 = 5, GL_DISP3 = 6, GL_DIV = 7, GL_CURL = 8 };

template <int OP> struct GlTraits {
    static constexpr int qcols = (OP == GL_MEAN1 || OP == GL_DISP1) ? 1 : (OP == GL_RHO ? 0 : 3);
    static constexpr int ocols = (OP == GL_MEAN3 || OP == GL_CURL) ? 3 : 1;
    static constexpr bool two_pass = (OP == GL_DISP1 || OP == GL_DISP3);
    static constexpr bool grad = (OP == GL_DIV || OP == GL_CURL);
};

// packed per-particle records, tree order
template <typename T> struct alignas(4 * sizeof(T)) Rec4 { T a, b, c, d; };

__device__ __forceinline__ Rec4<float> load_rec(const Rec4<float> *p)
{
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    Rec4<float> r; r.a = v.x; r.b = v.y; r.c = v.z; r.d = v.w; return r;
}
__device__ __forceinline__ Rec4<double> load_rec(const Rec4<double> *p)
{
    const double2 u = __ldg(reinterpret_cast<const double2 *>(p)), v = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    Rec4<double> r; r.a = u.x; r.b = u.y; r.c = v.x; r.d = v.y; return r;
}

template <typename T>
__global__ void k_pack_pos(int64_t n, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
                           const T *__restrict__ m, Rec4<T> *__restrict__ out)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    Rec4<T> r; r.a = x[p]; r.b = y[p]; r.c = z[p]; r.d = m[p];
    out[p] = r;
}
// {m/rho, q0, q1, q2} in double (the weight m/rho is formed once per particle, not once per pair)
template <typename T, typename Q>
__global__ void k_pack_qty(int64_t n, int qcols, const T *__restrict__ m, const T *__restrict__ rho,
                           const Q *__restrict__ q, Rec4<double> *__restrict__ out)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    Rec4<double> r;
    r.a = (double)m[p] / (double)rho[p];
    r.b = (double)q[p];
    r.c = qcols > 1 ? (double)q[n + p] : 0.0;
    r.d = qcols > 2 ? (double)q[2 * n + p] : 0.0;
    out[p] = r;
}

template <typename T, typename Q>
struct GlArgs {
    int64_t n, nsplit;
    const int32_t *bucket_start;
    const uint8_t *query_mask;
    const uint32_t *nn_ids;
    int k, K2;
    int kernel_id;
    double wendland_self;
    T period;
    const Rec4<T> *pos4;        // {x, y, z, m}
    const Rec4<double> *qty4;   // {m/rho, q0, q1, q2} (ops >= 3)
    const T *smooth_t;
    const Q *qty_t;             // [qcols][n] (own quantity of the query, gradient ops)
    void *out_t;
};

// separation along one axis as the reference forms it: the query is shifted by +-L first (kd.h:78-103), then
// the neighbour is subtracted
template <typename T, bool PERIODIC>
__device__ __forceinline__ T image_delta(T q, T x, T L, T half)
{
    T d = Ex<T>::sub(q, x);
    if (PERIODIC) {
        if (d > half) d = Ex<T>::sub(Ex<T>::sub(q, L), x);
        else if (d < -half) d = Ex<T>::sub(Ex<T>::add(q, L), x);
    }
    return d;
}

// one (query, neighbour) term of the reduction; PASS 1 = second sweep of the dispersion operators
template <typename T, typename Q, int OP, bool PERIODIC, int PASS>
__device__ __forceinline__ void gl_term(const GlArgs<T, Q> &a, const Rec4<T> &pj, const Rec4<double> &qj, T qx, T qy, T qz,
                                        T L, T half, T ball2, T ih2, T norm, const double (&own)[3], const Q (&mean)[3],
                                        double (&acc)[3])
{
    using OT = GlTraits<OP>;
    const T dx = image_delta<T, PERIODIC>(qx, pj.a, L, half), dy = image_delta<T, PERIODIC>(qy, pj.b, L, half),
            dz = image_delta<T, PERIODIC>(qz, pj.c, L, half);
    const T d2 = Ex<T>::add(Ex<T>::add(Ex<T>::mul(dx, dx), Ex<T>::mul(dy, dy)), Ex<T>::mul(dz, dz));
    if (!(d2 <= ball2)) return;                                     // smooth.h:310 (non-strict)
    if (OP == GL_RHO) {
        // rho_i += (W * norm) * m_j, each term rounded as the reference rounds it (smooth.h:640-645)
        T w = (T)kern_w(a.kernel_id, a.wendland_self, d2 * ih2);
        w *= norm;
        acc[0] += (double)(w * pj.d);
    } else if (!OT::grad) {
        T w = (T)kern_w(a.kernel_id, a.wendland_self, d2 * ih2);
        w *= norm;
        const double wm = (double)w * qj.a;
        if (PASS == 0) {
            acc[0] += (double)(Q)(wm * qj.b);
            if (OT::qcols > 1) { acc[1] += (double)(Q)(wm * qj.c); acc[2] += (double)(Q)(wm * qj.d); }
        } else {                                                    // squared deviations from the mean (smooth.h:814-918)
            Q tq = mean[0] - (Q)qj.b;
            acc[0] += (double)(Q)(wm * (double)tq * (double)tq);
            if (OT::qcols > 1) {
                tq = mean[1] - (Q)qj.c;
                acc[0] += (double)(Q)(wm * (double)tq * (double)tq);
                tq = mean[2] - (Q)qj.d;
                acc[0] += (double)(Q)(wm * (double)tq * (double)tq);
            }
        }
    } else {
        // raw UNWRAPPED separation (smooth.h:736-738, :791-793), wrapped r2 in the kernel gradient
        const T ux = qx - pj.a, uy = qy - pj.b, uz = qz - pj.c;
        T g = (T)kern_g(a.kernel_id, d2 * ih2, d2);
        g *= norm;
        const T dq0 = (T)(qj.b - own[0]), dq1 = (T)(qj.c - own[1]), dq2 = (T)(qj.d - own[2]);
        const double gm = (double)g * qj.a;
        if (OP == GL_DIV) {
            const T dv = ux * dq0 + uy * dq1 + uz * dq2;
            acc[0] += (double)(Q)(gm * (double)dv);
        } else {
            const T c0 = uy * dq2 - uz * dq1, c1 = uz * dq0 - ux * dq2, c2 = ux * dq1 - uy * dq0;
            acc[0] += (double)(Q)(gm * (double)c0);
            acc[1] += (double)(Q)(gm * (double)c1);
            acc[2] += (double)(Q)(gm * (double)c2);
        }
    }
}

// one sweep over the lane's neighbour list, U list slots at a time: the U id loads, then the U (+U) record gathers are
// issued before any of them is consumed, so that each warp keeps several L1/L2 round trips in flight (the kernel is
// bound by the latency of these gathers, not by bytes or issue slots: profiles/r2_gather_list_summary.txt)
template <typename T, typename Q, int OP, bool PERIODIC, int PASS>
__device__ __forceinline__ void gl_sweep(const GlArgs<T, Q> &a, const uint32_t *ids, uint32_t self, T qx, T qy, T qz,
                                         T L, T half, T ball2, T ih2, T norm, const double (&own)[3], const Q (&mean)[3],
                                         double (&acc)[3])
{
    constexpr int U = OP == GL_RHO ? 4 : 2;
    for (int e0 = 0; e0 < a.k; e0 += U) {
        uint32_t j[U];
        Rec4<T> pj[U];
        Rec4<double> qj[U];
#pragma unroll
        for (int u = 0; u < U; ++u) j[u] = e0 + u < a.k ? ids[(e0 + u) * 32] : self;     // (filler slots are skipped below)
#pragma unroll
        for (int u = 0; u < U; ++u) pj[u] = load_rec(a.pos4 + j[u]);
        if (OP != GL_RHO) {
#pragma unroll
            for (int u = 0; u < U; ++u) qj[u] = load_rec(a.qty4 + j[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (e0 + u < a.k)
                gl_term<T, Q, OP, PERIODIC, PASS>(a, pj[u], qj[u], qx, qy, qz, L, half, ball2, ih2, norm, own, mean, acc);
    }
}

template <typename T, typename Q, int OP, bool PERIODIC>
__global__ void __launch_bounds__(GL_WARPS * 32, 4) k_gather_list(GlArgs<T, Q> a)
{
    using OT = GlTraits<OP>;
    const int lane = threadIdx.x & 31;
    const int64_t bucket = (int64_t)blockIdx.x * GL_WARPS + (threadIdx.x >> 5);
    if (bucket >= a.nsplit) return;
    const int start = a.bucket_start[bucket];
    const int cnt = a.bucket_start[bucket + 1] - start;
    const int64_t p = start + lane;
    if (!(lane < cnt && (!a.query_mask || a.query_mask[p]))) return;

    const Rec4<T> me = load_rec(a.pos4 + p);
    const T qx = me.a, qy = me.b, qz = me.c;
    const T h = a.smooth_t[p];
    const T ball2 = Ex<T>::mul(Ex<T>::mul((T)4, h), h);              // kdmain.cpp:779
    const T ih = (T)(1.0 / (double)h);
    const T ih2 = ih * ih;
    // 1/(pi h^3) for values, 1/(pi h^4) for gradients (smooth.h:634-636, :724-726)
    const T norm = OT::grad ? (T)(0.31830988618379067154 * (double)ih2 * (double)ih2)
                            : (T)(0.31830988618379067154 * (double)ih * (double)ih2);
    const T L = a.period, half = (T)0.5 * a.period;
    double own[3] = {0, 0, 0};
    if (OT::grad)
        for (int c = 0; c < 3; ++c) own[c] = (double)(T)a.qty_t[(int64_t)c * a.n + p];   // (T) cast: smooth.h:727-733
    const uint32_t *ids = a.nn_ids + (size_t)bucket * a.K2 * 32 + lane;

    double acc[3] = {0, 0, 0};
    Q mean[3] = {0, 0, 0};
    gl_sweep<T, Q, OP, PERIODIC, 0>(a, ids, (uint32_t)p, qx, qy, qz, L, half, ball2, ih2, norm, own, mean, acc);
    if (OT::two_pass) {
        mean[0] = (Q)acc[0]; mean[1] = (Q)acc[1]; mean[2] = (Q)acc[2];
        acc[0] = acc[1] = acc[2] = 0;
        gl_sweep<T, Q, OP, PERIODIC, 1>(a, ids, (uint32_t)p, qx, qy, qz, L, half, ball2, ih2, norm, own, mean, acc);
        reinterpret_cast<Q *>(a.out_t)[p] = (Q)sqrt(acc[0]);            // smooth.h:866, :917
        return;
    }
    if (OP == GL_RHO) {
        reinterpret_cast<T *>(a.out_t)[p] = (T)acc[0];
    } else {
#pragma unroll
        for (int c = 0; c < OT::ocols; ++c) reinterpret_cast<Q *>(a.out_t)[(int64_t)c * a.n + p] = (Q)acc[c];
    }
}

template <typename T, typename Q, int OP>
static void launch_gl(Tree &t, GlArgs<T, Q> &a, bool periodic)
{
    const unsigned grid = (unsigned)((t.nsplit + GL_WARPS - 1) / GL_WARPS);
    KernelTimer timer(1, t.stream);
    if (periodic) PNB_LAUNCH((k_gather_list<T, Q, OP, true>), grid, GL_WARPS * 32, 0, t.stream, a);
    else PNB_LAUNCH((k_gather_list<T, Q, OP, false>), grid, GL_WARPS * 32, 0, t.stream, a);
    timer.stop();
}

template <typename T>
static void ensure_pos4(Tree &t)
{
    if (t.pos4.p) return;
    t.pos4.alloc(sizeof(Rec4<T>) * (size_t)t.n);
    PNB_LAUNCH((k_pack_pos<T>), (unsigned)((t.n + 255) / 256), 256, 0, t.stream, t.n, t.x.as<T>(), t.y.as<T>(),
               t.z.as<T>(), t.mass.as<T>(), t.pos4.as<Rec4<T>>());
}

template <typename T, typename Q>
static void run_gl_typed(Tree &t, int propid, int k, double period, const KernelParams &kp, const void *smooth_tree,
                         const void *rho_tree, const void *qty_tree, void *out_tree)
{
    ensure_pos4<T>(t);
    GlArgs<T, Q> a;
    a.n = t.n; a.nsplit = t.nsplit;
    a.bucket_start = t.bucket_start.as<int32_t>();
    a.query_mask = t.query_mask.p ? t.query_mask.as<uint8_t>() : nullptr;
    a.nn_ids = t.nn_ids.as<uint32_t>();
    a.k = k; a.K2 = t.nn_K2;
    a.kernel_id = kp.kernel_id;
    a.wendland_self = kp.wendland_self;
    a.period = period > 0 ? (T)period : (T)0;
    a.pos4 = t.pos4.as<Rec4<T>>();
    a.qty4 = nullptr;
    a.smooth_t = (const T *)smooth_tree;
    a.qty_t = (const Q *)qty_tree;
    a.out_t = out_tree;
    DevBuf q4;
    if (propid != GL_RHO) {
        const int qcols = (propid == GL_MEAN1 || propid == GL_DISP1) ? 1 : 3;
        q4.alloc(sizeof(Rec4<double>) * (size_t)t.n);
        PNB_LAUNCH((k_pack_qty<T, Q>), (unsigned)((t.n + 255) / 256), 256, 0, t.stream, t.n, qcols, t.mass.as<T>(),
                   (const T *)rho_tree, (const Q *)qty_tree, q4.as<Rec4<double>>());
        a.qty4 = q4.as<Rec4<double>>();
    }
    const bool per = period > 0;
    switch (propid) {
    case GL_RHO: launch_gl<T, Q, GL_RHO>(t, a, per); break;
    case GL_MEAN1: launch_gl<T, Q, GL_MEAN1>(t, a, per); break;
    case GL_MEAN3: launch_gl<T, Q, GL_MEAN3>(t, a, per); break;
    case GL_DISP1: launch_gl<T, Q, GL_DISP1>(t, a, per); break;
    case GL_DISP3: launch_gl<T, Q, GL_DISP3>(t, a, per); break;
    case GL_DIV: launch_gl<T, Q, GL_DIV>(t, a, per); break;
    case GL_CURL: launch_gl<T, Q, GL_CURL>(t, a, per); break;
    default: throw Error(PNB_ERR_VALUE, "unknown smoothing property id");
    }
    PNB_CUDA(cudaStreamSynchronize(t.stream));        // q4 goes back to the block cache on scope exit
}

void run_gather_lists(Tree &t, int propid, int k, double period, const KernelParams &kp, const void *smooth_tree,
                      const void *rho_tree, const void *qty_tree, int qdtype, void *out_tree)
{
    if (!t.has_mass) throw Error(PNB_ERR_VALUE, "mass array has not been set");
    if (!t.nn_valid || t.nn_k != k) throw Error(PNB_ERR_RUNTIME, "neighbour lists are not available for this pass");
    if (t.dtype == PNB_F64) {
        if (qdtype == PNB_F64 || propid == GL_RHO)
            run_gl_typed<double, double>(t, propid, k, period, kp, smooth_tree, rho_tree, qty_tree, out_tree);
        else run_gl_typed<double, float>(t, propid, k, period, kp, smooth_tree, rho_tree, qty_tree, out_tree);
    } else {
        if (qdtype == PNB_F32 || propid == GL_RHO)
            run_gl_typed<float, float>(t, propid, k, period, kp, smooth_tree, rho_tree, qty_tree, out_tree);
        else run_gl_typed<float, double>(t, propid, k, period, kp, smooth_tree, rho_tree, qty_tree, out_tree);
    }
}

}  // namespace pnb
