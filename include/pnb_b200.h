/* pnb_b200.h -- C ABI of libpnb_b200.so, the B200 (sm_100a) implementation of pynbody's SPH
 * smoothing hot path.  Plain pointers and sizes only; no torch / numpy types.
 *
 * Each entry point names the reference interface it replaces (paths relative to the pynbody
 * source tree).  The reference's native layer is two CPython extension modules:
 *     pynbody.kdtree.kdmain   (method table pynbody/kdtree/kdmain.cpp:63-85)
 *     pynbody.sph._render     (pynbody/sph/_render.pyx:209, :373)
 * INTEGRATION.md shows the ctypes binding a pynbody maintainer would add.
 *
 * Conventions
 *   - every function returns PNB_OK (0) or a negative PNB_ERR_* code; pnb_last_error() gives the
 *     message (thread-local).  The Python host layer maps codes to the exception types the
 *     reference raises (ValueError / TypeError / RuntimeError, kdmain.cpp:136-182,367-372,795-800).
 *   - dtype codes: PNB_F32 = 0, PNB_F64 = 1.
 *   - mem codes:  PNB_HOST = 0 (pageable or pinned host memory), PNB_DEVICE = 1 (device memory on
 *     the current CUDA device, e.g. a torch tensor's data_ptr()).
 *   - strides are in BYTES, as in the numpy buffers the reference borrows (kd.h:163-190).
 *   - all per-particle outputs are written in FILE order (the order of the input arrays), as the
 *     reference does through particleOffsets (kd.h:205-206).
 *   - there is no CPU fallback: every compute entry point fails with PNB_ERR_CUDA if no sm_100
 *     device is usable.
 */
#ifndef PNB_B200_H
#define PNB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNB_OK            0
#define PNB_ERR_VALUE    -1   /* -> ValueError   */
#define PNB_ERR_TYPE     -2   /* -> TypeError    */
#define PNB_ERR_RUNTIME  -3   /* -> RuntimeError (gather overflow, smooth.h:253-267) */
#define PNB_ERR_CUDA     -4   /* -> RuntimeError (CUDA failure / no device) */
#define PNB_ERR_MEMORY   -5   /* -> MemoryError  */
/* Every function that takes a `pnb_tree *` returns PNB_ERR_VALUE for a null handle; pnb_tree_destroy(NULL)
 * is a no-op.  pnb_last_error() returns the message of the last failure on the calling thread. */

#define PNB_F32 0
#define PNB_F64 1
#define PNB_HOST 0
#define PNB_DEVICE 1

/* array slots of kdmain.set_arrayref (kdmain.cpp:511-579) */
#define PNB_ARR_SMOOTH 0
#define PNB_ARR_RHO    1
#define PNB_ARR_MASS   2
#define PNB_ARR_QTY    3
#define PNB_ARR_QTY_SM 4

/* property ids of kdmain.populate (kdmain.cpp:53-60; pynbody/kdtree/__init__.py:70-77) */
#define PNB_PROP_HSM         1
#define PNB_PROP_RHO         2
#define PNB_PROP_QTYMEAN_1D  3
#define PNB_PROP_QTYMEAN_ND  4
#define PNB_PROP_QTYDISP_1D  5
#define PNB_PROP_QTYDISP_ND  6
#define PNB_PROP_QTYDIV      7
#define PNB_PROP_QTYCURL     8

/* kernel ids of sph/kernels.hpp:21-33 */
#define PNB_KERNEL_CUBIC_SPLINE 0
#define PNB_KERNEL_WENDLAND_C2  1

typedef struct pnb_tree pnb_tree;

/* ---- library -------------------------------------------------------------------------------- */
const char *pnb_last_error(void);
const char *pnb_version(void);
/* number of usable CUDA devices (0 if none); never fails */
int pnb_device_count(void);
/* Process-wide tuning switches (no reference counterpart):
 *   "neighbour_lists"  -1 (default) keep each particle's k neighbour ids on the device after the kNN pass when
 *                      they fit in half of the free device memory, so that rho / mean / dispersion / div / curl
 *                      stream over them instead of searching again; 0 never; 1 always.  (env PNB_NEIGHBOUR_LISTS)
 *   "gather_walk"      1: the reductions always use the tree-walking fixed-radius gather.  (env PNB_GATHER_WALK)
 *   "tree_build"       how the device tree is built; results do not depend on it.  2 (default): median-split kd-tree from
 *                      three per-axis sorted lists, one fused partition pass per level, the lowest 6 levels per sub-tree in
 *                      shared memory; 3 and 4: the SAME tree by the per-level formulations it replaced (mark + prefix sum +
 *                      partition / fused pass for every level); 1: the lowest levels by an in-CTA sort per node; 0: one
 *                      Morton-key sort for the upper levels + shared-memory kd levels (cheapest build, slower searches).
 *                      (env PNB_TREE_BUILD)
 *   "streams"          1 (default): every host thread that calls the library works on its own non-blocking CUDA stream
 *                      (ordered after whatever the caller had enqueued on the default stream when the call was made,
 *                      drained before the call returns), so calls from two threads overlap on the device;
 *                      0: everything on the default stream.  (env PNB_STREAMS)
 *   "knn_blocks_per_sm" > 0 caps the resident warps per SM of the (persistent) neighbour-search kernel (2 warps per unit)
 *                      so that kernels launched from another thread find room beside it; 0 (default): as many as fit.
 * PNB_ERR_VALUE for an unknown name. */
int pnb_set_option(const char *name, int value);
/* select the device subsequent calls from this thread use */
int pnb_set_device(int device);

/* ---- tree: kdmain.init + kdmain.build (kdmain.cpp:189-302; kd.cpp:98-238) --------------------
 * Uploads pos (N x 3) and mass (N), builds a balanced median-split kd-tree on the GPU and keeps
 * the particles resident in tree order as structure-of-arrays.  `leafsize` is the reference's
 * user-visible leaf size: it only shapes pnb_tree_node_count / pnb_tree_export; the
 * device tree always uses 32-particle buckets (one warp lane per particle).
 * `boxsize` <= 0 means non-periodic (kdmain.cpp:360-361).  mass may be NULL. */
pnb_tree *pnb_tree_create(const void *pos, int64_t pos_stride0, int64_t pos_stride1,
                          const void *mass, int64_t mass_stride,
                          int64_t n, int dtype, int leafsize, int mem);
/* kdmain.free (kdmain.cpp:304-342) */
void pnb_tree_destroy(pnb_tree *t);

int64_t pnb_tree_num_particles(const pnb_tree *t);
/* tight bounds of all particles: double[6] = min xyz, max xyz (root cell, kd.cpp:120-135) */
int pnb_tree_bounds(const pnb_tree *t, double *out6);
/* kdmain.get_node_count (kd.cpp:98-111): 2 * nSplit for the user's leafsize */
int64_t pnb_tree_node_count(const pnb_tree *t);
/* device-tree order (tree position -> file index), int64[N], host memory.  This is the order of the
 * 32-particle-bucket tree the kernels traverse; see pnb_tree_export for the reference-shaped tree. */
int pnb_tree_order(pnb_tree *t, int64_t *out);
/* What KDTree.serialize() returns (pynbody/kdtree/__init__.py:136-163): reference-layout
 * KDNode[node_count] records (kd.h:29-40, 80 bytes each) and the matching particle_offsets int64[N]
 * (tree order -> file order) of a tree with the reference's shape for the user's leafsize (median
 * splits m=(lo+hi)/2, kd.cpp:98-111,176-199), built on the GPU on demand.  Either output may be NULL.
 * Host memory. */
int pnb_tree_export(pnb_tree *t, void *kdnodes_out, int64_t n_nodes, int64_t *offsets_out);
/* device-tree introspection used by the tests: number of 32-particle buckets, their start offsets
 * (int64[n_buckets + 1]) and tight bounds (double[n_buckets][6] = min xyz, max xyz), host memory */
int64_t pnb_tree_num_buckets(const pnb_tree *t);
int pnb_tree_buckets(pnb_tree *t, int64_t *start, double *bounds);
/* tight bounds of EVERY node of the device tree in heap order (root = 1, children 2i / 2i+1, buckets at
 * n_buckets .. 2 n_buckets - 1; node 0 unused): double[2 * n_buckets][6] = min xyz, max xyz, host memory */
int pnb_tree_node_bounds(pnb_tree *t, double *bounds);

/* ---- kdmain.set_arrayref (kdmain.cpp:511-579) ------------------------------------------------
 * Registers a borrowed array for slot `id`.  ncols is 1 (N) or 3 (N x 3, qty / qty_sm only).
 * smooth/rho/mass must have the tree dtype, qty and qty_sm a common dtype (TypeError otherwise,
 * pynbody/kdtree/__init__.py:262-267). */
int pnb_set_array(pnb_tree *t, int id, void *ptr, int64_t stride0, int64_t stride1,
                  int ncols, int dtype, int mem);

/* ---- kdmain.nn_start + domain_decomposition + populate + nn_stop ------------------------------
 * (kdmain.cpp:344-382, 622-653, 681-806; smooth.h:331-480, 626-918).
 * Runs property `propid` with `k` neighbours (k counts the particle itself) and SPH kernel
 * `kernel_id`, writing the registered output array in file order.  `period` <= 0 disables
 * periodicity; if the particles span more than `period` along any axis periodicity is disabled
 * and *period_ignored (if non-NULL) is set to 1 -- the caller issues the reference's
 * RuntimeWarning (smooth.h:92-102).  k > N -> PNB_ERR_VALUE (kdmain.cpp:367-372); more than
 * k + 500 particles inside 2h -> PNB_ERR_RUNTIME (smooth.h:15,253-267). */
int pnb_populate(pnb_tree *t, int propid, int k, int kernel_id, double period, int *period_ignored);

/* ---- multi-GPU host-layer helpers (no reference counterpart: the reference is single-process) ----
 * pnb_set_query_mask: restrict subsequent pnb_populate / pnb_knn calls to the particles whose mask
 * byte (FILE order, n bytes) is non-zero; output entries of the other particles are left unspecified.
 * NULL removes the restriction.  Used to re-run only the boundary particles after a halo exchange.
 * pnb_mark_in_balls: flags_out[i] (FILE order, n bytes) = 1 iff particle i lies within radii[q] of
 * centres[q] (periodic) for any q < nq: the halo particles another rank's boundary queries need. */
int pnb_set_query_mask(pnb_tree *t, const uint8_t *mask, int mem);
/* The same with two switches for searches done in several passes over complementary subsets: `invert` selects the
 * particles whose mask byte is ZERO; `keep_results` leaves the resident results of earlier passes (smoothing
 * lengths, fused densities, neighbour lists) marked valid -- the caller vouches that the passes together cover
 * every particle it later asks about.  mask == NULL removes the restriction. */
int pnb_set_query_mask2(pnb_tree *t, const uint8_t *mask, int mem, int invert, int keep_results);
/* pnb_mark_near_faces: flags_out[i] (FILE order, n bytes) = 1 iff the k-neighbour ball of particle i MIGHT reach a
 * face of the cuboid [lo3, hi3) (+-inf = no face): the smallest ancestor of its bucket holding >= k particles bounds
 * the ball radius by its diagonal; a bucket further than that from every face cannot.  Particles flagged 0 are
 * guaranteed to find all their k neighbours inside the cuboid -- the interior of a rank's domain, searched while the
 * halo exchange for the flagged ones is under way. */
int pnb_mark_near_faces(pnb_tree *t, int k, const double *lo3, const double *hi3, uint8_t *flags_out, int mem);
int pnb_mark_in_balls(pnb_tree *t, const void *centres, int64_t stride0, int64_t stride1,
                      const void *radii, int64_t rstride, int64_t nq, double period,
                      uint8_t *flags_out, int mem);

/* pnb_route_plan / pnb_route_pack: moving particles to the rank that owns their spatial domain (device memory only).
 * The domains are the leaves of a recursive coordinate bisection given as n_splits nodes: particles with
 * coordinate[split_axis[s]] >= split_plane[s] go to split_child[2s+1], the others to split_child[2s]; a child >= 0
 * is another node, a child < 0 is rank -(child + 1); node 0 is the root (n_splits == 0: everything is rank 0's).
 * plan:  owner_out uint8[n] (device) and counts_out int64[n_ranks] (HOST): particles per destination.
 * pack:  packed_out T[n][4] (device) = rows {x, y, z, m} destination-major -- rank 0's first, in their original
 *        order (a stable multi-split) -- and src_index_out int64[n] (device) = original index of every row; the
 *        ranges [sum(counts[:r]), sum(counts[:r+1])) are what an all-to-all sends to rank r.  mass may be NULL. */
int pnb_route_plan(const void *pos, int64_t stride0, int64_t stride1, int64_t n, int dtype,
                   const int32_t *split_axis, const double *split_plane, const int32_t *split_child, int n_splits,
                   int n_ranks, uint8_t *owner_out, int64_t *counts_out);
int pnb_route_pack(const void *pos, int64_t stride0, int64_t stride1, const void *mass, int64_t mass_stride, int64_t n,
                   int dtype, const uint8_t *owner, int n_ranks, void *packed_out, int64_t *src_index_out);

/* ---- kdmain.nn_next, all particles at once (kdmain.cpp:390-440) -------------------------------
 * Neighbour lists for every particle: idx_out int64[N][k] (file-order ids, unsorted, self
 * included with d2 = 0), d2_out T[N][k], h_out T[N] (may be NULL); rows in file order; host memory. */
int pnb_knn(pnb_tree *t, int k, double period, int64_t *idx_out, void *d2_out, void *h_out,
            int *period_ignored);

/* The same, lazily, as the reference's nn_start / nn_next iteration (kdmain.cpp:344-382, 390-440): pnb_knn_start
 * runs the search once and leaves every particle's neighbour list resident on the device (ids 4 bytes +
 * d2 sizeof(T) bytes per neighbour; the registered smooth array, if any, is overwritten with h as the
 * reference does, smooth.h:429); pnb_knn_rows then copies out the rows of the particles with file index
 * first .. first+count-1 (idx_out int64[count][k], d2_out T[count][k], h_out T[count] or NULL; host memory), so
 * a caller iterating over 10^7 particles never holds more than one chunk.  PNB_ERR_VALUE if no list is
 * resident or the range is outside [0, N]. */
int pnb_knn_start(pnb_tree *t, int k, double period, int *period_ignored);
int pnb_knn_rows(pnb_tree *t, int64_t first, int64_t count, int64_t *idx_out, void *d2_out, void *h_out);

/* ---- nCnt of typed_populate's gather (kdmain.cpp:779: smBallGather(4*hsm*hsm), smooth.h:270-324) -----------
 * Number of particles with d2 <= fl(4 h h) around every particle, for the registered smooth array (file order,
 * int32[N]; k-1, k, or more on exact ties -- SURVEY.md section 7b).  The reference keeps this count internal
 * (it decides the k+500 overflow error, smooth.h:253-267); it is exposed so that parity of the gather's
 * neighbour SET SIZES can be tested.  Always computed by the tree walk, never from the neighbour lists. */
int pnb_gather_counts(pnb_tree *t, double period, int32_t *counts_out, int mem);

/* ---- kdmain.particles_in_sphere (kdmain.cpp:655-679) ------------------------------------------
 * File-order ids of particles with d2 <= r*r of (cx,cy,cz) (centre and radius are rounded to the
 * tree dtype first, as the reference's 'f'/'d' argument parsing does).  Writes at most `cap` ids
 * to out (host memory, ascending) and the total count to *n_found. */
int pnb_ball_query(pnb_tree *t, double cx, double cy, double cz, double r, double period,
                   int64_t *out, int64_t cap, int64_t *n_found);

/* ---- _render.render_image (pynbody/sph/_render.pyx:209-365) -----------------------------------
 * x,y,z share pos_dtype and a common byte stride; sm/qty/mass/rho are 1-D with their own dtype
 * and byte stride (the five fused types of _render.pyx:20-38).  lut holds n_lut float32 kernel
 * samples (kernel.get_samples()), h_power = kernel.h_power, max_d = kernel.max_d.  wrap_x/wrap_y
 * are the periodic offsets (converted to float32 as the reference does).  out: float32[ny][nx]. */
typedef struct pnb_render_arrays {
    const void *x, *y, *z;      int64_t pos_stride; int pos_dtype;
    const void *sm;             int64_t sm_stride;  int sm_dtype;
    const void *qty;            int64_t qty_stride; int qty_dtype;
    const void *mass;           int64_t mass_stride; int mass_dtype;
    const void *rho;            int64_t rho_stride; int rho_dtype;
    int64_t n;
    int mem;                    /* PNB_HOST or PNB_DEVICE, applies to all seven arrays */
} pnb_render_arrays;

int pnb_render_image(int nx, int ny, const pnb_render_arrays *a,
                     double x1, double x2, double y1, double y2, double z_camera, double z0,
                     double smooth_lo, double smooth_hi, double z_lo, double z_hi, double min_smooth,
                     const float *lut, int n_lut, int h_power, double max_d,
                     const double *wrap_x, int n_wrap_x, const double *wrap_y, int n_wrap_y,
                     float *out, int out_mem);

/* ---- _render.to_3d_grid (pynbody/sph/_render.pyx:373-497) -------------------------------------
 * out: float32[nx][ny][nz]. */
int pnb_render_grid(int nx, int ny, int nz, const pnb_render_arrays *a,
                    double x1, double x2, double y1, double y2, double z1, double z2,
                    double smooth_lo, double smooth_hi,
                    const float *lut, int n_lut, int h_power, double max_d,
                    const double *wrap_x, int n_wrap_x, const double *wrap_y, int n_wrap_y,
                    const double *wrap_z, int n_wrap_z,
                    float *out, int out_mem);

/* ---- _render.render_spherical_image_core (pynbody/sph/_render.pyx:51-179, sph/healpix.c) -----------
 * Healpix (RING) image around the origin: projected column per solid angle (h_power == 2) or a thin
 * shell at shell_radius sampled with the 3-D kernel (h_power == 3).  `a->sm` holds the smoothing
 * lengths h.  out: float32[12 * nside * nside]. */
int pnb_render_healpix(const pnb_render_arrays *a, int nside,
                       const float *lut, int n_lut, int h_power, double max_d,
                       double shell_radius, float *out, int out_mem);

/* ---- filt.geometry_selection.selection (pynbody/filt/geometry_selection.pyx:31-103, .hpp:3-119) ----
 * Tree-less sphere / cuboid selection over a C-contiguous [n][3] position array of `dtype`, in host or
 * device memory.  params: (x0,y0,z0,r_max) for PNB_REGION_SPHERE, (x0,y0,z0,x1,y1,z1) for PNB_REGION_CUBE,
 * each cast to `dtype` as the reference casts them; wrap > 0 selects the periodic variant (box size, cast
 * to `dtype`).  out: uint8[n], 1 = inside (strict inequalities, arithmetic in `dtype`, bit-identical to
 * the reference).  PNB_ERR_VALUE for a wrong parameter count or an unknown region. */
#define PNB_REGION_SPHERE 0
#define PNB_REGION_CUBE   1
int pnb_select_region(const void *pos, int64_t n, int dtype, int mem, int region, const double *params,
                      int n_params, double wrap, uint8_t *out, int out_mem);

/* Return the device memory the library keeps cached for re-use (freed work buffers) to the driver. */
void pnb_release_cached_memory(void);

/* ---- timing hooks (bench.py) ------------------------------------------------------------------
 * Device time in milliseconds of the kernels launched by the last pnb_tree_create (stage 0),
 * pnb_populate / pnb_knn (stage 1), render call (stage 2) or pnb_select_region (stage 3) on this thread, measured with CUDA
 * events on the library's stream; also the number of kernel launches of that call. */
double pnb_last_device_ms(int stage);
int64_t pnb_last_launch_count(int stage);
/* Duration (ms, CUDA events on the launching stream) of the last launch of one of the dominant
 * kernels on this thread: 0 = kNN (+fused density) kernel, 1 = gather/reduction kernel,
 * 2 = image/grid render kernel, 3 = region selection kernel.  This is what bench.py's roofline block divides by. */
double pnb_last_kernel_ms(int which);
/* cumulative number of kernel launches issued by this thread through the library */
int64_t pnb_total_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* PNB_B200_H */
