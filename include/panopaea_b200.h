/*
 * panopaea_b200.h -- C ABI of the B200-native grid fluid step.
 *
 * This is the drop-in boundary for ONE hot path of msiglreith/panopaea: the
 * 2D staggered-grid (MAC) fluid step driven by examples/dec_fluid.rs.  The
 * reference has no FFI of its own (it is a pure-Rust crate); every entry point
 * below names the Rust item (file:line under the reference tree) it replaces,
 * and INTEGRATION.md shows the `extern "C"` block + safe wrappers a maintainer
 * adds on the Rust side.
 *
 * Conventions
 *  - plain C types only; every function returns int: 0 = PANO_OK, else a
 *    PANO_ERR_* code.  Nothing throws or aborts across the boundary;
 *    pano_last_error() returns a thread-local human-readable message.
 *  - field handles own DEVICE memory laid out exactly like the reference's
 *    containers (panopaea/src/dec/grid.rs:10, 37-62, 76):
 *       Simplex2 (h,w): row-major (y,x), h*w values
 *       Simplex1 (h,w): ONE flat buffer, vy (h+1,w) first, then vx (h,w+1)
 *                       at element offset w*(h+1)
 *       Simplex0 (h,w): row-major (h+1, w+1)
 *    so upload/download of the flat `view_linear()` slice is a single copy.
 *  - dtype: PANO_F64 (the example's type) or PANO_F32 (the crate is generic;
 *    one reference test runs in f32).  Scalars cross the ABI as double.
 *  - calls are asynchronous on the context's stream unless they return a scalar
 *    or copy to host memory.  One host thread per context.
 *  - there is NO CPU fallback: every compute entry point fails with
 *    PANO_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef PANOPAEA_B200_H
#define PANOPAEA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PANO_API __attribute__((visibility("default")))
#else
#define PANO_API
#endif

enum {
    PANO_OK = 0,
    PANO_ERR_INVALID = 1,        /* null handle, bad enum, bad argument            */
    PANO_ERR_SHAPE = 2,          /* shape/dtype mismatch (the reference panics in ndarray Zip/assign) */
    PANO_ERR_CUDA = 3,           /* CUDA runtime/driver failure, or no device       */
    PANO_ERR_UNIMPLEMENTED = 4,  /* mirrors `unimplemented!()` in the reference     */
    PANO_ERR_TIMEOUT = 5,        /* a device-side wait exceeded its bound           */
    PANO_ERR_COMM = 6            /* multi-GPU setup/exchange failure                */
};

enum { PANO_F64 = 0, PANO_F32 = 1 };
enum { PANO_SIMPLEX0 = 0, PANO_SIMPLEX1 = 1, PANO_SIMPLEX2 = 2 };
/* component selector for rectangle fills on a Simplex1 (split(): dec/grid.rs:48-61) */
enum { PANO_COMP_ALL = 0, PANO_COMP_VY = 1, PANO_COMP_VX = 2 };
/* Preconditioner kinds (trait panopaea/src/pcg.rs:4-6).  Only `()` = identity exists in the reference (pcg.rs:8-12);
 * Jacobi and the multigrid V-cycle are additions behind the same seam, f64 only. */
enum { PANO_PRECOND_IDENTITY = 0, PANO_PRECOND_JACOBI = 1, PANO_PRECOND_MULTIGRID = 2 };

typedef struct pano_ctx pano_ctx;      /* device + stream + scratch                  */
typedef struct pano_field pano_field;  /* device-resident Simplex0/1/2               */
typedef struct pano_mg pano_mg;        /* multigrid preconditioner of one grid        */

/* half-open index rectangle rows [y0,y1) x cols [x0,x1); empty if y0>=y1 or x0>=x1 */
typedef struct pano_rect { int64_t y0, y1, x0, x1; } pano_rect;

/* what the reference only prints (pcg.rs:36, 61) */
typedef struct pano_pcg_info {
    int32_t iterations;      /* index i at the `break` (pcg.rs:60-63); == max_iterations if the loop ran out; -1 on the early-out (pcg.rs:35-38) */
    int32_t applies;         /* number of operator applications                                 */
    double final_residual;   /* last max|r| evaluated (max|b| on the early-out)                 */
    double rhs_max;          /* max|b| (pcg.rs:35)                                              */
} pano_pcg_info;

/* every literal of examples/dec_fluid.rs:27, 43-44, 51-54, 72-73, 95 */
typedef struct pano_step_params {
    double timestep;          /* :43  0.05 */
    double threshold;         /* :44  0.1  */
    int32_t max_iterations;   /* :95  100  */
    int32_t precond;          /* PANO_PRECOND_IDENTITY (:92 `&()`), or one of the other kinds */
    pano_rect inflow;         /* :51-52  rows 5..20, cols 54..64 */
    double inflow_density;    /* :53  1.0  */
    double inflow_vy;         /* :54  20.0 */
    pano_rect obstacle;       /* :72-73, :106-107  rows 70..80, cols 50..70 */
} pano_step_params;

/* ------------------------------------------------------------------ library */
PANO_API const char *pano_version(void);
PANO_API const char *pano_last_error(void);
/* number of CUDA devices visible; 0 (and PANO_OK) when there is no driver */
PANO_API int pano_device_count(int *count);

/* ------------------------------------------------------------------ context */
/* stream: a cudaStream_t to enqueue on, or NULL to create a private one. */
PANO_API int pano_ctx_create(int device, void *stream, pano_ctx **out);
/* Fields of the context that are still alive keep its device state alive: it is released with the last of them, so handles
 * may be freed in any order. */
PANO_API int pano_ctx_destroy(pano_ctx *ctx);
PANO_API int pano_ctx_sync(pano_ctx *ctx);
PANO_API int pano_ctx_stream(pano_ctx *ctx, void **stream);
PANO_API int pano_ctx_num_sms(pano_ctx *ctx, int *n);
/* kernels launched by this library on this context since creation */
PANO_API int pano_ctx_launch_count(pano_ctx *ctx, uint64_t *n);
/* CUDA-event timing on the context's stream (the stream the kernels run on) */
PANO_API int pano_timer_start(pano_ctx *ctx);
PANO_API int pano_timer_stop_ms(pano_ctx *ctx, double *ms);   /* records, synchronises, returns elapsed */
/* Lap timing without host synchronisation between laps: pano_timer_mark records one more event on
 * the stream (up to 4096 between two reads); pano_timer_marks_ms synchronises, writes the elapsed
 * milliseconds between consecutive marks (count - 1 values, at most `cap`) and forgets the marks. */
PANO_API int pano_timer_mark(pano_ctx *ctx);
PANO_API int pano_timer_marks_ms(pano_ctx *ctx, double *ms_out, int cap, int *count);
/* Per-phase device time of pano_fluid_step, measured with CUDA events on the context's
 * stream while option "step_timing" is 1.  phases: 0 inflow fills, 1 advect_all,
 * 2 neg_divergence, 3 CG solve, 4 project.  Returns the accumulated milliseconds and the
 * number of steps accumulated since the last reset (this call synchronises, then resets). */
#define PANO_STEP_PHASES 5
PANO_API int pano_ctx_step_times(pano_ctx *ctx, double ms_out[PANO_STEP_PHASES], int64_t *steps);
/* Diagnostic: with option "cg_profile" = 1 the SM-resident CG kernel accumulates the SM clock cycles CTA 0 spends in
 * each section of the last solve: 0 mailbox poll, 1 search update, 2 stencil, 3/6 CTA reductions, 4/7 grid all-reduces,
 * 5 x/r update + mailbox post. */
PANO_API int pano_ctx_cg_profile(pano_ctx *ctx, int64_t cycles_out[8]);
/* Same option, streaming kernel: cycles every CTA spent in its P1 tile loop ([0, G)) and P2 tile loop ([G, 2G)), G = CTAs. */
PANO_API int pano_ctx_cg_profile_ctas(pano_ctx *ctx, int64_t *cycles_out, int n);
/* tuning knobs: "cg_kernel" 0 auto / 1 generic / 2 TMA streaming (two reductions per iteration, as pcg.rs is written) /
 * 3 SM-resident / 4 SM-resident v1 / 5 one cluster / 6 TMA streaming with ONE reduction per iteration (Chronopoulos-Gear
 * arrangement of the same iteration) / 7 SM-resident with one reduction; "cg_single_reduction" -1 (default: kernel 7 on chip, kernel 6 on the slabs
 * of pano_dist, kernel 2 on one GPU's large grids, each where it measured fastest) / 1 (7 and 6 everywhere) / 0 (3 and 2); "cg_ldcg" 0/1,
 * "cg_zigzag" 0/1, "cg_blocks_per_sm" n, "cg_profile" 0/1, "step_timing" 0/1; streaming kernels: "cg_dynamic" -1 auto (from 24
 * tiles per CTA) / 0 fixed tile lists / 1 claimed tiles, "cg_batch" n (claim unit, 0 auto), "cg_fence" bit 0 fence.acq_rel
 * instead of fence.sc, bit 1 system scope only in CTAs that stored into a peer; multi-GPU: "cg_xflags" 1 halo flags (default) /
 * 0 fenced root exchange, "cg_halo_first" 0/1, "dist_fused_halos" 1 (default with the single-reduction solver: ONE ghost-row
 * exchange per step, everything else recomputed locally or mirrored from inside the solver) / 0 four exchanges, "slab_kernels"
 * 0 marching kernels on the slabs (default) / 1 first-generation kernels; SM-resident kernel: "cg_push" 0 root all-reduce
 * (default) / 1 per-CTA inboxes; advection: "advect_kernel" 0 auto / 1 k_advect / 2 k_advect_march / 3 k_advect_march3 /
 * 4 k_advect_tma (TMA-staged persistent kernel; f64, even sizes, at least 128 x 256), "advect_dynamic" 1 tiles claimed from a
 * counter (default) / 0 round-robin, "advect_ctas" n (cap on the persistent grid), "advect_rows" 2/4/8, "advect_prefetch" 0..3,
 * "advect_minblocks" 3..5 (marching kernel); "fused_wide" 0 / 8 / 16: -div and projection blocks of 256 consecutive columns x that
 * many rows (default 8 from 1024 columns on, else the 32-column blocks); multi-GPU: "dist_overlap" 0 (default) / 1 interior rows
 * advected while the ghost rows are in flight; Grid3d: "cg3_kernel" 0 auto (plane tiles through a cp.async ring for even widths,
 * else the column kernel) / 1 column kernel / 2 plane tiles, "cg3_zc" planes per tile (0 auto).
 * A key that was not set with this call is looked up in the environment as PANO_OPT_<key> before its default applies. */
PANO_API int pano_ctx_set_option(pano_ctx *ctx, const char *key, int64_t value);
PANO_API int pano_ctx_get_option(pano_ctx *ctx, const char *key, int64_t *value);

/* pinned host memory for the caller-owned mirrors of the fields */
PANO_API int pano_host_alloc(size_t bytes, void **out);
PANO_API int pano_host_free(void *p);

/* ------------------------------------------------------------------- fields
 * Manifold2d::new_simplex_0/1/2 (dec/grid.rs:358-371): zero-initialised.
 * Manifold2d::num_elem_0/1/2 (dec/grid.rs:346-356). */
PANO_API int pano_field_new(pano_ctx *ctx, int kind, int dtype, size_t h, size_t w, pano_field **out);
PANO_API int pano_field_free(pano_field *f);
PANO_API int pano_field_num_elem(int kind, size_t h, size_t w, size_t *n);
PANO_API int pano_field_info(const pano_field *f, int *kind, int *dtype, size_t *h, size_t *w, size_t *n);
PANO_API int pano_field_device_ptr(const pano_field *f, void **ptr);
/* host <-> device over the flat view (math/linear_view.rs:5-10); n_elems must equal num_elem */
PANO_API int pano_field_upload(pano_field *f, const void *host, size_t n_elems);
PANO_API int pano_field_download(const pano_field *f, void *host, size_t n_elems);
/* view_linear_mut().fill(v)  (dec_fluid.rs:65-66, 89; pcg.rs:32) */
PANO_API int pano_field_fill(pano_field *f, double value);
/* the example's index loops over a rectangle (dec_fluid.rs:48-57, 70-78, 104-112):
 * comp selects vy / vx / both for a Simplex1 (both: the same (y,x) in each array) */
PANO_API int pano_field_fill_rect(pano_field *f, int comp, pano_rect rect, double value);
/* view_linear_mut().assign(&src.view_linear())  (dec_fluid.rs:62-63, pcg.rs:10, 40, 42) */
PANO_API int pano_field_assign(pano_field *dst, const pano_field *src);
/* O(1) exchange of the device buffers of two same-shaped fields (replaces copy-back :62-63) */
PANO_API int pano_field_swap(pano_field *a, pano_field *b);
/* y = y + alpha * x   (ndarray scaled_add; pcg.rs:55-56, dec_fluid.rs:126) */
PANO_API int pano_field_scaled_add(pano_field *y, double alpha, const pano_field *x);
/* x = x * alpha       (dec_fluid.rs:81-83 with alpha=-1, :116-118 with alpha=timestep) */
PANO_API int pano_field_scale(pano_field *x, double alpha);
/* dst = a + beta * dst  (the search update, pcg.rs:75-77) */
PANO_API int pano_field_xpby(pano_field *dst, const pano_field *a, double beta);
/* LinearViewReal::dot_linear / norm_max (math/linear_view.rs:12-30) */
PANO_API int pano_field_dot(const pano_field *a, const pano_field *b, double *out);
PANO_API int pano_field_norm_max(const pano_field *a, double *out);

/* ------------------------------------------- Manifold2d operators, one to one
 * (dec/manifold.rs:46-83 dispatch; impls in dec/grid.rs) */
PANO_API int pano_hodge_0_primal(pano_field *dual, const pano_field *primal);      /* grid.rs:106-148 */
PANO_API int pano_hodge_2_dual(pano_field *primal, const pano_field *dual);        /* grid.rs:149-191 */
PANO_API int pano_hodge_1_primal(pano_field *dual, const pano_field *primal);      /* grid.rs:206-221 */
PANO_API int pano_hodge_1_dual(pano_field *primal, const pano_field *dual);        /* grid.rs:223-238 */
PANO_API int pano_hodge_2_primal(pano_field *dual, const pano_field *primal);      /* grid.rs:253-255 */
PANO_API int pano_hodge_0_dual(pano_field *primal, const pano_field *dual);        /* grid.rs:257-259 */
PANO_API int pano_derivative_0_primal(pano_field *edges, const pano_field *vertices);  /* grid.rs:274-288 */
PANO_API int pano_derivative_1_primal(pano_field *faces, const pano_field *edges);     /* grid.rs:295-305 */
PANO_API int pano_derivative_0_dual(pano_field *edges, const pano_field *faces);       /* grid.rs:318-334, interior edges only */
PANO_API int pano_derivative_1_dual(pano_field *vertices, const pano_field *edges);    /* grid.rs:308-312: always PANO_ERR_UNIMPLEMENTED */

/* ----------------------------------------------------------- fused hot path */
/* advect (dec_fluid.rs:173-211).  dst must not alias src. */
PANO_API int pano_advect(pano_field *dst, const pano_field *src, double timestep, const pano_field *vel);
/* advect_mac (dec_fluid.rs:213-291).  dst must not alias src or vel. */
PANO_API int pano_advect_mac(pano_field *dst, const pano_field *src, double timestep, const pano_field *vel);
/* both of the above in ONE pass over the grid (self-advection, dec_fluid.rs:59-60):
 * q_dst <- advect(q_src, vel), vel_dst <- advect_mac(vel, vel). */
PANO_API int pano_advect_all(pano_field *q_dst, pano_field *vel_dst, const pano_field *q_src,
                             const pano_field *vel, double timestep);
/* b = -div(vel) with the obstacle's edges treated as zero (dec_fluid.rs:69-83).
 * rhs_max (nullable) receives max|b|; when non-null the call synchronises. */
PANO_API int pano_neg_divergence(pano_field *b, const pano_field *vel, pano_rect obstacle, double *rhs_max);
/* z = A(s): the matrix-free Laplacian closure (dec_fluid.rs:100-119) in one pass */
PANO_API int pano_laplacian_apply(pano_field *z, const pano_field *s, double timestep, pano_rect obstacle);
/* vel += dt * d0_dual(p) on interior edges, then the wall loops (dec_fluid.rs:124-141) */
PANO_API int pano_project(pano_field *vel, const pano_field *pressure, double timestep);

/* ------------------------------------------------------------------ solver
 * pcg::precond_conjugate_gradient (pcg.rs:14-82), argument order kept, with the
 * operator fixed to the dec_fluid Laplacian closure (timestep, obstacle).
 * Caller owns x, b and the three scratch fields, exactly as in the reference.
 * On return x, residual and search hold what the reference leaves in them;
 * `auxiliary` is scratch (contents unspecified).  info (nullable): when non-null
 * the call synchronises and fills it. */
PANO_API int pano_pcg_solve(int precond, pano_field *x, const pano_field *b, int32_t max_iterations,
                            double threshold, pano_field *residual, pano_field *auxiliary,
                            pano_field *search, double timestep, pano_rect obstacle, pano_pcg_info *info);

/* ---------------------------------------------------------- preconditioners
 * Objects with the reference's `Preconditioner<L>::apply(&self, dst, src)` (pcg.rs:4-6) for the dec_fluid operator
 * A = timestep * Laplacian(open faces; walls and `obstacle` closed, examples/dec_fluid.rs:100-119).
 *   pano_jacobi_apply   dst = src / diag(A)   (0 on a cell whose four faces are closed)
 *   pano_mg_*           geometric multigrid V-cycle from a zero guess (symmetric positive definite): 2x2 aggregation,
 *                       damped-Jacobi smoothing; specified bit for bit in DESIGN.md 5b.
 * pano_mg_create builds the level hierarchy once per (grid, timestep, obstacle); the object belongs to `ctx` and is
 * released by pano_mg_destroy or with the context.  pano_pcg_solve / pano_fluid_step accept the kinds above: the
 * identity runs the persistent CG kernels, the others run the loop of pcg.rs:32-80 from the host over the same
 * device primitives (every scalar read back, like the reference's sigma/alpha/beta). */
PANO_API int pano_jacobi_apply(pano_field *dst, const pano_field *src, double timestep, pano_rect obstacle);
PANO_API int pano_mg_create(pano_ctx *ctx, size_t h, size_t w, double timestep, pano_rect obstacle, pano_mg **out);
PANO_API int pano_mg_destroy(pano_mg *mg);
PANO_API int pano_mg_apply(pano_mg *mg, pano_field *dst, const pano_field *src);
/* number of levels, and how many of them run inside the single-CTA tail kernel (nullable) */
PANO_API int pano_mg_levels(const pano_mg *mg, int *levels, int *tail_levels);

/* --------------------------------------------------------------------- step
 * One pass of the example's loop body (dec_fluid.rs:46-141, PNG dump excluded)
 * entirely on the device.  temp / vel_temp / residual / auxiliary / search are
 * the caller-owned scratch fields of dec_fluid.rs:33-41.  The copy-backs of
 * :62-63 are buffer swaps, so after the call `density`/`vel` hold the new state
 * and the scratch contents are unspecified.  info nullable as above. */
PANO_API int pano_fluid_step(const pano_step_params *params, pano_field *density, pano_field *vel,
                             pano_field *pressure, pano_field *temp, pano_field *vel_temp,
                             pano_field *residual, pano_field *auxiliary, pano_field *search,
                             pano_pcg_info *info);

/* The same step for callers that keep the fields in HOST memory, as the Rust
 * crate does: uploads density and vel (flat views), runs pano_fluid_step,
 * downloads density, vel and -- unless `pressure` is NULL -- pressure (the example never reads it
 * between steps, examples/dec_fluid.rs:143-164; the solve starts from zero, pcg.rs:32).  All scratch
 * lives in a workspace the context caches per (h, w).  Host buffers should come from pano_host_alloc. */
PANO_API int pano_fluid_step_host(pano_ctx *ctx, const pano_step_params *params, size_t h, size_t w,
                                  double *density, double *vel, double *pressure, pano_pcg_info *info);

/* u8 transfer of a Simplex2 for the PNG dump (panopaea_utils/src/imgproc.rs:2-5 with the
 * vertical flip of png.rs:12): out[h*w] bytes on the host. */
PANO_API int pano_density_to_u8(const pano_field *density, double lower, double upper, uint8_t *host_out);

/* ---------------------------------------------------------------- multi-GPU
 * The step slab-decomposed along y over the GPUs of one node, one process per GPU.
 * Rank g owns rows [y0, y1) given by pano_slab_range.  All halo traffic and the CG reductions move
 * through peer-mapped device memory (CUDA IPC over NVLink) written from inside the kernels; the host
 * only hands the IPC handles around once.  Collective calls: every rank makes the same calls in the
 * same order.  pano_dist_step is asynchronous; pano_dist_sync waits and reports. */
typedef struct pano_dist pano_dist;
#define PANO_IPC_HANDLE_BYTES 64
PANO_API int pano_slab_range(size_t h, int rank, int nranks, size_t *y0, size_t *y1);
PANO_API int pano_dist_create(pano_ctx *ctx, size_t h, size_t w, int rank, int nranks, const pano_step_params *params,
                              pano_dist **out);
PANO_API int pano_dist_destroy(pano_dist *d);
/* this rank's exchange window: raw device pointer (ranks sharing a process) and IPC handle (one process per GPU) */
PANO_API int pano_dist_window(pano_dist *d, void **ptr, size_t *bytes);
PANO_API int pano_dist_ipc_handle(pano_dist *d, void *handle_out /* PANO_IPC_HANDLE_BYTES */);
/* kind 0: peers = void*[nranks] raw window pointers; kind 1: peers = nranks consecutive IPC handles (own entry ignored) */
PANO_API int pano_dist_connect(pano_dist *d, int kind, const void *peers);
/* cap on the CTAs of the solver kernel (loop-back tests that run several ranks on one GPU); 0 = one per SM.
 * Must be the same on every rank (like the device type): a rank derives its neighbours' CTA counts, whose halo flags
 * it waits for, from its own launch parameters. */
PANO_API int pano_dist_set_max_ctas(pano_dist *d, int max_ctas);
/* owned rows of a field: which = 0 density, 1 vy, 2 vx, 3 pressure; host points at this rank's first row */
PANO_API int pano_dist_upload(pano_dist *d, int which, const double *host_rows);
PANO_API int pano_dist_download(pano_dist *d, int which, double *host_rows, size_t *rows);
PANO_API int pano_dist_step(pano_dist *d);
/* The pressure solve alone (pcg.rs:14-82 on this rank's slab), again on the right-hand side of the
 * last step: BASELINE configs[4], "pressure Poisson solve strong scaling".  Collective, asynchronous. */
PANO_API int pano_dist_solve(pano_dist *d);
PANO_API int pano_dist_sync(pano_dist *d, pano_pcg_info *info);
/* The slab step for callers that keep their rows in HOST memory, as pano_fluid_step_host on one GPU: uploads this rank's rows of
 * density, vy and vx (layout and row counts of pano_dist_upload; pinned buffers), steps, brings them back -- the density while the
 * solver runs, the velocity after the projection -- and synchronises (pano_dist_sync).  Collective. */
PANO_API int pano_dist_step_host(pano_dist *d, double *density_rows, double *vy_rows, double *vx_rows, pano_pcg_info *info);

/* --------------------------------------------------------------------- Grid3d
 * SURVEY.md 8(f) rank 4.  The reference holds two 3-D items and nothing else: the struct
 * `Grid3d { dim: (z, y, x) }` (panopaea/src/domain/grid.rs:17-20, no methods) and the free function
 * `trilinear` (panopaea/src/math/interp.rs:23-36, never called).  The family below is the dec_fluid loop body
 * (examples/dec_fluid.rs:46-141) carried to three dimensions by the same rules, one rule per 2-D line; there is
 * no reference code to match, so DESIGN.md 5c IS the definition (restated by oracle/pano_oracle3.inc and,
 * independently, by oracle/np_oracle3.py).  f64 only.
 *   PANO_CELL3 (d,h,w): cell-centred scalar, row-major (z,y,x), x contiguous            -- Simplex2's role
 *   PANO_FACE3 (d,h,w): ONE flat buffer, vz (d+1,h,w), then vy (d,h+1,w), then vx (d,h,w+1) -- Simplex1's role
 * Handles are pano_field: upload/download/fill/assign/swap/scaled_add/scale/xpby/dot/norm_max above work on the
 * flat view unchanged; the 2-D operators reject them with PANO_ERR_SHAPE. */
enum { PANO_CELL3 = 3, PANO_FACE3 = 4 };
enum { PANO_COMP_VZ = 3 };
/* half-open index box [z0,z1) x [y0,y1) x [x0,x1) */
typedef struct pano_box { int64_t z0, z1, y0, y1, x0, x1; } pano_box;
typedef struct pano_step3_params {
    double timestep;          /* 0.05 */
    double threshold;         /* 0.1  */
    int32_t max_iterations;   /* 100  */
    int32_t precond;          /* PANO_PRECOND_IDENTITY only */
    pano_box inflow;          /* density = inflow_density, vy = inflow_vy on the box before advection */
    double inflow_density;
    double inflow_vy;
    pano_box obstacle;        /* faces vz/vy/vx with index in the box are closed in b and in A (not in the projection) */
} pano_step3_params;
PANO_API int pano_field3_new(pano_ctx *ctx, int kind, size_t d, size_t h, size_t w, pano_field **out);
PANO_API int pano_field3_num_elem(int kind, size_t d, size_t h, size_t w, size_t *n);
PANO_API int pano_field3_dim(const pano_field *f, size_t *d, size_t *h, size_t *w);
/* comp: PANO_COMP_ALL (a cell field, or the same (z,y,x) in all three face arrays) / VZ / VY / VX */
PANO_API int pano_field3_fill_box(pano_field *f, int comp, pano_box box, double value);
/* math::trilinear (interp.rs:23-36) evaluated on the host with the library's own expression (argument order kept) */
PANO_API double pano_trilinear(double a000, double a001, double a010, double a011, double a100, double a101,
                               double a110, double a111, double s, double t, double u);
/* advect / advect_mac / both in one pass, as the 2-D entry points above */
PANO_API int pano_advect3(pano_field *dst, const pano_field *src, double timestep, const pano_field *vel);
PANO_API int pano_advect3_mac(pano_field *dst, const pano_field *src, double timestep, const pano_field *vel);
PANO_API int pano_advect3_all(pano_field *q_dst, pano_field *vel_dst, const pano_field *q_src, const pano_field *vel,
                              double timestep);
PANO_API int pano_neg_divergence3(pano_field *b, const pano_field *vel, pano_box obstacle, double *rhs_max);
/* the 7-point Laplacian: z = A(s), walls and obstacle faces closed */
PANO_API int pano_laplacian3_apply(pano_field *z, const pano_field *s, double timestep, pano_box obstacle);
PANO_API int pano_project3(pano_field *vel, const pano_field *pressure, double timestep);
/* pcg.rs:14-82 with the 7-point closure, ONE persistent kernel; arguments as pano_pcg_solve */
PANO_API int pano_pcg3_solve(int precond, pano_field *x, const pano_field *b, int32_t max_iterations, double threshold,
                             pano_field *residual, pano_field *auxiliary, pano_field *search, double timestep,
                             pano_box obstacle, pano_pcg_info *info);
PANO_API int pano_fluid3_step(const pano_step3_params *params, pano_field *density, pano_field *vel, pano_field *pressure,
                              pano_field *temp, pano_field *vel_temp, pano_field *residual, pano_field *auxiliary,
                              pano_field *search, pano_pcg_info *info);
/* pano_fluid_step_host on a Grid3d: host density (d*h*w) and vel (flat face layout) in and out, the pressure unless NULL */
PANO_API int pano_fluid3_step_host(pano_ctx *ctx, const pano_step3_params *params, size_t d, size_t h, size_t w, double *density,
                                   double *vel, double *pressure, pano_pcg_info *info);

#ifdef __cplusplus
}
#endif
#endif /* PANOPAEA_B200_H */
