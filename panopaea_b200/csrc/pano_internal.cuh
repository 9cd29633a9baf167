// pano_internal.cuh -- shared declarations of the B200 grid-fluid library (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/panopaea_b200.h"

// ---------------------------------------------------------------------------------- errors
void pano_set_error(const char *fmt, ...);

#define PANO_FAIL(code, ...)        \
    do {                            \
        pano_set_error(__VA_ARGS__); \
        return (code);              \
    } while (0)

#define PANO_CUDA(expr)                                                                         \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            pano_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return PANO_ERR_CUDA;                                                               \
        }                                                                                       \
    } while (0)

#define PANO_TRY(expr)             \
    do {                           \
        int _rc = (expr);          \
        if (_rc != PANO_OK) return _rc; \
    } while (0)

// ---------------------------------------------------------------------------------- handles
// Device-side control block of the persistent CG kernel (one per context).
struct PanoCgControl {
    // != 0: a bounded wait expired; everybody leaves.  STICKY: launches clear only what follows this word
    // (pano_cg_control_reset), so that a failure inside an asynchronous step is still there when the host next
    // synchronises (pano_check_device_error); kernels launched meanwhile see it at their first wait and leave at once.
    unsigned int error;
    unsigned int pad_;
    unsigned long long barrier;   // monotonically increasing arrival counter
    int iterations;               // as pano_pcg_info
    int applies;
    double final_residual;
    double rhs_max;
    long long prof[8];            // per-section clock64 totals of CTA 0 (option "cg_profile", SM-resident kernel)
};

constexpr size_t kPanoInboxBytes = (size_t)2 * 192 * 3 * 192 * 16;   // = pano_sm100::kInboxUnits * sizeof(ReduceUnit)

struct pano_field;
// cached device fields of pano_fluid_step_host (per (h, w)) and pano_fluid3_step_host (per (d, h, w))
struct PanoWorkspace {
    size_t dep = 0, h = 0, w = 0;
    pano_field *density = nullptr, *vel = nullptr, *pressure = nullptr;
    pano_field *temp = nullptr, *vel_temp = nullptr, *residual = nullptr, *auxiliary = nullptr, *search = nullptr;
};
struct pano_mg;         // multigrid preconditioner (pano_mg.cu)

struct pano_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 0;
    int cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0;
    uint64_t launches = 0;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    std::vector<cudaEvent_t> marks;              // pano_timer_mark laps
    int marks_used = 0;
    cudaStream_t copy_stream = nullptr;          // second stream of pano_fluid_step_host (downloads under the solve)
    cudaEvent_t ev_advect = nullptr, ev_copy = nullptr;
    // reductions: block partials (device) + final scalars (device, pinned host mirror)
    double *d_partials = nullptr;
    size_t partials_cap = 0;      // in doubles
    double *d_scalars = nullptr;  // 8 doubles
    double *h_scalars = nullptr;  // pinned, 8 doubles
    PanoCgControl *d_cg = nullptr;
    PanoCgControl *h_cg = nullptr;   // pinned
    double *d_mail = nullptr;        // boundary-line mailboxes of the SM-resident CG kernel
    size_t mail_cap = 0;             // in doubles
    double *d_sr_scratch = nullptr;  // two more h x w arrays of the single-reduction CG kernel (second r and s buffers)
    size_t sr_scratch_cap = 0;       // in doubles
    int *d_sr_order = nullptr;       // tile order of the single-reduction CG kernel (pano_cg_sr.cu: build_tile_order)
    size_t sr_order_cap = 0;
    long long sr_order_key[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int sr_order_n = 0, sr_order_lo = 0, sr_order_hi = 0;
    void *d_tparts = nullptr;        // per-(tile, warp) reduction units of the dynamically scheduled streaming CG kernel
    size_t tparts_cap = 0;           // in 16-byte units
    unsigned long long *d_claim = nullptr;   // 4 tile-claim counters (one per phase in flight)
    unsigned int *d_adv_claim = nullptr;     // two tile counters of the TMA advection kernel, used alternately
    unsigned long long adv_epoch = 0;
    void *d_units = nullptr;         // publish+poll all-reduce units of the persistent kernels
    void *d_inbox = nullptr;         // per-CTA inboxes of the push all-reduce (pano_sm100.cuh), kPanoInboxBytes
    unsigned long long launch_epoch = 0;
    // optional per-phase timing of pano_fluid_step ("step_timing" option)
    std::vector<cudaEvent_t> phase_events;   // ring of (PANO_STEP_PHASES + 1) events per slot
    int phase_slots = 0, phase_used = 0;
    double phase_ms[PANO_STEP_PHASES] = {0, 0, 0, 0, 0};
    int64_t phase_steps = 0;
    std::map<std::string, int64_t> options;
    std::map<std::pair<size_t, size_t>, PanoWorkspace *> workspaces;
    std::map<std::pair<size_t, std::pair<size_t, size_t>>, PanoWorkspace *> workspaces3;
    std::vector<pano_mg *> mg_cache;             // preconditioners built on this context (pano_mg_create)
    // handles may be released in any order (garbage-collected hosts): fields keep the context alive
    int64_t live_fields = 0;
    bool destroy_pending = false;
};

struct pano_field {
    pano_ctx *ctx = nullptr;
    int kind = 0, dtype = 0;
    size_t h = 0, w = 0, n = 0;
    size_t dep = 0;   // depth of a Grid3d field (kinds PANO_CELL3 / PANO_FACE3), 0 on a Grid2d
    void *d = nullptr;
};

inline size_t pano_dtype_size(int dtype) { return dtype == PANO_F32 ? 4 : 8; }
inline size_t pano_num_elem(int kind, size_t h, size_t w) {
    switch (kind) {
        case PANO_SIMPLEX0: return (h + 1) * (w + 1);
        case PANO_SIMPLEX1: return (h + 1) * w + h * (w + 1);
        default: return h * w;
    }
}

void pano_ctx_field_born(pano_ctx *ctx);   // a field handle now refers to ctx
int pano_check_field(const pano_field *f, const char *name);
int pano_check_same(const pano_field *a, const pano_field *b, const char *what);
int pano_check_kind(const pano_field *f, int kind, const char *name);
int pano_check_grid(const pano_field *a, const pano_field *b, const char *what);   // same ctx/dtype/(h,w)
int pano_activate(pano_ctx *ctx);                                                  // cudaSetDevice
int pano_after_launch(pano_ctx *ctx, const char *what);                            // cudaGetLastError + counter
int pano_ensure_partials(pano_ctx *ctx, size_t doubles);
int64_t pano_option(pano_ctx *ctx, const char *key, int64_t dflt);

// Slab of a larger grid handed to the streaming CG kernel by the multi-GPU step (pano_dist.cu).
struct PanoCgSlab {
    int row0;                     // array row of the first owned row (number of ghost rows above)
    int rows_total;               // rows of every array, ghosts included
    int gy0, gh;                  // global row of the first owned row, global grid height
    double *up_r, *up_s0, *up_s1; // upper neighbour's ghost row below its slab (peer memory), or null
    double *dn_r, *dn_s0, *dn_s1; // lower neighbour's ghost row above its slab (peer memory), or null
    int rank, nranks;
    unsigned long long xseq_base; // identical on every rank
    void *xunits_local;
    void *xunits_peer[8];
    int max_ctas;                 // 0: one CTA per SM; loop-back tests share one GPU between ranks
};

// The same for the single-reduction kernel (pano_cg_sr.cu): r and s are double-buffered, r has TWO halo rows.
struct PanoCgSrSlab {
    int row0, rows_total, gy0, gh;
    double *up_r[2], *up_s[2];    // upper neighbour's r / s buffers at the ghost row that mirrors my row 0, or null
    double *dn_r[2], *dn_s[2];    // lower neighbour's buffers at the ghost row that mirrors my row h-2 (r) / h-1 (s), or null
    double *dn_x;                 // lower neighbour's x at the ghost row that mirrors my row h-1, or null (then p is exchanged outside)
    int rank, nranks;
    unsigned long long xseq_base;
    void *xunits_local;
    void *xunits_peer[8];
    int max_ctas;
};

// internal (non-ABI) entry points shared between translation units
// the example's rectangle loops index vy[(y,x)] / vx[(y,x)] / d[(y,x)] directly: a rectangle that
// leaves the (rows, cols) grid panics in the reference, so it is an error here too
int pano_check_rect_within(const pano_rect &r, size_t rows, size_t cols, const char *what);
void pano_workspace_free_all(pano_ctx *ctx);
void pano_mg_free_all(pano_ctx *ctx);
int pano_pcg_precond_raw(pano_ctx *ctx, int precond, pano_field *x, const pano_field *b, int max_iterations, double threshold,
                         pano_field *residual, pano_field *auxiliary, pano_field *search, double dt, pano_rect ob, pano_pcg_info *info);
int pano_norm_max_raw(pano_ctx *ctx, int dtype, const void *a, size_t n, double *out);   // max|a[k]| (synchronises)
struct RectI;
int pano_cg_control_reset(pano_ctx *ctx);          // before a CG launch: clear the control block, keep the sticky error word
int pano_check_device_error(pano_ctx *ctx, const char *where);   // after a stream sync that copied d_cg into h_cg
int pano_phase_mark(pano_ctx *ctx, int phase);     // record event #phase of the current step (no-op unless step_timing)
int pano_phase_drain(pano_ctx *ctx);               // synchronise and fold recorded events into phase_ms

// ---------------------------------------------------------------------------------- device helpers
struct RectI {
    int y0, y1, x0, x1;
};
inline RectI pano_clip_rect(pano_rect r, size_t hmax, size_t wmax) {
    RectI o;
    auto clip = [](int64_t v, size_t hi) -> int { return v < 0 ? 0 : (v > (int64_t)hi ? (int)hi : (int)v); };
    o.y0 = clip(r.y0, hmax); o.y1 = clip(r.y1, hmax);
    o.x0 = clip(r.x0, wmax); o.x1 = clip(r.x1, wmax);
    if (o.y1 <= o.y0 || o.x1 <= o.x0) o = RectI{0, 0, 0, 0};
    return o;
}

__device__ __forceinline__ bool in_rect(const RectI &r, int y, int x) {
    return y >= r.y0 && y < r.y1 && x >= r.x0 && x < r.x1;
}

template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <class T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    return v;
}

// Deterministic block reductions (fixed shuffle tree, fixed warp order). `scratch` holds >= 32 T.
// Result valid in every thread.
template <class T>
__device__ __forceinline__ T block_sum(T v, T *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    T t = 0;
    for (int i = 0; i < nw; ++i) t += scratch[i];
    return t;
}
template <class T>
__device__ __forceinline__ T block_max(T v, T *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    T t = 0;
    for (int i = 0; i < nw; ++i) t = scratch[i] > t ? scratch[i] : t;
    return t;
}
