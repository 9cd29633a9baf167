// pano_abi.cu -- context, field handles, copies: the plumbing half of the C ABI
// (include/panopaea_b200.h).  Compute entry points live in pano_prim.cu,
// pano_fused.cu, pano_cg.cu and pano_step.cu.
#include <cstddef>

#include "pano_internal.cuh"

static thread_local char g_err[1024] = "";

void pano_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int pano_cg_control_reset(pano_ctx *ctx) {
    static_assert(offsetof(PanoCgControl, error) == 0, "the sticky error word leads the control block");
    const size_t skip = offsetof(PanoCgControl, barrier);
    PANO_CUDA(cudaMemsetAsync(reinterpret_cast<char *>(ctx->d_cg) + skip, 0, sizeof(PanoCgControl) - skip, ctx->stream));
    return PANO_OK;
}

// h_cg holds a fresh copy of (at least) the error word.  A raised flag is reported once and cleared, so the context stays usable.
int pano_check_device_error(pano_ctx *ctx, const char *where) {
    if (!ctx->h_cg->error) return PANO_OK;
    ctx->h_cg->error = 0;
    cudaMemsetAsync(&ctx->d_cg->error, 0, sizeof(unsigned int), ctx->stream);
    PANO_FAIL(PANO_ERR_TIMEOUT, "%s: a bounded wait expired inside a persistent kernel (grid barrier / pipeline timeout)", where);
}

extern "C" {

const char *pano_version(void) { return "panopaea_b200 0.1 (sm_100a)"; }
const char *pano_last_error(void) { return g_err; }

int pano_device_count(int *count) {
    if (!count) PANO_FAIL(PANO_ERR_INVALID, "pano_device_count: null out pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return PANO_OK;
}

static int ctx_init(pano_ctx *c, int device, const cudaDeviceProp &prop, void *stream);

int pano_ctx_create(int device, void *stream, pano_ctx **out) {
    if (!out) PANO_FAIL(PANO_ERR_INVALID, "pano_ctx_create: null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        PANO_FAIL(PANO_ERR_CUDA, "pano_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n) PANO_FAIL(PANO_ERR_INVALID, "pano_ctx_create: device %d out of range [0,%d)", device, n);
    PANO_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PANO_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        PANO_FAIL(PANO_ERR_CUDA, "pano_ctx_create: device %d is sm_%d%d; this build targets sm_100a only", device,
                  prop.major, prop.minor);
    pano_ctx *c = new pano_ctx();
    const int rc = ctx_init(c, device, prop, stream);
    if (rc != PANO_OK) {   // nothing half-built survives a failed create
        pano_ctx_destroy(c);
        return rc;
    }
    *out = c;
    return PANO_OK;
}

static int ctx_init(pano_ctx *c, int device, const cudaDeviceProp &prop, void *stream) {
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        PANO_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    PANO_CUDA(cudaEventCreate(&c->ev_start));
    PANO_CUDA(cudaEventCreate(&c->ev_stop));
    PANO_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    PANO_CUDA(cudaEventCreateWithFlags(&c->ev_advect, cudaEventDisableTiming));
    PANO_CUDA(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    PANO_CUDA(cudaMalloc(&c->d_scalars, 8 * sizeof(double)));
    PANO_CUDA(cudaMallocHost(&c->h_scalars, 8 * sizeof(double)));
    PANO_CUDA(cudaMalloc(&c->d_cg, sizeof(PanoCgControl)));
    PANO_CUDA(cudaMemset(c->d_cg, 0, sizeof(PanoCgControl)));
    PANO_CUDA(cudaMallocHost(&c->h_cg, sizeof(PanoCgControl)));
    PANO_TRY(pano_ensure_partials(c, 3 * 4096));
    // all-reduce units of the persistent kernels: allocated here so that no launch path ever calls cudaMalloc
    PANO_CUDA(cudaMalloc(&c->d_units, 4096 * 16));
    PANO_CUDA(cudaMemset(c->d_units, 0, 4096 * 16));
    PANO_CUDA(cudaMalloc(&c->d_inbox, kPanoInboxBytes));            // push exchange of the SM-resident CG kernel
    PANO_CUDA(cudaMemset(c->d_inbox, 0, kPanoInboxBytes));          // sequence 0 never matches
    return PANO_OK;
}

// releases everything the context owns; no field handle refers to it any more
static void ctx_finalize(pano_ctx *ctx) {
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_partials);
    cudaFree(ctx->d_scalars);
    cudaFreeHost(ctx->h_scalars);
    cudaFree(ctx->d_cg);
    cudaFree(ctx->d_units);
    cudaFree(ctx->d_inbox);
    cudaFree(ctx->d_mail);
    cudaFree(ctx->d_tparts);
    cudaFree(ctx->d_sr_scratch);
    cudaFree(ctx->d_sr_order);
    cudaFree(ctx->d_claim);
    cudaFree(ctx->d_adv_claim);
    cudaFreeHost(ctx->h_cg);
    for (cudaEvent_t e : ctx->phase_events) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->marks) cudaEventDestroy(e);
    for (cudaEvent_t e : {ctx->ev_start, ctx->ev_stop, ctx->ev_advect, ctx->ev_copy})
        if (e) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// Fields created on the context that are still alive keep its device state alive: the context goes away with the last of
// them (hosts with garbage collection release handles in no particular order; nothing may dangle across the boundary).
int pano_ctx_destroy(pano_ctx *ctx) {
    if (!ctx || ctx->destroy_pending) return PANO_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    pano_workspace_free_all(ctx);
    pano_mg_free_all(ctx);
    if (ctx->live_fields > 0) {
        ctx->destroy_pending = true;
        return PANO_OK;
    }
    ctx_finalize(ctx);
    return PANO_OK;
}

int pano_ctx_sync(pano_ctx *ctx) {
    if (!ctx) PANO_FAIL(PANO_ERR_INVALID, "pano_ctx_sync: null context");
    PANO_TRY(pano_activate(ctx));
    // the sticky device error word rides along: an asynchronous step whose solver timed out is reported here
    PANO_CUDA(cudaMemcpyAsync(&ctx->h_cg->error, &ctx->d_cg->error, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    PANO_CUDA(cudaStreamSynchronize(ctx->stream));
    return pano_check_device_error(ctx, "pano_ctx_sync");
}

int pano_ctx_stream(pano_ctx *ctx, void **stream) {
    if (!ctx || !stream) PANO_FAIL(PANO_ERR_INVALID, "pano_ctx_stream: null argument");
    *stream = (void *)ctx->stream;
    return PANO_OK;
}

int pano_ctx_num_sms(pano_ctx *ctx, int *n) {
    if (!ctx || !n) PANO_FAIL(PANO_ERR_INVALID, "pano_ctx_num_sms: null argument");
    *n = ctx->num_sms;
    return PANO_OK;
}

int pano_ctx_launch_count(pano_ctx *ctx, uint64_t *n) {
    if (!ctx || !n) PANO_FAIL(PANO_ERR_INVALID, "pano_ctx_launch_count: null argument");
    *n = ctx->launches;
    return PANO_OK;
}

int pano_timer_start(pano_ctx *ctx) {
    if (!ctx) PANO_FAIL(PANO_ERR_INVALID, "pano_timer_start: null context");
    PANO_TRY(pano_activate(ctx));
    PANO_CUDA(cudaEventRecord(ctx->ev_start, ctx->stream));
    return PANO_OK;
}

int pano_timer_stop_ms(pano_ctx *ctx, double *ms) {
    if (!ctx || !ms) PANO_FAIL(PANO_ERR_INVALID, "pano_timer_stop_ms: null argument");
    PANO_TRY(pano_activate(ctx));
    PANO_CUDA(cudaEventRecord(ctx->ev_stop, ctx->stream));
    PANO_CUDA(cudaEventSynchronize(ctx->ev_stop));
    float f = 0.f;
    PANO_CUDA(cudaEventElapsedTime(&f, ctx->ev_start, ctx->ev_stop));
    *ms = (double)f;
    return PANO_OK;
}

int pano_timer_mark(pano_ctx *ctx) {
    if (!ctx) PANO_FAIL(PANO_ERR_INVALID, "pano_timer_mark: null context");
    PANO_TRY(pano_activate(ctx));
    if (ctx->marks_used >= 4096) PANO_FAIL(PANO_ERR_INVALID, "pano_timer_mark: 4096 marks pending; read them with pano_timer_marks_ms");
    if (ctx->marks_used == (int)ctx->marks.size()) {
        cudaEvent_t e;
        PANO_CUDA(cudaEventCreate(&e));
        ctx->marks.push_back(e);
    }
    PANO_CUDA(cudaEventRecord(ctx->marks[ctx->marks_used++], ctx->stream));
    return PANO_OK;
}

int pano_timer_marks_ms(pano_ctx *ctx, double *ms_out, int cap, int *count) {
    if (!ctx || !ms_out || !count) PANO_FAIL(PANO_ERR_INVALID, "pano_timer_marks_ms: null argument");
    PANO_TRY(pano_activate(ctx));
    const int n = ctx->marks_used;
    *count = n;
    ctx->marks_used = 0;
    if (n == 0) return PANO_OK;
    PANO_CUDA(cudaEventSynchronize(ctx->marks[n - 1]));
    for (int i = 0; i + 1 < n && i < cap; ++i) {
        float f = 0.f;
        PANO_CUDA(cudaEventElapsedTime(&f, ctx->marks[i], ctx->marks[i + 1]));
        ms_out[i] = (double)f;
    }
    return PANO_OK;
}

int pano_ctx_step_times(pano_ctx *ctx, double ms_out[PANO_STEP_PHASES], int64_t *steps) {
    if (!ctx || !ms_out || !steps) PANO_FAIL(PANO_ERR_INVALID, "pano_ctx_step_times: null argument");
    PANO_TRY(pano_activate(ctx));
    PANO_TRY(pano_phase_drain(ctx));
    for (int i = 0; i < PANO_STEP_PHASES; ++i) {
        ms_out[i] = ctx->phase_ms[i];
        ctx->phase_ms[i] = 0.0;
    }
    *steps = ctx->phase_steps;
    ctx->phase_steps = 0;
    return PANO_OK;
}

int pano_ctx_cg_profile(pano_ctx *ctx, int64_t cycles_out[8]) {
    if (!ctx || !cycles_out) PANO_FAIL(PANO_ERR_INVALID, "pano_ctx_cg_profile: null argument");
    PANO_TRY(pano_activate(ctx));
    PANO_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(PanoCgControl), cudaMemcpyDeviceToHost, ctx->stream));
    PANO_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 8; ++i) cycles_out[i] = (int64_t)ctx->h_cg->prof[i];
    return PANO_OK;
}

int pano_ctx_cg_profile_ctas(pano_ctx *ctx, int64_t *cycles_out, int n) {
    if (!ctx || !cycles_out || n < 0 || (size_t)n > ctx->partials_cap) PANO_FAIL(PANO_ERR_INVALID, "pano_ctx_cg_profile_ctas: bad argument");
    PANO_TRY(pano_activate(ctx));
    PANO_CUDA(cudaMemcpyAsync(cycles_out, ctx->d_partials, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PANO_CUDA(cudaStreamSynchronize(ctx->stream));
    return PANO_OK;
}

int pano_ctx_set_option(pano_ctx *ctx, const char *key, int64_t value) {
    if (!ctx || !key) PANO_FAIL(PANO_ERR_INVALID, "pano_ctx_set_option: null argument");
    ctx->options[key] = value;
    return PANO_OK;
}

int pano_ctx_get_option(pano_ctx *ctx, const char *key, int64_t *value) {
    if (!ctx || !key || !value) PANO_FAIL(PANO_ERR_INVALID, "pano_ctx_get_option: null argument");
    auto it = ctx->options.find(key);
    if (it == ctx->options.end()) PANO_FAIL(PANO_ERR_INVALID, "pano_ctx_get_option: unknown option '%s'", key);
    *value = it->second;
    return PANO_OK;
}

int pano_host_alloc(size_t bytes, void **out) {
    if (!out) PANO_FAIL(PANO_ERR_INVALID, "pano_host_alloc: null out pointer");
    *out = nullptr;
    PANO_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return PANO_OK;
}

int pano_host_free(void *p) {
    if (!p) return PANO_OK;
    PANO_CUDA(cudaFreeHost(p));
    return PANO_OK;
}

// ------------------------------------------------------------------------------------ fields
int pano_field_num_elem(int kind, size_t h, size_t w, size_t *n) {
    if (!n) PANO_FAIL(PANO_ERR_INVALID, "pano_field_num_elem: null out pointer");
    if (kind < PANO_SIMPLEX0 || kind > PANO_SIMPLEX2) PANO_FAIL(PANO_ERR_INVALID, "pano_field_num_elem: bad kind %d", kind);
    *n = pano_num_elem(kind, h, w);
    return PANO_OK;
}

int pano_field_new(pano_ctx *ctx, int kind, int dtype, size_t h, size_t w, pano_field **out) {
    if (!ctx || !out) PANO_FAIL(PANO_ERR_INVALID, "pano_field_new: null argument");
    *out = nullptr;
    if (kind < PANO_SIMPLEX0 || kind > PANO_SIMPLEX2) PANO_FAIL(PANO_ERR_INVALID, "pano_field_new: bad kind %d", kind);
    if (dtype != PANO_F64 && dtype != PANO_F32) PANO_FAIL(PANO_ERR_INVALID, "pano_field_new: bad dtype %d", dtype);
    if (h > (size_t)1 << 20 || w > (size_t)1 << 20)
        PANO_FAIL(PANO_ERR_INVALID, "pano_field_new: grid %zux%zu exceeds the 2^20 per-axis limit", h, w);
    PANO_TRY(pano_activate(ctx));
    pano_field *f = new pano_field();
    f->ctx = ctx;
    f->kind = kind;
    f->dtype = dtype;
    f->h = h;
    f->w = w;
    f->n = pano_num_elem(kind, h, w);
    size_t bytes = f->n * pano_dtype_size(dtype);
    // +256 B slack: vector/TMA paths may touch (never use) a few elements past the end
    cudaError_t e = cudaMalloc(&f->d, bytes + 256);
    if (e != cudaSuccess) {
        delete f;
        PANO_FAIL(PANO_ERR_CUDA, "pano_field_new: cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
    }
    e = cudaMemsetAsync(f->d, 0, bytes + 256, ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(f->d);
        delete f;
        PANO_FAIL(PANO_ERR_CUDA, "pano_field_new: cudaMemsetAsync -> %s", cudaGetErrorString(e));
    }
    pano_ctx_field_born(ctx);
    *out = f;
    return PANO_OK;
}

int pano_field_free(pano_field *f) {
    if (!f) return PANO_OK;
    pano_ctx *ctx = f->ctx;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(f->d);
    delete f;
    if (ctx && --ctx->live_fields == 0 && ctx->destroy_pending) ctx_finalize(ctx);   // the context was waiting for this handle
    return PANO_OK;
}

int pano_field_info(const pano_field *f, int *kind, int *dtype, size_t *h, size_t *w, size_t *n) {
    PANO_TRY(pano_check_field(f, "pano_field_info"));
    if (kind) *kind = f->kind;
    if (dtype) *dtype = f->dtype;
    if (h) *h = f->h;
    if (w) *w = f->w;
    if (n) *n = f->n;
    return PANO_OK;
}

int pano_field_device_ptr(const pano_field *f, void **ptr) {
    PANO_TRY(pano_check_field(f, "pano_field_device_ptr"));
    if (!ptr) PANO_FAIL(PANO_ERR_INVALID, "pano_field_device_ptr: null out pointer");
    *ptr = f->d;
    return PANO_OK;
}

int pano_field_upload(pano_field *f, const void *host, size_t n_elems) {
    PANO_TRY(pano_check_field(f, "pano_field_upload"));
    if (!host && n_elems) PANO_FAIL(PANO_ERR_INVALID, "pano_field_upload: null host pointer");
    if (n_elems != f->n) PANO_FAIL(PANO_ERR_SHAPE, "pano_field_upload: %zu elements given, field has %zu", n_elems, f->n);
    PANO_TRY(pano_activate(f->ctx));
    PANO_CUDA(cudaMemcpyAsync(f->d, host, f->n * pano_dtype_size(f->dtype), cudaMemcpyHostToDevice, f->ctx->stream));
    PANO_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return PANO_OK;
}

int pano_field_download(const pano_field *f, void *host, size_t n_elems) {
    PANO_TRY(pano_check_field(f, "pano_field_download"));
    if (!host && n_elems) PANO_FAIL(PANO_ERR_INVALID, "pano_field_download: null host pointer");
    if (n_elems != f->n) PANO_FAIL(PANO_ERR_SHAPE, "pano_field_download: %zu elements asked, field has %zu", n_elems, f->n);
    PANO_TRY(pano_activate(f->ctx));
    PANO_CUDA(cudaMemcpyAsync(host, f->d, f->n * pano_dtype_size(f->dtype), cudaMemcpyDeviceToHost, f->ctx->stream));
    PANO_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return PANO_OK;
}

int pano_field_assign(pano_field *dst, const pano_field *src) {
    PANO_TRY(pano_check_field(dst, "pano_field_assign(dst)"));
    PANO_TRY(pano_check_field(src, "pano_field_assign(src)"));
    // the reference assigns flat views: only length and dtype must agree (dec_fluid.rs:63)
    if (dst->n != src->n || dst->dtype != src->dtype || dst->ctx != src->ctx)
        PANO_FAIL(PANO_ERR_SHAPE, "pano_field_assign: flat views differ (%zu vs %zu elements, dtype %d vs %d)", dst->n,
                  src->n, dst->dtype, src->dtype);
    if (dst->d == src->d) return PANO_OK;
    PANO_TRY(pano_activate(dst->ctx));
    PANO_CUDA(cudaMemcpyAsync(dst->d, src->d, dst->n * pano_dtype_size(dst->dtype), cudaMemcpyDeviceToDevice,
                              dst->ctx->stream));
    return PANO_OK;
}

int pano_field_swap(pano_field *a, pano_field *b) {
    PANO_TRY(pano_check_field(a, "pano_field_swap(a)"));
    PANO_TRY(pano_check_field(b, "pano_field_swap(b)"));
    PANO_TRY(pano_check_same(a, b, "pano_field_swap"));
    void *t = a->d;
    a->d = b->d;
    b->d = t;
    return PANO_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------ helpers
void pano_ctx_field_born(pano_ctx *ctx) { ++ctx->live_fields; }

int pano_check_field(const pano_field *f, const char *name) {
    if (!f || !f->ctx || !f->d) PANO_FAIL(PANO_ERR_INVALID, "%s: null or freed field handle", name);
    return PANO_OK;
}

int pano_check_kind(const pano_field *f, int kind, const char *name) {
    PANO_TRY(pano_check_field(f, name));
    if (f->kind != kind) PANO_FAIL(PANO_ERR_SHAPE, "%s: expected a Simplex%d, got a Simplex%d", name, kind, f->kind);
    return PANO_OK;
}

int pano_check_same(const pano_field *a, const pano_field *b, const char *what) {
    if (a->ctx != b->ctx) PANO_FAIL(PANO_ERR_INVALID, "%s: fields belong to different contexts", what);
    if (a->kind != b->kind || a->dtype != b->dtype || a->h != b->h || a->w != b->w || a->dep != b->dep)
        PANO_FAIL(PANO_ERR_SHAPE, "%s: shape mismatch: Simplex%d %zux%zu dtype %d vs Simplex%d %zux%zu dtype %d", what,
                  a->kind, a->h, a->w, a->dtype, b->kind, b->h, b->w, b->dtype);
    return PANO_OK;
}

int pano_check_grid(const pano_field *a, const pano_field *b, const char *what) {
    if (a->ctx != b->ctx) PANO_FAIL(PANO_ERR_INVALID, "%s: fields belong to different contexts", what);
    if (a->dtype != b->dtype || a->h != b->h || a->w != b->w || a->dep != b->dep)
        PANO_FAIL(PANO_ERR_SHAPE, "%s: grid mismatch: %zux%zu dtype %d vs %zux%zu dtype %d", what, a->h, a->w, a->dtype,
                  b->h, b->w, b->dtype);
    return PANO_OK;
}

int pano_check_rect_within(const pano_rect &r, size_t rows, size_t cols, const char *what) {
    if (r.y0 < 0 || r.x0 < 0 || r.y1 < r.y0 || r.x1 < r.x0) PANO_FAIL(PANO_ERR_INVALID, "%s: malformed rectangle", what);
    if (r.y1 > r.y0 && r.x1 > r.x0 && ((size_t)r.y1 > rows || (size_t)r.x1 > cols))
        PANO_FAIL(PANO_ERR_SHAPE, "%s: rectangle [%lld,%lld)x[%lld,%lld) exceeds the %zux%zu grid (the reference would panic on the index)",
                  what, (long long)r.y0, (long long)r.y1, (long long)r.x0, (long long)r.x1, rows, cols);
    return PANO_OK;
}

int pano_activate(pano_ctx *ctx) {
    PANO_CUDA(cudaSetDevice(ctx->device));
    return PANO_OK;
}

int pano_after_launch(pano_ctx *ctx, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) PANO_FAIL(PANO_ERR_CUDA, "%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    ctx->launches++;
    return PANO_OK;
}

int pano_ensure_partials(pano_ctx *ctx, size_t doubles) {
    if (doubles <= ctx->partials_cap) return PANO_OK;
    if (ctx->d_partials) {
        PANO_CUDA(cudaStreamSynchronize(ctx->stream));
        PANO_CUDA(cudaFree(ctx->d_partials));
        ctx->d_partials = nullptr;
        ctx->partials_cap = 0;
    }
    PANO_CUDA(cudaMalloc(&ctx->d_partials, doubles * sizeof(double)));
    ctx->partials_cap = doubles;
    return PANO_OK;
}

// ---- per-phase step timing: slot k holds PANO_STEP_PHASES + 1 events (phase boundaries)
int pano_phase_mark(pano_ctx *ctx, int phase) {
    if (pano_option(ctx, "step_timing", 0) == 0) return PANO_OK;
    constexpr int per = PANO_STEP_PHASES + 1, kSlots = 64;
    if (ctx->phase_events.empty()) {
        ctx->phase_events.resize((size_t)per * kSlots);
        for (auto &e : ctx->phase_events) PANO_CUDA(cudaEventCreate(&e));
        ctx->phase_slots = kSlots;
    }
    if (phase == 0 && ctx->phase_used == ctx->phase_slots) PANO_TRY(pano_phase_drain(ctx));
    PANO_CUDA(cudaEventRecord(ctx->phase_events[(size_t)ctx->phase_used * per + phase], ctx->stream));
    if (phase == PANO_STEP_PHASES) ctx->phase_used++;
    return PANO_OK;
}

int pano_phase_drain(pano_ctx *ctx) {
    constexpr int per = PANO_STEP_PHASES + 1;
    if (ctx->phase_used == 0) return PANO_OK;
    PANO_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int s = 0; s < ctx->phase_used; ++s)
        for (int p = 0; p < PANO_STEP_PHASES; ++p) {
            float ms = 0.f;
            PANO_CUDA(cudaEventElapsedTime(&ms, ctx->phase_events[(size_t)s * per + p], ctx->phase_events[(size_t)s * per + p + 1]));
            ctx->phase_ms[p] += (double)ms;
        }
    ctx->phase_steps += ctx->phase_used;
    ctx->phase_used = 0;
    return PANO_OK;
}

// Look-up order: pano_ctx_set_option, then the environment variable PANO_OPT_<key> (lets a whole test run exercise a
// non-default kernel variant without touching the callers), then the built-in default.
int64_t pano_option(pano_ctx *ctx, const char *key, int64_t dflt) {
    auto it = ctx->options.find(key);
    if (it != ctx->options.end()) return it->second;
    const std::string name = std::string("PANO_OPT_") + key;
    if (const char *env = getenv(name.c_str())) {
        char *end = nullptr;
        const long long v = strtoll(env, &end, 10);
        if (end != env) return (int64_t)v;
    }
    return dflt;
}
