// pano_prim.cu -- one kernel per reference operator: the Manifold2d methods on Grid2d
// (panopaea/src/dec/grid.rs:102-335) and the flat-view linear algebra the CG loop and the
// example use (panopaea/src/math/linear_view.rs:12-30, ndarray fill/assign/scaled_add).
// These keep the reference's composed call sequence working on device-resident fields and
// cross-check the fused kernels; the hot path itself runs through pano_fused.cu / pano_cg.cu.
#include "pano_internal.cuh"

namespace {

constexpr int kThreads = 256;

inline int flat_grid(pano_ctx *ctx, size_t n, int per_thread = 4) {
    size_t blocks = (n + (size_t)kThreads * per_thread - 1) / ((size_t)kThreads * per_thread);
    size_t cap = (size_t)ctx->num_sms * 8;
    if (blocks < 1) blocks = 1;
    return (int)(blocks < cap ? blocks : cap);
}

// ------------------------------------------------------------------ flat elementwise kernels
template <class T>
__global__ void k_fill(T *__restrict__ dst, size_t n, T v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

template <class T>
__global__ void k_fill_rect(T *__restrict__ dst, size_t pitch, int y0, int y1, int x0, int x1, T v) {
    const int rw = x1 - x0;
    const size_t n = (size_t)(y1 - y0) * rw;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int y = y0 + (int)(i / rw), x = x0 + (int)(i % rw);
        dst[(size_t)y * pitch + x] = v;
    }
}

template <class T>
__global__ void k_scaled_add(T *__restrict__ y, T alpha, const T *__restrict__ x, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = y[i] + alpha * x[i];
}

template <class T>
__global__ void k_scale(T *__restrict__ x, T alpha, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] = x[i] * alpha;
}

template <class T>
__global__ void k_xpby(T *__restrict__ dst, const T *__restrict__ a, T beta, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = a[i] + beta * dst[i];
}

// sign: +1 copy, -1 negate (the two halves of Hodge<Simplex1>)
template <class T>
__global__ void k_copy_signed(T *__restrict__ dst, const T *__restrict__ src, size_t n0, size_t n, bool neg_first,
                              bool neg_second) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T v = src[i];
        bool neg = i < n0 ? neg_first : neg_second;
        dst[i] = neg ? -v : v;
    }
}

// ------------------------------------------------------------------ reductions (deterministic)
template <class T>
__global__ void k_dot_partial(const T *__restrict__ a, const T *__restrict__ b, size_t n, double *__restrict__ partial) {
    __shared__ T scratch[32];
    T acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc = acc + a[i] * b[i];
    T t = block_sum(acc, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = (double)t;
}

template <class T>
__global__ void k_absmax_partial(const T *__restrict__ a, size_t n, double *__restrict__ partial) {
    __shared__ T scratch[32];
    T acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T v = a[i];
        v = v < 0 ? -v : v;
        acc = v > acc ? v : acc;
    }
    T t = block_max(acc, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = (double)t;
}

// T selects the accumulation type of the second stage (the reference accumulates in T)
template <class T, bool kMax>
__global__ void k_reduce_final(const double *__restrict__ partial, int n, double *__restrict__ out) {
    __shared__ T scratch[32];
    T acc = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        T v = (T)partial[i];
        if (kMax) acc = v > acc ? v : acc;
        else acc = acc + v;
    }
    T t = kMax ? block_max(acc, scratch) : block_sum(acc, scratch);
    if (threadIdx.x == 0) *out = (double)t;
}

// ------------------------------------------------------------------ stencil-shaped operators
// derivative_1_primal  (dec/grid.rs:295-305): face = -bottom + top - left + right
template <class T>
__global__ void k_derivative_1_primal(T *__restrict__ faces, const T *__restrict__ edges, int h, int w) {
    const T *e0 = edges, *e1 = edges + (size_t)w * (h + 1);
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    if (x >= w) return;
    for (int y = blockIdx.y * 8 + (threadIdx.x >> 5); y < h; y += gridDim.y * 8) {
        T top = e0[(size_t)y * w + x], bottom = e0[(size_t)(y + 1) * w + x];
        T left = e1[(size_t)y * (w + 1) + x], right = e1[(size_t)y * (w + 1) + x + 1];
        faces[(size_t)y * w + x] = -bottom + top - left + right;
    }
}

// derivative_0_dual  (dec/grid.rs:318-334): interior edges only, boundary edges untouched
template <class T>
__global__ void k_derivative_0_dual(T *__restrict__ edges, const T *__restrict__ faces, int h, int w) {
    T *e0 = edges, *e1 = edges + (size_t)w * (h + 1);
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    if (x >= w) return;
    for (int y = blockIdx.y * 8 + (threadIdx.x >> 5); y < h; y += gridDim.y * 8) {
        T f = faces[(size_t)y * w + x];
        if (y >= 1) e0[(size_t)y * w + x] = -(f - faces[(size_t)(y - 1) * w + x]);
        if (x >= 1) e1[(size_t)y * (w + 1) + x] = faces[(size_t)y * w + x - 1] - f;
    }
}

// derivative_0_primal  (dec/grid.rs:274-288)
template <class T>
__global__ void k_derivative_0_primal(T *__restrict__ edges, const T *__restrict__ v, int h, int w) {
    T *e0 = edges, *e1 = edges + (size_t)w * (h + 1);
    const int W = w + 1;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    if (x > w) return;
    for (int y = blockIdx.y * 8 + (threadIdx.x >> 5); y <= h; y += gridDim.y * 8) {
        T c = v[(size_t)y * W + x];
        if (x < w) e0[(size_t)y * w + x] = v[(size_t)y * W + x + 1] - c;
        if (y < h) e1[(size_t)y * W + x] = v[(size_t)(y + 1) * W + x] - c;
    }
}

// Hodge<Simplex0> (dec/grid.rs:106-191) including its addressing quirk: the four "corners"
// are taken at CELL dims (0|h-1, 0|w-1) of the (h+1, w+1) vertex array and are written first,
// so later side / inner slices override them; the true far corners are never written.
template <class T>
__global__ void k_hodge_0(T *__restrict__ dst, const T *__restrict__ src, int h, int w, bool inverse) {
    const int H = h + 1, W = w + 1;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    if (x >= W) return;
    const T two = (T)2, four = (T)4;
    for (int y = blockIdx.y * 8 + (threadIdx.x >> 5); y < H; y += gridDim.y * 8) {
        const bool ymid = y >= 1 && y <= H - 2, xmid = x >= 1 && x <= W - 2;
        const T v = src[(size_t)y * W + x];
        if (ymid && xmid) dst[(size_t)y * W + x] = v;
        else if (((y == 0 || y == H - 1) && xmid) || ((x == 0 || x == W - 1) && ymid))
            dst[(size_t)y * W + x] = inverse ? v * two : v / two;
        else if ((y == 0 || y == h - 1) && (x == 0 || x == w - 1))
            dst[(size_t)y * W + x] = inverse ? v * four : v / four;
    }
}

inline dim3 grid2d(int rows, int cols) {
    int gy = (rows + 7) / 8;
    if (gy > 4096) gy = 4096;
    if (gy < 1) gy = 1;
    int gx = (cols + 31) / 32;
    if (gx < 1) gx = 1;
    return dim3((unsigned)gx, (unsigned)gy);
}

template <class F64, class F32>
int dispatch(int dtype, F64 f64, F32 f32) {
    if (dtype == PANO_F64) f64();
    else f32();
    return PANO_OK;
}

int reduce_to_host(pano_ctx *ctx, int dtype, int nblocks, bool is_max, double *out) {
    if (dtype == PANO_F64) {
        if (is_max) k_reduce_final<double, true><<<1, kThreads, 0, ctx->stream>>>(ctx->d_partials, nblocks, ctx->d_scalars);
        else k_reduce_final<double, false><<<1, kThreads, 0, ctx->stream>>>(ctx->d_partials, nblocks, ctx->d_scalars);
    } else {
        if (is_max) k_reduce_final<float, true><<<1, kThreads, 0, ctx->stream>>>(ctx->d_partials, nblocks, ctx->d_scalars);
        else k_reduce_final<float, false><<<1, kThreads, 0, ctx->stream>>>(ctx->d_partials, nblocks, ctx->d_scalars);
    }
    PANO_TRY(pano_after_launch(ctx, "reduce_final"));
    PANO_CUDA(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PANO_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = ctx->h_scalars[0];
    return PANO_OK;
}

}  // namespace

// max|a[k]| of a flat device array (synchronises); shared with the solver's empty-loop path (pano_cg.cu)
int pano_norm_max_raw(pano_ctx *ctx, int dtype, const void *a, size_t n, double *out) {
    if (n == 0) {
        *out = 0.0;
        return PANO_OK;
    }
    const int g = flat_grid(ctx, n);
    PANO_TRY(pano_ensure_partials(ctx, (size_t)g));
    if (dtype == PANO_F64) k_absmax_partial<double><<<g, kThreads, 0, ctx->stream>>>((const double *)a, n, ctx->d_partials);
    else k_absmax_partial<float><<<g, kThreads, 0, ctx->stream>>>((const float *)a, n, ctx->d_partials);
    PANO_TRY(pano_after_launch(ctx, "pano_field_norm_max"));
    return reduce_to_host(ctx, dtype, g, true, out);
}

extern "C" {

int pano_field_fill(pano_field *f, double value) {
    PANO_TRY(pano_check_field(f, "pano_field_fill"));
    pano_ctx *ctx = f->ctx;
    PANO_TRY(pano_activate(ctx));
    if (f->n == 0) return PANO_OK;
    if (value == 0.0) {   // +0.0 is all-zero bits in both types
        PANO_CUDA(cudaMemsetAsync(f->d, 0, f->n * pano_dtype_size(f->dtype), ctx->stream));
        return PANO_OK;
    }
    const int g = flat_grid(ctx, f->n);
    if (f->dtype == PANO_F64) k_fill<double><<<g, kThreads, 0, ctx->stream>>>((double *)f->d, f->n, value);
    else k_fill<float><<<g, kThreads, 0, ctx->stream>>>((float *)f->d, f->n, (float)value);
    return pano_after_launch(ctx, "pano_field_fill");
}

int pano_field_fill_rect(pano_field *f, int comp, pano_rect rect, double value) {
    PANO_TRY(pano_check_field(f, "pano_field_fill_rect"));
    if (f->dep != 0) PANO_FAIL(PANO_ERR_SHAPE, "pano_field_fill_rect: a Grid3d field (use pano_field3_fill_box)");
    pano_ctx *ctx = f->ctx;
    PANO_TRY(pano_activate(ctx));
    if (rect.y0 < 0 || rect.x0 < 0 || rect.y1 < rect.y0 || rect.x1 < rect.x0)
        PANO_FAIL(PANO_ERR_INVALID, "pano_field_fill_rect: malformed rectangle");
    struct Part { size_t off, rows, cols; };
    Part parts[2];
    int np = 0;
    if (f->kind == PANO_SIMPLEX2) {
        if (comp != PANO_COMP_ALL) PANO_FAIL(PANO_ERR_INVALID, "pano_field_fill_rect: component %d on a Simplex2", comp);
        parts[np++] = Part{0, f->h, f->w};
    } else if (f->kind == PANO_SIMPLEX0) {
        if (comp != PANO_COMP_ALL) PANO_FAIL(PANO_ERR_INVALID, "pano_field_fill_rect: component %d on a Simplex0", comp);
        parts[np++] = Part{0, f->h + 1, f->w + 1};
    } else {
        if (comp == PANO_COMP_ALL || comp == PANO_COMP_VY) parts[np++] = Part{0, f->h + 1, f->w};
        if (comp == PANO_COMP_ALL || comp == PANO_COMP_VX) parts[np++] = Part{f->w * (f->h + 1), f->h, f->w + 1};
        if (np == 0) PANO_FAIL(PANO_ERR_INVALID, "pano_field_fill_rect: bad component %d", comp);
    }
    // the reference's indexed writes panic when out of bounds: report it instead of clipping
    for (int i = 0; i < np; ++i)
        if ((size_t)rect.y1 > parts[i].rows || (size_t)rect.x1 > parts[i].cols)
            PANO_FAIL(PANO_ERR_SHAPE, "pano_field_fill_rect: rectangle [%lld,%lld)x[%lld,%lld) exceeds %zux%zu",
                      (long long)rect.y0, (long long)rect.y1, (long long)rect.x0, (long long)rect.x1, parts[i].rows,
                      parts[i].cols);
    if (rect.y1 == rect.y0 || rect.x1 == rect.x0) return PANO_OK;
    const size_t cells = (size_t)(rect.y1 - rect.y0) * (size_t)(rect.x1 - rect.x0);
    const int g = flat_grid(ctx, cells, 1);
    for (int i = 0; i < np; ++i) {
        if (f->dtype == PANO_F64)
            k_fill_rect<double><<<g, kThreads, 0, ctx->stream>>>((double *)f->d + parts[i].off, parts[i].cols, (int)rect.y0,
                                                                 (int)rect.y1, (int)rect.x0, (int)rect.x1, value);
        else
            k_fill_rect<float><<<g, kThreads, 0, ctx->stream>>>((float *)f->d + parts[i].off, parts[i].cols, (int)rect.y0,
                                                                (int)rect.y1, (int)rect.x0, (int)rect.x1, (float)value);
        PANO_TRY(pano_after_launch(ctx, "pano_field_fill_rect"));
    }
    return PANO_OK;
}

int pano_field_scaled_add(pano_field *y, double alpha, const pano_field *x) {
    PANO_TRY(pano_check_field(y, "pano_field_scaled_add(y)"));
    PANO_TRY(pano_check_field(x, "pano_field_scaled_add(x)"));
    if (y->n != x->n || y->dtype != x->dtype || y->ctx != x->ctx)
        PANO_FAIL(PANO_ERR_SHAPE, "pano_field_scaled_add: flat views differ (%zu vs %zu elements)", y->n, x->n);
    pano_ctx *ctx = y->ctx;
    PANO_TRY(pano_activate(ctx));
    if (y->n == 0) return PANO_OK;
    const int g = flat_grid(ctx, y->n);
    if (y->dtype == PANO_F64)
        k_scaled_add<double><<<g, kThreads, 0, ctx->stream>>>((double *)y->d, alpha, (const double *)x->d, y->n);
    else k_scaled_add<float><<<g, kThreads, 0, ctx->stream>>>((float *)y->d, (float)alpha, (const float *)x->d, y->n);
    return pano_after_launch(ctx, "pano_field_scaled_add");
}

int pano_field_scale(pano_field *x, double alpha) {
    PANO_TRY(pano_check_field(x, "pano_field_scale"));
    pano_ctx *ctx = x->ctx;
    PANO_TRY(pano_activate(ctx));
    if (x->n == 0) return PANO_OK;
    const int g = flat_grid(ctx, x->n);
    if (x->dtype == PANO_F64) k_scale<double><<<g, kThreads, 0, ctx->stream>>>((double *)x->d, alpha, x->n);
    else k_scale<float><<<g, kThreads, 0, ctx->stream>>>((float *)x->d, (float)alpha, x->n);
    return pano_after_launch(ctx, "pano_field_scale");
}

int pano_field_xpby(pano_field *dst, const pano_field *a, double beta) {
    PANO_TRY(pano_check_field(dst, "pano_field_xpby(dst)"));
    PANO_TRY(pano_check_field(a, "pano_field_xpby(a)"));
    if (dst->n != a->n || dst->dtype != a->dtype || dst->ctx != a->ctx)
        PANO_FAIL(PANO_ERR_SHAPE, "pano_field_xpby: flat views differ (%zu vs %zu elements)", dst->n, a->n);
    pano_ctx *ctx = dst->ctx;
    PANO_TRY(pano_activate(ctx));
    if (dst->n == 0) return PANO_OK;
    const int g = flat_grid(ctx, dst->n);
    if (dst->dtype == PANO_F64)
        k_xpby<double><<<g, kThreads, 0, ctx->stream>>>((double *)dst->d, (const double *)a->d, beta, dst->n);
    else k_xpby<float><<<g, kThreads, 0, ctx->stream>>>((float *)dst->d, (const float *)a->d, (float)beta, dst->n);
    return pano_after_launch(ctx, "pano_field_xpby");
}

int pano_field_dot(const pano_field *a, const pano_field *b, double *out) {
    PANO_TRY(pano_check_field(a, "pano_field_dot(a)"));
    PANO_TRY(pano_check_field(b, "pano_field_dot(b)"));
    if (!out) PANO_FAIL(PANO_ERR_INVALID, "pano_field_dot: null out pointer");
    if (a->n != b->n || a->dtype != b->dtype || a->ctx != b->ctx)
        PANO_FAIL(PANO_ERR_SHAPE, "pano_field_dot: flat views differ (%zu vs %zu elements)", a->n, b->n);
    pano_ctx *ctx = a->ctx;
    PANO_TRY(pano_activate(ctx));
    if (a->n == 0) {
        *out = 0.0;
        return PANO_OK;
    }
    const int g = flat_grid(ctx, a->n);
    PANO_TRY(pano_ensure_partials(ctx, (size_t)g));
    if (a->dtype == PANO_F64)
        k_dot_partial<double><<<g, kThreads, 0, ctx->stream>>>((const double *)a->d, (const double *)b->d, a->n, ctx->d_partials);
    else k_dot_partial<float><<<g, kThreads, 0, ctx->stream>>>((const float *)a->d, (const float *)b->d, a->n, ctx->d_partials);
    PANO_TRY(pano_after_launch(ctx, "pano_field_dot"));
    return reduce_to_host(ctx, a->dtype, g, false, out);
}

int pano_field_norm_max(const pano_field *a, double *out) {
    PANO_TRY(pano_check_field(a, "pano_field_norm_max"));
    if (!out) PANO_FAIL(PANO_ERR_INVALID, "pano_field_norm_max: null out pointer");
    PANO_TRY(pano_activate(a->ctx));
    return pano_norm_max_raw(a->ctx, a->dtype, a->d, a->n, out);
}

// ------------------------------------------------------------------ Hodge stars
static int hodge_1(pano_field *dst, const pano_field *src, bool neg_vy, bool neg_vx, const char *name) {
    PANO_TRY(pano_check_kind(dst, PANO_SIMPLEX1, name));
    PANO_TRY(pano_check_kind(src, PANO_SIMPLEX1, name));
    PANO_TRY(pano_check_same(dst, src, name));
    pano_ctx *ctx = dst->ctx;
    PANO_TRY(pano_activate(ctx));
    if (dst->n == 0) return PANO_OK;
    const size_t n0 = dst->w * (dst->h + 1);
    const int g = flat_grid(ctx, dst->n);
    if (dst->dtype == PANO_F64)
        k_copy_signed<double><<<g, kThreads, 0, ctx->stream>>>((double *)dst->d, (const double *)src->d, n0, dst->n, neg_vy, neg_vx);
    else k_copy_signed<float><<<g, kThreads, 0, ctx->stream>>>((float *)dst->d, (const float *)src->d, n0, dst->n, neg_vy, neg_vx);
    return pano_after_launch(ctx, name);
}

int pano_hodge_1_primal(pano_field *dual, const pano_field *primal) {
    return hodge_1(dual, primal, false, true, "pano_hodge_1_primal");   // vy copied, vx negated
}
int pano_hodge_1_dual(pano_field *primal, const pano_field *dual) {
    return hodge_1(primal, dual, true, false, "pano_hodge_1_dual");     // vy negated, vx copied
}

static int hodge_2(pano_field *dst, const pano_field *src, const char *name) {
    PANO_TRY(pano_check_kind(dst, PANO_SIMPLEX2, name));
    PANO_TRY(pano_check_kind(src, PANO_SIMPLEX2, name));
    PANO_TRY(pano_check_same(dst, src, name));
    return pano_field_assign(dst, src);
}
int pano_hodge_2_primal(pano_field *dual, const pano_field *primal) { return hodge_2(dual, primal, "pano_hodge_2_primal"); }
int pano_hodge_0_dual(pano_field *primal, const pano_field *dual) { return hodge_2(primal, dual, "pano_hodge_0_dual"); }

static int hodge_0(pano_field *dst, const pano_field *src, bool inverse, const char *name) {
    PANO_TRY(pano_check_kind(dst, PANO_SIMPLEX0, name));
    PANO_TRY(pano_check_kind(src, PANO_SIMPLEX0, name));
    PANO_TRY(pano_check_same(dst, src, name));
    if (dst->h < 1 || dst->w < 1) PANO_FAIL(PANO_ERR_SHAPE, "%s: empty grid (the reference would panic indexing [h-1])", name);
    pano_ctx *ctx = dst->ctx;
    PANO_TRY(pano_activate(ctx));
    const int h = (int)dst->h, w = (int)dst->w;
    dim3 g = grid2d(h + 1, w + 1);
    if (dst->dtype == PANO_F64) k_hodge_0<double><<<g, kThreads, 0, ctx->stream>>>((double *)dst->d, (const double *)src->d, h, w, inverse);
    else k_hodge_0<float><<<g, kThreads, 0, ctx->stream>>>((float *)dst->d, (const float *)src->d, h, w, inverse);
    return pano_after_launch(ctx, name);
}
int pano_hodge_0_primal(pano_field *dual, const pano_field *primal) { return hodge_0(dual, primal, false, "pano_hodge_0_primal"); }
int pano_hodge_2_dual(pano_field *primal, const pano_field *dual) { return hodge_0(primal, dual, true, "pano_hodge_2_dual"); }

// ------------------------------------------------------------------ exterior derivatives
int pano_derivative_0_primal(pano_field *edges, const pano_field *vertices) {
    PANO_TRY(pano_check_kind(edges, PANO_SIMPLEX1, "pano_derivative_0_primal(edges)"));
    PANO_TRY(pano_check_kind(vertices, PANO_SIMPLEX0, "pano_derivative_0_primal(vertices)"));
    PANO_TRY(pano_check_grid(edges, vertices, "pano_derivative_0_primal"));
    pano_ctx *ctx = edges->ctx;
    PANO_TRY(pano_activate(ctx));
    const int h = (int)edges->h, w = (int)edges->w;
    dim3 g = grid2d(h + 1, w + 1);
    if (edges->dtype == PANO_F64)
        k_derivative_0_primal<double><<<g, kThreads, 0, ctx->stream>>>((double *)edges->d, (const double *)vertices->d, h, w);
    else k_derivative_0_primal<float><<<g, kThreads, 0, ctx->stream>>>((float *)edges->d, (const float *)vertices->d, h, w);
    return pano_after_launch(ctx, "pano_derivative_0_primal");
}

int pano_derivative_1_primal(pano_field *faces, const pano_field *edges) {
    PANO_TRY(pano_check_kind(faces, PANO_SIMPLEX2, "pano_derivative_1_primal(faces)"));
    PANO_TRY(pano_check_kind(edges, PANO_SIMPLEX1, "pano_derivative_1_primal(edges)"));
    PANO_TRY(pano_check_grid(faces, edges, "pano_derivative_1_primal"));
    pano_ctx *ctx = faces->ctx;
    PANO_TRY(pano_activate(ctx));
    if (faces->n == 0) return PANO_OK;
    const int h = (int)faces->h, w = (int)faces->w;
    dim3 g = grid2d(h, w);
    if (faces->dtype == PANO_F64)
        k_derivative_1_primal<double><<<g, kThreads, 0, ctx->stream>>>((double *)faces->d, (const double *)edges->d, h, w);
    else k_derivative_1_primal<float><<<g, kThreads, 0, ctx->stream>>>((float *)faces->d, (const float *)edges->d, h, w);
    return pano_after_launch(ctx, "pano_derivative_1_primal");
}

int pano_derivative_0_dual(pano_field *edges, const pano_field *faces) {
    PANO_TRY(pano_check_kind(edges, PANO_SIMPLEX1, "pano_derivative_0_dual(edges)"));
    PANO_TRY(pano_check_kind(faces, PANO_SIMPLEX2, "pano_derivative_0_dual(faces)"));
    PANO_TRY(pano_check_grid(edges, faces, "pano_derivative_0_dual"));
    pano_ctx *ctx = edges->ctx;
    PANO_TRY(pano_activate(ctx));
    if (faces->n == 0) return PANO_OK;
    const int h = (int)faces->h, w = (int)faces->w;
    dim3 g = grid2d(h, w);
    if (edges->dtype == PANO_F64)
        k_derivative_0_dual<double><<<g, kThreads, 0, ctx->stream>>>((double *)edges->d, (const double *)faces->d, h, w);
    else k_derivative_0_dual<float><<<g, kThreads, 0, ctx->stream>>>((float *)edges->d, (const float *)faces->d, h, w);
    return pano_after_launch(ctx, "pano_derivative_0_dual");
}

int pano_derivative_1_dual(pano_field *vertices, const pano_field *edges) {
    (void)vertices;
    (void)edges;
    PANO_FAIL(PANO_ERR_UNIMPLEMENTED,
              "pano_derivative_1_dual: not implemented, as in the reference (panopaea/src/dec/grid.rs:308-312 is unimplemented!())");
}

}  // extern "C"
