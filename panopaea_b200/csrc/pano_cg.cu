// pano_cg.cu -- the pressure solve: pcg::precond_conjugate_gradient (panopaea/src/pcg.rs:14-82)
// with the identity preconditioner `()` (pcg.rs:8-12) and the matrix-free Laplacian closure of
// examples/dec_fluid.rs:100-119, as ONE persistent cooperative kernel.
//
// Reference iteration (pcg.rs:48-80)                 This kernel
//   z = A s                                            phase P1: s' = r + beta*s (previous iteration's
//   alpha = sigma / (z . s)                                      search update, folded in), z = A s' is
//   x += alpha s ; r -= alpha z                                   evaluated but NOT stored, z.s' reduced
//   if max|r| < threshold: break                       -- grid barrier, every CTA reduces the partials --
//   sigma' = r . r ; beta = sigma'/sigma               phase P2: z recomputed from s', x += alpha s',
//   s = r + beta s                                               r -= alpha z, r.r and max|r| reduced
//                                                      -- grid barrier, convergence test --
// HBM traffic per cell and iteration: P1 reads r, s and writes s' (24 B), P2 reads s', r, x and
// writes r, x (40 B) = 64 B, against 240 B for the reference's 12 passes and 88 B for the
// three-kernel split of SURVEY.md 8(d).  s is double-buffered (search <-> auxiliary) because P1
// recomputes s' on the one-cell halo of every tile from the neighbours' OLD values.
//
// All scalars (sigma, alpha, beta, the convergence decision) stay on the device; every CTA
// reduces the per-CTA partials in the same fixed order, so all CTAs hold bit-identical scalars
// and the result is deterministic run to run.
#include <cooperative_groups.h>

#include "pano_cell_math.h"
#include "pano_internal.cuh"
#include "pano_gridsync.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kTileX = 32, kTileY = 32;             // generic kernel: 32 x 32 cells per tile, 4 rows per thread

template <class T>
struct CgArgs {
    T *x;
    const T *b;
    T *r;
    T *s0;   // `search`
    T *s1;   // `auxiliary`, the second search buffer
    int h, w;
    T dt, threshold;
    int max_iter;
    RectI m;
    double *partials;   // 5 * G doubles
    PanoCgControl *ctl;
    int tiles_x, tiles_y;
};

template <bool kCg, class T>
__device__ __forceinline__ T ld(const T *p) {
    if (kCg) return __ldcg(p);
    return *p;
}

// kCg: route every field load through L2 only (ld.global.cg); otherwise rely on the barrier's
// gpu-scope fence to invalidate L1 (same contract as cooperative_groups::grid_group::sync()).
template <class T, bool kCg>
__global__ void __launch_bounds__(kThreads) k_cg_generic(CgArgs<T> a) {
    __shared__ T scratch[32];
    __shared__ int s_flag;
    const int G = gridDim.x;
    const int h = a.h, w = a.w;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int ntiles = a.tiles_x * a.tiles_y;
    double *pA = a.partials, *pB = a.partials + G, *pC = a.partials + 2 * G;       // P1: z.s, b.b, max|b|
    double *pD = a.partials + 3 * G, *pE = a.partials + 4 * G;                     // P2: r.r, max|r|
    unsigned long long nbar = 0;

    T sigma = 0, alpha = 0, beta = 0, rmax = 0, bmax = 0;
    int it = 0, applies = 0;
    bool converged = false;
    T *s_cur = a.s0;   // buffer that receives s' in the current P1
    T *s_old = a.s1;

    for (it = 0; it < a.max_iter; ++it) {
        const bool first = it == 0;
        const T *r_src = first ? a.b : a.r;
        // ------------------------------------------------------------------ P1
        T acc_zs = 0, acc_bb = 0, acc_bmax = 0;
        for (int t = blockIdx.x; t < ntiles; t += G) {
            const int x = (t % a.tiles_x) * kTileX + tx;
            const int ybase = (t / a.tiles_x) * kTileY + ty;
            if (x >= w) continue;
#pragma unroll
            for (int k = 0; k < kTileY / 8; ++k) {
                const int y = ybase + k * 8;
                if (y >= h) break;
                const size_t i = (size_t)y * w + x;
                const bool oN = y > 0 && !in_rect(a.m, y, x), oS = y < h - 1 && !in_rect(a.m, y + 1, x);
                const bool oW = x > 0 && !in_rect(a.m, y, x), oE = x < w - 1 && !in_rect(a.m, y, x + 1);
                T c, n = 0, s = 0, l = 0, e = 0;
                if (first) {
                    c = ld<kCg>(r_src + i);
                    if (oN) n = ld<kCg>(r_src + i - w);
                    if (oS) s = ld<kCg>(r_src + i + w);
                    if (oW) l = ld<kCg>(r_src + i - 1);
                    if (oE) e = ld<kCg>(r_src + i + 1);
                    T ab = c < 0 ? -c : c;
                    acc_bmax = ab > acc_bmax ? ab : acc_bmax;
                    acc_bb = acc_bb + c * c;
                } else {
                    c = ld<kCg>(r_src + i) + beta * ld<kCg>(s_old + i);
                    if (oN) n = ld<kCg>(r_src + i - w) + beta * ld<kCg>(s_old + i - w);
                    if (oS) s = ld<kCg>(r_src + i + w) + beta * ld<kCg>(s_old + i + w);
                    if (oW) l = ld<kCg>(r_src + i - 1) + beta * ld<kCg>(s_old + i - 1);
                    if (oE) e = ld<kCg>(r_src + i + 1) + beta * ld<kCg>(s_old + i + 1);
                }
                if (!first) s_cur[i] = c;   // iteration 0: s = b is materialised in P2, after the early-out test
                const T z = pano::laplacian_cell<T>(c, n, s, l, e, oN, oS, oW, oE, a.dt);
                acc_zs = acc_zs + z * c;
            }
        }
        {
            T v = block_sum(acc_zs, scratch);
            if (threadIdx.x == 0) pA[blockIdx.x] = (double)v;
            if (first) {
                T v2 = block_sum(acc_bb, scratch);
                T v3 = block_max(acc_bmax, scratch);
                if (threadIdx.x == 0) {
                    pB[blockIdx.x] = (double)v2;
                    pC[blockIdx.x] = (double)v3;
                }
            }
        }
        if (!grid_barrier(a.ctl, (++nbar) * (unsigned long long)G, &s_flag)) return;
        const T zs = sum_partials<T>(pA, G, scratch);
        if (first) {
            sigma = sum_partials<T>(pB, G, scratch);          // pcg.rs:46  (aux = r = b)
            bmax = max_partials<T>(pC, G, scratch);           // pcg.rs:35
            rmax = bmax;
            if (bmax < a.threshold) {                         // early out: x stays zero, nothing else is touched
                for (int t = blockIdx.x; t < ntiles; t += G) {
                    const int x = (t % a.tiles_x) * kTileX + tx;
                    const int ybase = (t / a.tiles_x) * kTileY + ty;
                    if (x >= w) continue;
                    for (int k = 0; k < kTileY / 8; ++k) {
                        const int y = ybase + k * 8;
                        if (y < h) a.x[(size_t)y * w + x] = (T)0;
                    }
                }
                if (blockIdx.x == 0 && threadIdx.x == 0) {
                    a.ctl->iterations = -1;
                    a.ctl->applies = 0;
                    a.ctl->final_residual = (double)bmax;
                    a.ctl->rhs_max = (double)bmax;
                }
                return;
            }
        }
        ++applies;
        alpha = sigma / zs;                                   // pcg.rs:53
        const T nalpha = -alpha;
        // ------------------------------------------------------------------ P2
        T acc_rr = 0, acc_rmax = 0;
        const T *s_rd = first ? a.b : s_cur;                  // pcg.rs:40-42: s = aux = r = b
        for (int t = blockIdx.x; t < ntiles; t += G) {
            const int x = (t % a.tiles_x) * kTileX + tx;
            const int ybase = (t / a.tiles_x) * kTileY + ty;
            if (x >= w) continue;
#pragma unroll
            for (int k = 0; k < kTileY / 8; ++k) {
                const int y = ybase + k * 8;
                if (y >= h) break;
                const size_t i = (size_t)y * w + x;
                const bool oN = y > 0 && !in_rect(a.m, y, x), oS = y < h - 1 && !in_rect(a.m, y + 1, x);
                const bool oW = x > 0 && !in_rect(a.m, y, x), oE = x < w - 1 && !in_rect(a.m, y, x + 1);
                const T c = ld<kCg>(s_rd + i);
                const T n = oN ? ld<kCg>(s_rd + i - w) : (T)0, s = oS ? ld<kCg>(s_rd + i + w) : (T)0;
                const T l = oW ? ld<kCg>(s_rd + i - 1) : (T)0, e = oE ? ld<kCg>(s_rd + i + 1) : (T)0;
                const T z = pano::laplacian_cell<T>(c, n, s, l, e, oN, oS, oW, oE, a.dt);
                if (first) a.s0[i] = c;
                const T xo = first ? (T)0 : ld<kCg>(a.x + i);
                a.x[i] = xo + alpha * c;                      // pcg.rs:55
                const T rn = ld<kCg>(r_src + i) + nalpha * z; // pcg.rs:56
                a.r[i] = rn;
                const T ar = rn < 0 ? -rn : rn;
                acc_rmax = ar > acc_rmax ? ar : acc_rmax;
                acc_rr = acc_rr + rn * rn;
            }
        }
        {
            T v = block_sum(acc_rr, scratch);
            T v2 = block_max(acc_rmax, scratch);
            if (threadIdx.x == 0) {
                pD[blockIdx.x] = (double)v;
                pE[blockIdx.x] = (double)v2;
            }
        }
        if (!grid_barrier(a.ctl, (++nbar) * (unsigned long long)G, &s_flag)) return;
        const T rr = sum_partials<T>(pD, G, scratch);
        rmax = max_partials<T>(pE, G, scratch);               // pcg.rs:58
        if (rmax < a.threshold) {                             // pcg.rs:60-63
            converged = true;
            break;
        }
        beta = rr / sigma;                                    // pcg.rs:67-68
        sigma = rr;                                           // pcg.rs:79
        T *tmp = s_cur;
        s_cur = s_old;
        s_old = tmp;
    }
    // After the loop `s_fin` is the buffer holding the last search direction that was applied.
    // converged: it is s_cur.  exhausted: the swap already happened, so it is s_old, and the
    // reference still performs the search update (pcg.rs:72-77) before leaving the loop.
    const T *s_fin = converged ? s_cur : s_old;
    if (a.max_iter > 0 && (!converged || s_fin != a.s0)) {
        for (int t = blockIdx.x; t < ntiles; t += G) {
            const int x = (t % a.tiles_x) * kTileX + tx;
            const int ybase = (t / a.tiles_x) * kTileY + ty;
            if (x >= w) continue;
            for (int k = 0; k < kTileY / 8; ++k) {
                const int y = ybase + k * 8;
                if (y >= h) break;
                const size_t i = (size_t)y * w + x;
                const T sv = ld<kCg>(s_fin + i);
                a.s0[i] = converged ? sv : ld<kCg>(a.r + i) + beta * sv;
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.ctl->iterations = converged ? it : a.max_iter;
        a.ctl->applies = applies;
        a.ctl->final_residual = (double)rmax;
        a.ctl->rhs_max = (double)bmax;
    }
}

template <class T>
int launch_generic(pano_ctx *ctx, CgArgs<T> &args, bool use_cg_loads) {
    const void *fn = use_cg_loads ? (const void *)k_cg_generic<T, true> : (const void *)k_cg_generic<T, false>;
    int per_sm = 0;
    PANO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0));
    if (per_sm < 1) PANO_FAIL(PANO_ERR_CUDA, "cg: kernel does not fit on an SM");
    int64_t cap = pano_option(ctx, "cg_blocks_per_sm", 0);
    if (cap > 0 && cap < per_sm) per_sm = (int)cap;
    const int ntiles = args.tiles_x * args.tiles_y;
    int G = ctx->num_sms * per_sm;
    if (G > ntiles) G = ntiles;
    if (G < 1) G = 1;
    PANO_TRY(pano_ensure_partials(ctx, 5 * (size_t)G));
    args.partials = ctx->d_partials;
    PANO_TRY(pano_cg_control_reset(ctx));
    void *kargs[] = {(void *)&args};
    PANO_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)G), dim3(kThreads), kargs, 0, ctx->stream));
    return pano_after_launch(ctx, "cg_generic");
}

}  // namespace

bool pano_cg_stream_supported(size_t h, size_t w, const void *x, const void *b, const void *r, const void *s0, const void *s1);
int pano_cg_stream_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, double *s1, size_t h, size_t w,
                          int max_iterations, double threshold, double timestep, RectI m, const PanoCgSlab *slab);

int pano_cg_sr_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *p, double *s0, double *r1, double *s1, size_t h,
                      size_t w, int max_iterations, double threshold, double timestep, RectI m, const PanoCgSrSlab *slab);

bool pano_cg_resident_supported(pano_ctx *ctx, size_t h, size_t w);
int pano_cg_resident_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, size_t h, size_t w,
                            int max_iterations, double threshold, double timestep, RectI m);

bool pano_cg_resident2_supported(pano_ctx *ctx, size_t h, size_t w);
int pano_cg_resident2_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, size_t h, size_t w,
                             int max_iterations, double threshold, double timestep, RectI m);

bool pano_cg_resident_sr_supported(pano_ctx *ctx, size_t h, size_t w);
int pano_cg_resident_sr_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, size_t h, size_t w,
                               int max_iterations, double threshold, double timestep, RectI m);

bool pano_cg_cluster_supported(pano_ctx *ctx, size_t h, size_t w);
int pano_cg_cluster_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, size_t h, size_t w, int max_iterations,
                           double threshold, double timestep, RectI m);

// Solve on raw device pointers.  s0 = search, s1 = auxiliary.  info nullable (non-null => sync).
int pano_cg_solve_raw(pano_ctx *ctx, int dtype, void *x, const void *b, void *r, void *s0, void *s1, size_t h, size_t w,
                      int max_iterations, double threshold, double timestep, pano_rect obstacle, pano_pcg_info *info) {
    if (h == 0 || w == 0) {
        if (info) *info = pano_pcg_info{-1, 0, 0.0, 0.0};
        return PANO_OK;
    }
    if (max_iterations <= 0) {
        // pcg.rs:32-46 with an empty loop: x = 0; then the early-out of :35-38 (scratch untouched) or r = s = b.
        // Rare enough to afford a host round trip for max|b|.
        const size_t bytes = h * w * pano_dtype_size(dtype);
        double bmax = 0.0;
        PANO_TRY(pano_norm_max_raw(ctx, dtype, b, h * w, &bmax));
        PANO_CUDA(cudaMemsetAsync(x, 0, bytes, ctx->stream));
        const bool early = bmax < threshold;
        if (!early) {
            PANO_CUDA(cudaMemcpyAsync(r, b, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            PANO_CUDA(cudaMemcpyAsync(s0, b, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        if (info) {
            PANO_CUDA(cudaStreamSynchronize(ctx->stream));
            *info = pano_pcg_info{early ? -1 : 0, 0, bmax, bmax};
        }
        return PANO_OK;
    }
    const bool cg_loads = pano_option(ctx, "cg_ldcg", 0) != 0;
    const RectI m = pano_clip_rect(obstacle, h + 1, w + 1);
    // kernel choice: "cg_kernel" 0 = auto, 1 = generic (any shape / dtype), 2 = TMA streaming (f64, even width),
    // 3 = SM-resident (f64, grids that fit on chip), 4 = first-generation SM-resident kernel (run-time tile geometry),
    // 5 = one thread-block cluster (f64, grids up to ~40 k cells), 6 = TMA streaming with ONE reduction per iteration
    // (pano_cg_sr.cu; f64, even width), 7 = SM-resident with ONE reduction per iteration (pano_cg_resident_sr.cu).
    // auto: cluster if it fits, else resident if it fits, else streaming (the single-reduction form when option
    // "cg_single_reduction" is 1, the default), else generic.
    const int64_t want = pano_option(ctx, "cg_kernel", 0);
    const bool stream_ok = dtype == PANO_F64 && pano_cg_stream_supported(h, w, x, b, r, s0, s1);
    const bool resident_ok = dtype == PANO_F64 && pano_cg_resident_supported(ctx, h, w);
    const bool resident2_ok = dtype == PANO_F64 && pano_cg_resident2_supported(ctx, h, w);
    const bool cluster_ok = dtype == PANO_F64 && pano_cg_cluster_supported(ctx, h, w);
    const bool resident_sr_ok = dtype == PANO_F64 && pano_cg_resident_sr_supported(ctx, h, w);
    // "cg_single_reduction": -1 auto (default), 0, 1.  Measured on B200 (profiles/r02_*): the single-reduction arrangement wins where
    // the reductions are a large share of an iteration -- on chip (1024^2: 1142 vs 980 Mcell-steps/s) and on the slabs of a
    // multi-GPU grid (pano_dist.cu) -- and is 1.5-4 % behind the two-reduction kernel on one GPU's large grids (it prefetches half of a
    // phase's boxes across its reductions, which a pass that rewrites all four vectors cannot).
    const int64_t sr_opt = pano_option(ctx, "cg_single_reduction", -1);
    const bool single = sr_opt > 0, single_auto = sr_opt < 0;
    if (want == 5 && !cluster_ok)
        PANO_FAIL(PANO_ERR_INVALID, "cg_kernel=5 (cluster) needs an f64 grid of at most 8 x 5120 cells and width <= 1024 (%zux%zu given)", h, w);
    if (want == 6 && !stream_ok)
        PANO_FAIL(PANO_ERR_INVALID, "cg_kernel=6 (single-reduction TMA streaming) needs f64 fields with an even width (%zux%zu given)", h, w);
    if (want == 2 && !stream_ok)
        PANO_FAIL(PANO_ERR_INVALID, "cg_kernel=2 (TMA streaming) needs f64 fields with an even width (%zux%zu given)", h, w);
    if ((want == 3 && !resident2_ok) || (want == 4 && !resident_ok) || (want == 7 && !resident_sr_ok))
        PANO_FAIL(PANO_ERR_INVALID, "cg_kernel=%d (SM-resident) needs an f64 grid that fits on chip (%zux%zu given)", (int)want, h, w);
    if (cluster_ok && (want == 0 || want == 5)) {
        PANO_TRY(pano_cg_cluster_launch(ctx, (double *)x, (const double *)b, (double *)r, (double *)s0, h, w, max_iterations, threshold,
                                        timestep, m));
    } else if (resident_sr_ok && (want == 7 || (want == 0 && (single || single_auto)))) {
        PANO_TRY(pano_cg_resident_sr_launch(ctx, (double *)x, (const double *)b, (double *)r, (double *)s0, h, w, max_iterations,
                                            threshold, timestep, m));
    } else if (resident2_ok && (want == 0 || want == 3)) {
        PANO_TRY(pano_cg_resident2_launch(ctx, (double *)x, (const double *)b, (double *)r, (double *)s0, h, w, max_iterations,
                                          threshold, timestep, m));
    } else if (resident_ok && (want == 0 || want == 4)) {
        PANO_TRY(pano_cg_resident_launch(ctx, (double *)x, (const double *)b, (double *)r, (double *)s0, h, w, max_iterations,
                                         threshold, timestep, m));
    } else if (stream_ok && (want == 6 || (want == 0 && single))) {
        // the second r and s buffers live in a scratch area the context keeps (two h x w arrays, 256-byte aligned)
        const size_t per = (h * w + 31) & ~(size_t)31;
        if (ctx->sr_scratch_cap < 2 * per) {
            if (ctx->d_sr_scratch) {
                PANO_CUDA(cudaStreamSynchronize(ctx->stream));
                PANO_CUDA(cudaFree(ctx->d_sr_scratch));
                ctx->d_sr_scratch = nullptr;
                ctx->sr_scratch_cap = 0;
            }
            PANO_CUDA(cudaMalloc((void **)&ctx->d_sr_scratch, 2 * per * sizeof(double)));
            ctx->sr_scratch_cap = 2 * per;
        }
        PANO_TRY(pano_cg_sr_launch(ctx, (double *)x, (const double *)b, (double *)r, (double *)s0, (double *)s1, ctx->d_sr_scratch,
                                   ctx->d_sr_scratch + per, h, w, max_iterations, threshold, timestep, m, nullptr));
    } else if (stream_ok && want != 1) {
        PANO_TRY(pano_cg_stream_launch(ctx, (double *)x, (const double *)b, (double *)r, (double *)s0, (double *)s1, h, w,
                                       max_iterations, threshold, timestep, m, nullptr));
    } else if (dtype == PANO_F64) {
        CgArgs<double> a{(double *)x, (const double *)b, (double *)r, (double *)s0, (double *)s1, (int)h, (int)w,
                         timestep, threshold, max_iterations, m, nullptr, ctx->d_cg,
                         ((int)w + kTileX - 1) / kTileX, ((int)h + kTileY - 1) / kTileY};
        PANO_TRY(launch_generic<double>(ctx, a, cg_loads));
    } else {
        CgArgs<float> a{(float *)x, (const float *)b, (float *)r, (float *)s0, (float *)s1, (int)h, (int)w,
                        (float)timestep, (float)threshold, max_iterations, m, nullptr, ctx->d_cg,
                        ((int)w + kTileX - 1) / kTileX, ((int)h + kTileY - 1) / kTileY};
        PANO_TRY(launch_generic<float>(ctx, a, cg_loads));
    }
    if (info) {
        PANO_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(PanoCgControl), cudaMemcpyDeviceToHost, ctx->stream));
        PANO_CUDA(cudaStreamSynchronize(ctx->stream));
        PANO_TRY(pano_check_device_error(ctx, "pano_pcg_solve"));
        info->iterations = ctx->h_cg->iterations;
        info->applies = ctx->h_cg->applies;
        info->final_residual = ctx->h_cg->final_residual;
        info->rhs_max = ctx->h_cg->rhs_max;
    }
    return PANO_OK;
}

extern "C" int pano_pcg_solve(int precond, pano_field *x, const pano_field *b, int32_t max_iterations, double threshold,
                              pano_field *residual, pano_field *auxiliary, pano_field *search, double timestep,
                              pano_rect obstacle, pano_pcg_info *info) {
    if (precond != PANO_PRECOND_IDENTITY && precond != PANO_PRECOND_JACOBI && precond != PANO_PRECOND_MULTIGRID)
        PANO_FAIL(PANO_ERR_INVALID, "pano_pcg_solve: unknown preconditioner kind %d", precond);
    const pano_field *all[] = {x, b, residual, auxiliary, search};
    const char *names[] = {"x", "b", "residual", "auxiliary", "search"};
    for (int i = 0; i < 5; ++i) {
        char nm[64];
        snprintf(nm, sizeof(nm), "pano_pcg_solve(%s)", names[i]);
        PANO_TRY(pano_check_kind(all[i], PANO_SIMPLEX2, nm));
        PANO_TRY(pano_check_same(all[0], all[i], "pano_pcg_solve"));
        for (int j = 0; j < i; ++j)
            if (all[i]->d == all[j]->d) PANO_FAIL(PANO_ERR_INVALID, "pano_pcg_solve: %s aliases %s", names[i], names[j]);
    }
    PANO_TRY(pano_check_rect_within(obstacle, x->h, x->w, "pano_pcg_solve(obstacle)"));
    pano_ctx *ctx = x->ctx;
    PANO_TRY(pano_activate(ctx));
    if (precond != PANO_PRECOND_IDENTITY)
        return pano_pcg_precond_raw(ctx, precond, x, b, max_iterations, threshold, residual, auxiliary, search, timestep, obstacle, info);
    return pano_cg_solve_raw(ctx, x->dtype, x->d, b->d, residual->d, search->d, auxiliary->d, x->h, x->w, max_iterations,
                             threshold, timestep, obstacle, info);
}
