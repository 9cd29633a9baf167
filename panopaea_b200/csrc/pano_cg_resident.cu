// pano_cg_resident.cu -- the pressure solve for grids whose CG state fits ON CHIP.
//
// 148 SMs x (256 KB registers + 227 KB shared memory) hold the three CG vectors of a grid of up
// to ~1.2 M cells (1024^2 = BASELINE configs[1]).  One persistent CTA per SM owns a fixed
// rectangular tile for the whole solve:
//   x, r, z   live in REGISTERS (each thread owns one column x KR consecutive rows)
//   s         lives in SHARED memory with a one-cell halo frame
// and HBM is touched only to read b once and to write x / r / s once at the end.  Per iteration
//   P1: s' = r + beta*s on the tile and on its halo frame (the frame needs the neighbours' boundary
//       r values: a 4-line "mailbox" per tile in global memory, served from L2), z = A s', z.s'
//   P2: x += alpha s', r -= alpha z, r.r, max|r|; boundary r values are posted to the mailbox
// with the two grid-wide reductions done by the publish+poll all-reduce of pano_sm100.cuh.
// Arithmetic (expression order, no FMA contraction) is identical to the other two CG kernels
// (pcg.rs:14-82 with the closure of dec_fluid.rs:100-119); only the reduction order differs.
#include "pano_cell_math.h"
#include "pano_sm100.cuh"

using namespace pano_sm100;

namespace {

constexpr int kThreads = 512, kWarps = 16;
constexpr int kMaxCtas = 192;

struct ResArgs {
    double *x;
    const double *b;
    double *r, *s0;
    int h, w;
    double dt, threshold;
    int max_iter;
    RectI m;
    int tw, tw_log2;        // tile width (power of two, 32..512); row groups RG = 512 / tw; tile height = RG * KR
    int tiles_x, tiles_y;
    double *mail;           // [tiles][2*tw + 2*th]: top row, bottom row, left column, right column of r
    ReduceUnit *units;      // [2 banks][3 values][kMaxCtas]
    unsigned long long seq_base;
    PanoCgControl *ctl;
};

struct ResShared {          // placed behind the s tile
    double vals[3][kMaxCtas];
    double out[4];
    double wsum[3][kWarps];
    int ok;
};

__device__ __forceinline__ double cta_sum(double v, double *wsum) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) t += wsum[i];
    return t;
}
__device__ __forceinline__ double cta_max(double v, double *wsum) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) t = wsum[i] > t ? wsum[i] : t;
    return t;
}

// grid-wide all-reduce of up to 3 CTA totals; bit max_mask<k> selects max for value k
__device__ __forceinline__ bool grid_allreduce(const ResArgs &a, ResShared *sh, unsigned long long n, int nvals, double v0,
                                               double v1, double v2, unsigned max_mask, double *out) {
    const int G = gridDim.x, tid = threadIdx.x;
    const unsigned long long seq = a.seq_base + n;
    ReduceUnit *bank = a.units + (n & 1) * 3 * kMaxCtas;
    if (tid == 0) {
        __threadfence();
        unit_store(bank + 0 * kMaxCtas + blockIdx.x, v0, seq);
        if (nvals > 1) unit_store(bank + 1 * kMaxCtas + blockIdx.x, v1, seq);
        if (nvals > 2) unit_store(bank + 2 * kMaxCtas + blockIdx.x, v2, seq);
    }
    if (tid < G) {
        volatile unsigned int *err = &a.ctl->error;
        bool ok = true;
        for (int k = 0; k < nvals && ok; ++k) {
            double v;
            ok = unit_poll(bank + k * kMaxCtas + tid, seq, v, err);
            sh->vals[k][tid] = v;
        }
        __threadfence();
        if (!ok) sh->ok = 0;
    }
    __syncthreads();
    const int wid = tid >> 5, lane = tid & 31;
    if (wid < nvals) {
        double r = ((max_mask >> wid) & 1u) ? warp_fixed_max(sh->vals[wid], G, lane) : warp_fixed_sum(sh->vals[wid], G, lane);
        if (lane == 0) sh->out[wid] = r;
    }
    __syncthreads();
    out[0] = sh->out[0];
    if (nvals > 1) out[1] = sh->out[1];
    if (nvals > 2) out[2] = sh->out[2];
    return sh->ok != 0;
}

template <int KR>
__global__ void __launch_bounds__(kThreads, 1) k_cg_resident(const ResArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int tw = a.tw, RG = kThreads >> a.tw_log2, th = RG * KR;
    const int P = tw + 2;                                    // pitch of the s tile (with halo frame)
    double *S = smem;
    ResShared *sh = reinterpret_cast<ResShared *>(smem + (size_t)(th + 2) * P);
    const int tile = blockIdx.x, tcx = tile % a.tiles_x, tcy = tile / a.tiles_x;
    const int x0 = tcx * tw, y0 = tcy * th;
    const int c = tid & (tw - 1), rg = tid >> a.tw_log2;
    const int gx = x0 + c, gy0 = y0 + rg * KR;
    const int h = a.h, w = a.w;
    const int own0 = (rg * KR + 1) * P + c + 1;              // S index of this thread's first cell
    const int mail_stride = 2 * tw + 2 * th;
    double *my_mail = a.mail + (size_t)tile * mail_stride;
    if (tid == 0) sh->ok = 1;

    // per-thread cell masks: validity and the four "edge open" flags (walls and obstacle), KR bits each
    unsigned vmask = 0, mN = 0, mS = 0, mW = 0, mE = 0;
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        const int gy = gy0 + k;
        const bool valid = gy < h && gx < w;
        if (valid) {
            vmask |= 1u << k;
            if (gy > 0 && !in_rect(a.m, gy, gx)) mN |= 1u << k;
            if (gy < h - 1 && !in_rect(a.m, gy + 1, gx)) mS |= 1u << k;
            if (gx > 0 && !in_rect(a.m, gy, gx)) mW |= 1u << k;
            if (gx < w - 1 && !in_rect(a.m, gy, gx + 1)) mE |= 1u << k;
        }
    }
    const bool has_n = tcy > 0, has_s = tcy + 1 < a.tiles_y, has_w = tcx > 0, has_e = tcx + 1 < a.tiles_x;

    // ---- state: r (= b), x, z in registers; s = b in shared memory, halo frame from global b
    double r[KR], x[KR], z[KR];
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        r[k] = ((vmask >> k) & 1u) ? a.b[(size_t)(gy0 + k) * w + gx] : 0.0;
        x[k] = 0.0;
        z[k] = 0.0;
        S[own0 + k * P] = r[k];
    }
    for (int i = tid; i < 2 * P + 2 * th; i += kThreads) {   // frame: top row, bottom row, left col, right col
        int fy, fx;
        if (i < P) { fy = -1; fx = i - 1; }
        else if (i < 2 * P) { fy = th; fx = i - P - 1; }
        else if (i < 2 * P + th) { fy = i - 2 * P; fx = -1; }
        else { fy = i - 2 * P - th; fx = tw; }
        const int gy = y0 + fy, gxx = x0 + fx;
        const bool inside = gy >= 0 && gy < h && gxx >= 0 && gxx < w;
        S[(fy + 1) * P + fx + 1] = inside ? a.b[(size_t)gy * w + gxx] : 0.0;
    }
    __syncthreads();

    unsigned long long nred = 0;
    double sigma = 0, alpha = 0, beta = 0, rmax = 0, bmax = 0;
    int it = 0, applies = 0;
    bool converged = false, early = false, failed = false;
    double red[3];

    for (it = 0; it < a.max_iter; ++it) {
        const bool first = it == 0;
        // ------------------------------------------------------------------ P1
        if (!first) {
            // s' = r + beta*s: own cells from registers, halo frame from the neighbours' mailboxes
#pragma unroll
            for (int k = 0; k < KR; ++k) S[own0 + k * P] = r[k] + beta * S[own0 + k * P];
            for (int i = tid; i < 2 * tw + 2 * th; i += kThreads) {
                double rv;
                int si;
                bool have;
                if (i < tw) {                     // top frame row <- north tile's bottom row
                    have = has_n;
                    rv = have ? __ldcg(my_mail - (size_t)a.tiles_x * mail_stride + tw + i) : 0.0;
                    si = i + 1;
                } else if (i < 2 * tw) {          // bottom frame row <- south tile's top row
                    have = has_s;
                    rv = have ? __ldcg(my_mail + (size_t)a.tiles_x * mail_stride + (i - tw)) : 0.0;
                    si = (th + 1) * P + (i - tw) + 1;
                } else if (i < 2 * tw + th) {     // left frame column <- west tile's right column
                    have = has_w;
                    rv = have ? __ldcg(my_mail - mail_stride + 2 * tw + th + (i - 2 * tw)) : 0.0;
                    si = (i - 2 * tw + 1) * P;
                } else {                          // right frame column <- east tile's left column
                    have = has_e;
                    rv = have ? __ldcg(my_mail + mail_stride + 2 * tw + (i - 2 * tw - th)) : 0.0;
                    si = (i - 2 * tw - th + 1) * P + tw + 1;
                }
                if (have) S[si] = rv + beta * S[si];
            }
            __syncthreads();
        }
        double acc_zs = 0, acc_bb = 0, acc_bmax = 0;
        {
            double up = S[own0 - P], cur = S[own0];
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                const double dn = S[own0 + (k + 1) * P], wv = S[own0 + k * P - 1], ev = S[own0 + k * P + 1];
                const double zz = pano::laplacian_cell<double>(cur, up, dn, wv, ev, (mN >> k) & 1u, (mS >> k) & 1u,
                                                               (mW >> k) & 1u, (mE >> k) & 1u, a.dt);
                if ((vmask >> k) & 1u) {
                    z[k] = zz;
                    acc_zs = acc_zs + zz * cur;
                    if (first) {
                        const double ab = cur < 0 ? -cur : cur;
                        acc_bmax = ab > acc_bmax ? ab : acc_bmax;
                        acc_bb = acc_bb + cur * cur;
                    }
                }
                up = cur;
                cur = dn;
            }
        }
        {
            const double v0 = cta_sum(acc_zs, sh->wsum[0]);
            double v1 = 0, v2 = 0;
            if (first) {
                v1 = cta_sum(acc_bb, sh->wsum[1]);
                v2 = cta_max(acc_bmax, sh->wsum[2]);
            }
            if (!grid_allreduce(a, sh, nred++, first ? 3 : 1, v0, v1, v2, 0x4u, red)) { failed = true; break; }
        }
        const double zs = red[0];
        if (first) {
            sigma = red[1];                                    // pcg.rs:46
            bmax = red[2];                                     // pcg.rs:35
            rmax = bmax;
            if (bmax < a.threshold) { early = true; break; }   // pcg.rs:35-38
        }
        ++applies;
        alpha = sigma / zs;                                    // pcg.rs:53
        const double nalpha = -alpha;
        // ------------------------------------------------------------------ P2
        double acc_rr = 0, acc_rmax = 0;
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            if ((vmask >> k) & 1u) {
                const double sc = S[own0 + k * P];
                x[k] = x[k] + alpha * sc;                      // pcg.rs:55 (x = 0 before iteration 0)
                const double rn = r[k] + nalpha * z[k];        // pcg.rs:56
                r[k] = rn;
                const double ar = rn < 0 ? -rn : rn;
                acc_rmax = ar > acc_rmax ? ar : acc_rmax;
                acc_rr = acc_rr + rn * rn;
            }
        }
        // post the tile's boundary lines of r for the neighbours' next P1
        if (rg == 0) my_mail[c] = r[0];
        if (rg == RG - 1) my_mail[tw + c] = r[KR - 1];
        if (c == 0) {
#pragma unroll
            for (int k = 0; k < KR; ++k) my_mail[2 * tw + rg * KR + k] = r[k];
        }
        if (c == tw - 1) {
#pragma unroll
            for (int k = 0; k < KR; ++k) my_mail[2 * tw + th + rg * KR + k] = r[k];
        }
        {
            const double v0 = cta_sum(acc_rr, sh->wsum[0]);
            const double v1 = cta_max(acc_rmax, sh->wsum[1]);
            if (!grid_allreduce(a, sh, nred++, 2, v0, v1, 0.0, 0x2u, red)) { failed = true; break; }
        }
        const double rr = red[0];
        rmax = red[1];                                         // pcg.rs:58
        if (rmax < a.threshold) { converged = true; break; }   // pcg.rs:60-63
        beta = rr / sigma;                                     // pcg.rs:67-68
        sigma = rr;                                            // pcg.rs:79
    }
    if (failed) return;

    // ------------------------------------------------------------------ write the state back once
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        if ((vmask >> k) & 1u) {
            const size_t gi = (size_t)(gy0 + k) * w + gx;
            if (early) {
                a.x[gi] = 0.0;                                 // nothing else is touched (pcg.rs:35-38)
            } else {
                a.x[gi] = x[k];
                a.r[gi] = r[k];
                const double sv = S[own0 + k * P];
                // exhausted loop: the reference still performs the search update (pcg.rs:72-77)
                a.s0[gi] = converged ? sv : r[k] + beta * sv;
            }
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        a.ctl->iterations = early ? -1 : (converged ? it : a.max_iter);
        a.ctl->applies = applies;
        a.ctl->final_residual = rmax;
        a.ctl->rhs_max = bmax;
    }
}

struct ResPlan {
    int tw, tw_log2, kr, tiles_x, tiles_y, th;
};

constexpr int kKrChoices[] = {1, 2, 4, 8, 14, 16};

// smallest tile (in cells) over tw in {32..512}, KR in kKrChoices such that the tiles fit on the SMs
bool plan_resident(size_t h, size_t w, int num_sms, ResPlan *out) {
    bool found = false;
    long best_cells = 0, best_perim = 0;
    for (int lg = 5; lg <= 9; ++lg) {
        const int tw = 1 << lg, rg = kThreads / tw;
        const long tiles_x = ((long)w + tw - 1) / tw;
        if (tiles_x > num_sms) continue;
        const long max_ty = num_sms / tiles_x;
        const long th_need = ((long)h + max_ty - 1) / max_ty;
        for (int kr : kKrChoices) {
            const long th = (long)rg * kr;
            if (th < th_need) continue;
            const long tiles_y = ((long)h + th - 1) / th;
            if (tiles_x * tiles_y > num_sms || tiles_x * tiles_y > kMaxCtas) continue;
            const long cells = th * tw, perim = th + tw;
            if (!found || cells < best_cells || (cells == best_cells && perim < best_perim)) {
                found = true;
                best_cells = cells;
                best_perim = perim;
                *out = ResPlan{tw, lg, kr, (int)tiles_x, (int)tiles_y, (int)th};
            }
            break;   // larger KR only makes the tile bigger for this tw
        }
    }
    return found;
}

template <int KR>
int launch_kr(pano_ctx *ctx, ResArgs &a, int grid, size_t smem_bytes) {
    PANO_CUDA(cudaFuncSetAttribute(k_cg_resident<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    void *kargs[] = {(void *)&a};
    PANO_CUDA(cudaLaunchCooperativeKernel((const void *)k_cg_resident<KR>, dim3((unsigned)grid), dim3(kThreads), kargs, smem_bytes,
                                          ctx->stream));
    return pano_after_launch(ctx, "cg_resident");
}

}  // namespace

bool pano_cg_resident_supported(pano_ctx *ctx, size_t h, size_t w) {
    ResPlan p;
    return h >= 1 && w >= 1 && plan_resident(h, w, ctx->num_sms, &p);
}

int pano_cg_resident_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, size_t h, size_t w,
                            int max_iterations, double threshold, double timestep, RectI m) {
    ResPlan p;
    if (!plan_resident(h, w, ctx->num_sms, &p)) PANO_FAIL(PANO_ERR_INVALID, "cg_resident: a %zux%zu grid does not fit on chip", h, w);
    ResArgs a;
    a.x = x; a.b = b; a.r = r; a.s0 = s0;
    a.h = (int)h; a.w = (int)w;
    a.dt = timestep; a.threshold = threshold; a.max_iter = max_iterations;
    a.m = m;
    a.tw = p.tw; a.tw_log2 = p.tw_log2;
    a.tiles_x = p.tiles_x; a.tiles_y = p.tiles_y;
    const int grid = p.tiles_x * p.tiles_y;
    const size_t mail_doubles = (size_t)grid * (2 * p.tw + 2 * p.th);
    if (mail_doubles > ctx->mail_cap) {
        if (ctx->d_mail) {
            PANO_CUDA(cudaStreamSynchronize(ctx->stream));
            PANO_CUDA(cudaFree(ctx->d_mail));
            ctx->d_mail = nullptr;
            ctx->mail_cap = 0;
        }
        PANO_CUDA(cudaMalloc(&ctx->d_mail, mail_doubles * sizeof(double)));
        ctx->mail_cap = mail_doubles;
    }
    a.mail = ctx->d_mail;
    if (!ctx->d_units) PANO_CUDA(cudaMalloc(&ctx->d_units, 2 * 3 * kMaxCtas * sizeof(ReduceUnit)));
    if (ctx->launch_epoch == 0) PANO_CUDA(cudaMemsetAsync(ctx->d_units, 0, 2 * 3 * kMaxCtas * sizeof(ReduceUnit), ctx->stream));
    a.units = (ReduceUnit *)ctx->d_units;
    a.seq_base = (++ctx->launch_epoch) << 32;
    a.ctl = ctx->d_cg;
    PANO_CUDA(cudaMemsetAsync(ctx->d_cg, 0, sizeof(PanoCgControl), ctx->stream));
    const size_t smem_bytes = ((size_t)(p.th + 2) * (p.tw + 2)) * sizeof(double) + sizeof(ResShared) + 16;
    switch (p.kr) {
        case 1: return launch_kr<1>(ctx, a, grid, smem_bytes);
        case 2: return launch_kr<2>(ctx, a, grid, smem_bytes);
        case 4: return launch_kr<4>(ctx, a, grid, smem_bytes);
        case 8: return launch_kr<8>(ctx, a, grid, smem_bytes);
        case 14: return launch_kr<14>(ctx, a, grid, smem_bytes);
        case 16: return launch_kr<16>(ctx, a, grid, smem_bytes);
    }
    PANO_FAIL(PANO_ERR_INVALID, "cg_resident: no kernel for KR=%d", p.kr);
}
