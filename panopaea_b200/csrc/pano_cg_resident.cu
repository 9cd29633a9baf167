// pano_cg_resident.cu -- the pressure solve for grids whose CG state fits ON CHIP.
//
// 148 SMs x (256 KB registers + 227 KB shared memory) hold the three CG vectors of a grid of up
// to ~1.2 M cells (1024^2 = BASELINE configs[1]).  One persistent CTA per SM owns a fixed
// rectangular tile for the whole solve:
//   r, z      live in REGISTERS (each thread owns one column x KR consecutive rows)
//   s, x      live in SHARED memory (s with a one-cell halo frame)
// and HBM is touched only to read b once and to write x / r / s once at the end.  Per iteration
//   P1: s' = r + beta*s on the tile and on its halo frame (the frame needs the neighbours' boundary
//       r values: a 4-line "mailbox" per tile in global memory, served from L2), z = A s', z.s'
//   P2: x += alpha s', r -= alpha z, r.r, max|r|; boundary r values are posted to the mailbox
// with the two grid-wide reductions done by the publish+poll all-reduce of pano_sm100.cuh.
// Everything that crosses CTAs -- reduction partials AND mailbox values -- travels as 16-byte
// {value, sequence} units written with one store and polled by the reader, so the loop contains
// no memory fence at all (the NCCL "LL" idea applied to a stencil halo).
// Arithmetic (expression order, no FMA contraction) is identical to the other two CG kernels
// (pcg.rs:14-82 with the closure of dec_fluid.rs:100-119); only the reduction order differs.
#include "pano_cell_math.h"
#include "pano_sm100.cuh"

using namespace pano_sm100;

namespace {

constexpr int kThreads = 512, kWarps = 16;

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
    ReduceUnit *mail;       // [tiles][2*tw + 2*th] {value, seq}: top row, bottom row, left column, right column of r
    ReduceUnit *units;      // kUnitsTotal units (pano_sm100.cuh)
    unsigned long long seq_base;
    PanoCgControl *ctl;
};

struct ResShared {          // placed behind the s tile
    double vals[3][kMaxCtas];
    double out[4];
    double wsum[3][kWarps];
    int ok;
};

// deterministic CTA reduction of (sum, sum, max) in one pass: 2 barriers for all three values
__device__ __forceinline__ void cta_reduce3(double &v0, double &v1, double &v2, int nvals, unsigned max_mask, ResShared *sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v0 = (max_mask & 1u) ? warp_max(v0) : warp_sum(v0);
    if (nvals > 1) v1 = (max_mask & 2u) ? warp_max(v1) : warp_sum(v1);
    if (nvals > 2) v2 = (max_mask & 4u) ? warp_max(v2) : warp_sum(v2);
    __syncthreads();
    if (lane == 0) {
        sh->wsum[0][wid] = v0;
        if (nvals > 1) sh->wsum[1][wid] = v1;
        if (nvals > 2) sh->wsum[2][wid] = v2;
    }
    __syncthreads();
    double t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) {
        const double a0 = sh->wsum[0][i];
        t0 = (max_mask & 1u) ? (a0 > t0 ? a0 : t0) : t0 + a0;
        if (nvals > 1) {
            const double a1 = sh->wsum[1][i];
            t1 = (max_mask & 2u) ? (a1 > t1 ? a1 : t1) : t1 + a1;
        }
        if (nvals > 2) {
            const double a2 = sh->wsum[2][i];
            t2 = (max_mask & 4u) ? (a2 > t2 ? a2 : t2) : t2 + a2;
        }
    }
    v0 = t0; v1 = t1; v2 = t2;
}

// grid-wide all-reduce of up to 3 CTA totals; bit max_mask<k> selects max for value k
__device__ __forceinline__ bool grid_allreduce(const ResArgs &a, ResShared *sh, unsigned long long n, int nvals, double v0,
                                               double v1, double v2, unsigned max_mask, double *out) {
    return grid_allreduce_units(a.units, a.seq_base + n, n, nvals, v0, v1, v2, max_mask, sh->vals, sh->out,
                                &sh->ok, &a.ctl->error, /*fenced=*/false, [] { __syncthreads(); }, out);
}

// z = A s for the KR cells of one thread's column strip, read from the shared s tile.
// kMasked = false: every edge of every cell of the strip is open (no select instructions).
template <int KR, bool kMasked, bool kFirst>
__device__ __forceinline__ void strip_p1(const double *S, int own0, int P, double dt, unsigned vmask, unsigned mN, unsigned mS,
                                         unsigned mW, unsigned mE, double (&z)[KR], double &acc_zs, double &acc_bb,
                                         double &acc_bmax) {
    double up = S[own0 - P], cur = S[own0];
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        const double dn = S[own0 + (k + 1) * P], wv = S[own0 + k * P - 1], ev = S[own0 + k * P + 1];
        double zz;
        bool valid = true;
        if (kMasked) {
            valid = (vmask >> k) & 1u;
            zz = pano::laplacian_cell<double>(cur, up, dn, wv, ev, (mN >> k) & 1u, (mS >> k) & 1u, (mW >> k) & 1u,
                                              (mE >> k) & 1u, dt);
        } else {
            zz = pano::laplacian_cell<double>(cur, up, dn, wv, ev, true, true, true, true, dt);
        }
        if (valid) {
            z[k] = zz;
            acc_zs = acc_zs + zz * cur;
            if (kFirst) {
                const double ab = cur < 0 ? -cur : cur;
                acc_bmax = ab > acc_bmax ? ab : acc_bmax;
                acc_bb = acc_bb + cur * cur;
            }
        }
        up = cur;
        cur = dn;
    }
}

template <int KR>
__global__ void __launch_bounds__(kThreads, 1) k_cg_resident(const ResArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int tw = a.tw, RG = kThreads >> a.tw_log2, th = RG * KR;
    const int P = tw + 2;                                    // pitch of the s tile (with halo frame)
    double *S = smem;                                        // (th + 2) x P   search direction + halo frame
    double *X = smem + (size_t)(th + 2) * P;                 // th x tw        solution
    ResShared *sh = reinterpret_cast<ResShared *>(X + (size_t)th * tw);
    const int tile = blockIdx.x, tcx = tile % a.tiles_x, tcy = tile / a.tiles_x;
    const int x0 = tcx * tw, y0 = tcy * th;
    const int c = tid & (tw - 1), rg = tid >> a.tw_log2;
    const int gx = x0 + c, gy0 = y0 + rg * KR;
    const int h = a.h, w = a.w;
    const int own0 = (rg * KR + 1) * P + c + 1;              // S index of this thread's first cell
    const int xown0 = rg * KR * tw + c;                      // X index of this thread's first cell
    const int mail_stride = 2 * tw + 2 * th;
    ReduceUnit *my_mail = a.mail + (size_t)tile * mail_stride;
    volatile unsigned int *err = &a.ctl->error;
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
    constexpr unsigned kFull = (1u << KR) - 1u;
    const bool all_open = (vmask & mN & mS & mW & mE) == kFull;   // this thread's strip needs no masking
    const bool has_n = tcy > 0, has_s = tcy + 1 < a.tiles_y, has_w = tcx > 0, has_e = tcx + 1 < a.tiles_x;

    // Each thread serves at most two entries of the halo frame (2*tw + 2*th <= 2*512 for every plan):
    // where the boundary r value comes from (a neighbour's mailbox unit) and where s' goes in S.
    const ReduceUnit *h_src[2] = {nullptr, nullptr};
    int h_dst[2] = {0, 0};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int i = tid + q * kThreads;
        if (i < tw) {                                   // top frame row <- north tile's bottom row
            if (has_n) { h_src[q] = my_mail - (size_t)a.tiles_x * mail_stride + tw + i; h_dst[q] = i + 1; }
        } else if (i < 2 * tw) {                        // bottom frame row <- south tile's top row
            if (has_s) { h_src[q] = my_mail + (size_t)a.tiles_x * mail_stride + (i - tw); h_dst[q] = (th + 1) * P + (i - tw) + 1; }
        } else if (i < 2 * tw + th) {                   // left frame column <- west tile's right column
            if (has_w) { h_src[q] = my_mail - mail_stride + 2 * tw + th + (i - 2 * tw); h_dst[q] = (i - 2 * tw + 1) * P; }
        } else if (i < 2 * tw + 2 * th) {               // right frame column <- east tile's left column
            if (has_e) { h_src[q] = my_mail + mail_stride + 2 * tw + (i - 2 * tw - th); h_dst[q] = (i - 2 * tw - th + 1) * P + tw + 1; }
        }
    }

    // ---- state: r (= b) and z in registers; s = b and x = 0 in shared memory, halo frame from global b
    double r[KR], z[KR];
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        r[k] = ((vmask >> k) & 1u) ? a.b[(size_t)(gy0 + k) * w + gx] : 0.0;
        z[k] = 0.0;
        S[own0 + k * P] = r[k];
        X[xown0 + k * tw] = 0.0;
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
        double acc_zs = 0, acc_bb = 0, acc_bmax = 0;
        // ------------------------------------------------------------------ P1
        if (first) {
            if (all_open) strip_p1<KR, false, true>(S, own0, P, a.dt, vmask, mN, mS, mW, mE, z, acc_zs, acc_bb, acc_bmax);
            else strip_p1<KR, true, true>(S, own0, P, a.dt, vmask, mN, mS, mW, mE, z, acc_zs, acc_bb, acc_bmax);
        } else {
            // boundary r of the neighbours, posted in their previous P2 (already there: one L2 round trip)
            const unsigned long long want = a.seq_base + (unsigned long long)it;
            double hv[2] = {0.0, 0.0};
            bool ok = true;
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (h_src[q]) ok = unit_poll(h_src[q], want, hv[q], err) && ok;
            if (!ok) sh->ok = 0;
            // s' = r + beta*s: own cells from registers, halo frame from the mailbox values
#pragma unroll
            for (int k = 0; k < KR; ++k) S[own0 + k * P] = r[k] + beta * S[own0 + k * P];
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (h_src[q]) S[h_dst[q]] = hv[q] + beta * S[h_dst[q]];
            __syncthreads();
            if (all_open) strip_p1<KR, false, false>(S, own0, P, a.dt, vmask, mN, mS, mW, mE, z, acc_zs, acc_bb, acc_bmax);
            else strip_p1<KR, true, false>(S, own0, P, a.dt, vmask, mN, mS, mW, mE, z, acc_zs, acc_bb, acc_bmax);
        }
        cta_reduce3(acc_zs, acc_bb, acc_bmax, first ? 3 : 1, 0x4u, sh);
        if (!grid_allreduce(a, sh, nred++, first ? 3 : 1, acc_zs, acc_bb, acc_bmax, 0x4u, red)) { failed = true; break; }
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
        double acc_rr = 0, acc_rmax = 0, unused = 0;
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            if (all_open || ((vmask >> k) & 1u)) {
                const double sc = S[own0 + k * P];
                X[xown0 + k * tw] = X[xown0 + k * tw] + alpha * sc;   // pcg.rs:55 (x = 0 before iteration 0)
                const double rn = r[k] + nalpha * z[k];        // pcg.rs:56
                r[k] = rn;
                const double ar = rn < 0 ? -rn : rn;
                acc_rmax = ar > acc_rmax ? ar : acc_rmax;
                acc_rr = acc_rr + rn * rn;
            }
        }
        // post the tile's boundary lines of r for the neighbours' next P1 (tag: next iteration index)
        {
            const unsigned long long tag = a.seq_base + (unsigned long long)(it + 1);
            if (rg == 0 && has_n) unit_store(my_mail + c, r[0], tag);
            if (rg == RG - 1 && has_s) unit_store(my_mail + tw + c, r[KR - 1], tag);
            if (c == 0 && has_w) {
#pragma unroll
                for (int k = 0; k < KR; ++k) unit_store(my_mail + 2 * tw + rg * KR + k, r[k], tag);
            }
            if (c == tw - 1 && has_e) {
#pragma unroll
                for (int k = 0; k < KR; ++k) unit_store(my_mail + 2 * tw + th + rg * KR + k, r[k], tag);
            }
        }
        cta_reduce3(acc_rr, acc_rmax, unused, 2, 0x2u, sh);
        if (!grid_allreduce(a, sh, nred++, 2, acc_rr, acc_rmax, 0.0, 0x2u, red)) { failed = true; break; }
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
                a.x[gi] = X[xown0 + k * tw];
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
    const size_t mail_doubles = 2 * (size_t)grid * (2 * p.tw + 2 * p.th);   // 16-byte units
    if (mail_doubles > ctx->mail_cap) {
        if (ctx->d_mail) {
            PANO_CUDA(cudaStreamSynchronize(ctx->stream));
            PANO_CUDA(cudaFree(ctx->d_mail));
            ctx->d_mail = nullptr;
            ctx->mail_cap = 0;
        }
        PANO_CUDA(cudaMalloc(&ctx->d_mail, mail_doubles * sizeof(double)));
        PANO_CUDA(cudaMemsetAsync(ctx->d_mail, 0, mail_doubles * sizeof(double), ctx->stream));   // sequence 0 never matches
        ctx->mail_cap = mail_doubles;
    }
    a.mail = reinterpret_cast<ReduceUnit *>(ctx->d_mail);
    static_assert(kUnitsTotal * sizeof(ReduceUnit) <= 4096 * 16, "d_units (allocated in pano_ctx_create) is too small");
    a.units = (ReduceUnit *)ctx->d_units;
    a.seq_base = (++ctx->launch_epoch) << 32;
    a.ctl = ctx->d_cg;
    PANO_TRY(pano_cg_control_reset(ctx));
    const size_t smem_bytes = ((size_t)(p.th + 2) * (p.tw + 2) + (size_t)p.th * p.tw) * sizeof(double) + sizeof(ResShared) + 16;
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
