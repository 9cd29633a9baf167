// pano_cg_resident2.cu -- second-generation SM-resident CG kernel (see pano_cg_resident.cu for the idea).
//
// Same algorithm, arithmetic and data placement (r, z in registers; s with a halo frame and x in shared
// memory; boundary r values through {value, sequence} mailboxes; fence-free root all-reduce), but laid
// out for instruction count, which is what bounded the first version (~100 instructions per cell and
// iteration, profiles/r01_v3_resident_cg_1024_ncu.txt):
//   * tile geometry is a compile-time constant (template <KR, TW, T>): every shared-memory access is
//     base register + immediate
//   * each thread owns TWO adjacent columns x KR rows: own data moves with 128-bit LDS/STS, only the
//     west / east neighbours need a 64-bit load (2.5 shared-memory instructions per cell and phase
//     instead of 5-6); the halo frame is two columns wide so that pairs stay 16-byte aligned
//   * strips without a wall / obstacle / ragged edge take a select-free path
// Pressure solve of examples/dec_fluid.rs:91-119 = pcg.rs:14-82 with the closure of :100-119.
#include "pano_cell_math.h"
#include "pano_sm100.cuh"

using namespace pano_sm100;

namespace {

struct Res2Args {
    double *x;
    const double *b;
    double *r, *s0;
    int h, w;
    double dt, threshold;
    int max_iter;
    RectI m;
    int tiles_x, tiles_y;
    ReduceUnit *mail;       // [tiles][2*TW + 2*TH] {value, seq}: top row, bottom row, left column, right column of r
    ReduceUnit *units;
    ReduceUnit *inbox;      // push all-reduce (option "cg_push"; measured 7 % slower than the root protocol at 148 CTAs,
                            // 45 k scattered 16-byte stores per exchange: default off), or null for the root protocol
    unsigned long long seq_base;
    PanoCgControl *ctl;
    long long *dbg;         // optional: per-section clock64 totals of CTA 0 (option "cg_profile")
};

struct Res2Shared {
    double vals[3][kMaxCtas];
    double out[4];
    double wsum[3][32];
    int ok;
};

template <int T>
__device__ __forceinline__ void cta_reduce3(double &v0, double &v1, double &v2, int nvals, unsigned max_mask, Res2Shared *sh) {
    constexpr int kWarps = T / 32;
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
        t0 = (max_mask & 1u) ? fmax(a0, t0) : t0 + a0;
        if (nvals > 1) {
            const double a1 = sh->wsum[1][i];
            t1 = (max_mask & 2u) ? fmax(a1, t1) : t1 + a1;
        }
        if (nvals > 2) {
            const double a2 = sh->wsum[2][i];
            t2 = (max_mask & 4u) ? fmax(a2, t2) : t2 + a2;
        }
    }
    v0 = t0; v1 = t1; v2 = t2;
}

// KR rows per thread, TW tile width (cells), T threads.  Column pairs per row: TW/2; row groups: RG = T/(TW/2);
// tile height TH = RG*KR.  Shared s tile: (TH+2) rows x P = TW+4 columns, cell (ty,tx) at (ty+1)*P + tx + 2.
template <int KR, int TW, int T>
__global__ void __launch_bounds__(T, 1) k_cg_resident2(const Res2Args a) {
    constexpr int CP = TW / 2, RG = T / CP, TH = RG * KR, P = TW + 4;
    static_assert(T % CP == 0 && RG >= 1, "thread layout");
    extern __shared__ __align__(16) double smem[];
    double *S = smem;                                  // (TH+2) x P
    double *X = smem + (TH + 2) * P;                   // TH x TW
    Res2Shared *sh = reinterpret_cast<Res2Shared *>(X + TH * TW);
    const int tid = threadIdx.x;
    const int cp = tid % CP, rg = tid / CP;            // column pair, row group
    const int tile = blockIdx.x, tcx = tile % a.tiles_x, tcy = tile / a.tiles_x;
    const int x0 = tcx * TW, y0 = tcy * TH;
    const int gx = x0 + 2 * cp, gy0 = y0 + rg * KR;
    const int h = a.h, w = a.w;
    double *Sown = S + (rg * KR + 1) * P + 2 * cp + 2; // this thread's first pair
    double *Xown = X + rg * KR * TW + 2 * cp;
    constexpr int kMailStride = 2 * TW + 2 * TH;
    ReduceUnit *my_mail = a.mail + (size_t)tile * kMailStride;
    volatile unsigned int *err = &a.ctl->error;
    if (tid == 0) sh->ok = 1;

    // masks, KR bits per column (index 0: column gx, 1: column gx+1): validity and the four edge-open flags
    unsigned vm[2] = {0, 0}, mN[2] = {0, 0}, mS[2] = {0, 0}, mW[2] = {0, 0}, mE[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            const int gy = gy0 + k, gxx = gx + j;
            if (gy < h && gxx < w) {
                vm[j] |= 1u << k;
                if (gy > 0 && !in_rect(a.m, gy, gxx)) mN[j] |= 1u << k;
                if (gy < h - 1 && !in_rect(a.m, gy + 1, gxx)) mS[j] |= 1u << k;
                if (gxx > 0 && !in_rect(a.m, gy, gxx)) mW[j] |= 1u << k;
                if (gxx < w - 1 && !in_rect(a.m, gy, gxx + 1)) mE[j] |= 1u << k;
            }
        }
    }
    constexpr unsigned kFull = (1u << KR) - 1u;
    const bool all_open = (vm[0] & mN[0] & mS[0] & mW[0] & mE[0] & vm[1] & mN[1] & mS[1] & mW[1] & mE[1]) == kFull;
    const bool has_n = tcy > 0, has_s = tcy + 1 < a.tiles_y, has_w = tcx > 0, has_e = tcx + 1 < a.tiles_x;

    // halo frame entries served by this thread: mailbox unit of the neighbour -> position in S
    constexpr int kFrame = 2 * TW + 2 * TH, kPerThread = (kFrame + T - 1) / T;
    const ReduceUnit *h_src[kPerThread];
    int h_dst[kPerThread];
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
        const int i = tid + q * T;
        h_src[q] = nullptr;
        h_dst[q] = 0;
        if (i < TW) {                                   // top frame row <- north tile's bottom row
            if (has_n) { h_src[q] = my_mail - (size_t)a.tiles_x * kMailStride + TW + i; h_dst[q] = i + 2; }
        } else if (i < 2 * TW) {                        // bottom frame row <- south tile's top row
            if (has_s) { h_src[q] = my_mail + (size_t)a.tiles_x * kMailStride + (i - TW); h_dst[q] = (TH + 1) * P + (i - TW) + 2; }
        } else if (i < 2 * TW + TH) {                   // left frame column <- west tile's right column
            if (has_w) { h_src[q] = my_mail - kMailStride + 2 * TW + TH + (i - 2 * TW); h_dst[q] = (i - 2 * TW + 1) * P + 1; }
        } else if (i < kFrame) {                        // right frame column <- east tile's left column
            if (has_e) { h_src[q] = my_mail + kMailStride + 2 * TW + (i - 2 * TW - TH); h_dst[q] = (i - 2 * TW - TH + 1) * P + TW + 2; }
        }
    }

    // ---- state: r (= b) and z in registers; s = b and x = 0 in shared memory; halo frame from global b
    double2 r[KR], z[KR];
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        const size_t gi = (size_t)(gy0 + k) * w + gx;
        r[k].x = ((vm[0] >> k) & 1u) ? a.b[gi] : 0.0;
        r[k].y = ((vm[1] >> k) & 1u) ? a.b[gi + 1] : 0.0;
        z[k] = make_double2(0.0, 0.0);
        *reinterpret_cast<double2 *>(Sown + k * P) = r[k];
        *reinterpret_cast<double2 *>(Xown + k * TW) = make_double2(0.0, 0.0);
    }
    for (int i = tid; i < 2 * P + 2 * TH; i += T) {     // frame cells (the outermost frame columns are never read)
        int fy, fx;
        if (i < P) { fy = -1; fx = i - 2; }
        else if (i < 2 * P) { fy = TH; fx = i - P - 2; }
        else if (i < 2 * P + TH) { fy = i - 2 * P; fx = -1; }
        else { fy = i - 2 * P - TH; fx = TW; }
        const int gy = y0 + fy, gxx = x0 + fx;
        const bool inside = gy >= 0 && gy < h && gxx >= 0 && gxx < w;
        S[(fy + 1) * P + fx + 2] = inside ? a.b[(size_t)gy * w + gxx] : 0.0;
    }
    __syncthreads();

    unsigned long long nred = 0;
    double sigma = 0, alpha = 0, beta = 0, rmax = 0, bmax = 0;
    int it = 0, applies = 0;
    bool converged = false, early = false, failed = false;
    double red[3];

    auto allreduce = [&](int nvals, double v0, double v1, double v2, unsigned max_mask) {
        return grid_allreduce_units(a.units, a.seq_base + nred, nred, nvals, v0, v1, v2, max_mask, sh->vals, sh->out, &sh->ok,
                                    &a.ctl->error, /*fenced=*/false, [] { __syncthreads(); }, red, nullptr, NoWork(), a.inbox);
    };

    const bool prof = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    long long tprev = prof ? clock64() : 0;
    auto stamp = [&](int slot) {
        if (prof) {
            const long long t = clock64();
            a.dbg[slot] += t - tprev;
            tprev = t;
        }
    };
    for (it = 0; it < a.max_iter; ++it) {
        const bool first = it == 0;
        double acc_zs = 0, acc_bb = 0, acc_bmax = 0;
        // ------------------------------------------------------------------ P1
        if (!first) {
            const unsigned long long want = a.seq_base + (unsigned long long)it;
            double hv[kPerThread];
            bool ok = true;
#pragma unroll
            for (int q = 0; q < kPerThread; ++q) {
                hv[q] = 0.0;
                if (h_src[q]) ok = unit_poll(h_src[q], want, hv[q], err) && ok;
            }
            if (!ok) sh->ok = 0;
            stamp(0);   // mailbox poll
#pragma unroll
            for (int k = 0; k < KR; ++k) {                         // s' = r + beta*s  (pcg.rs:72-77)
                double2 sv = *reinterpret_cast<const double2 *>(Sown + k * P);
                sv.x = r[k].x + beta * sv.x;
                sv.y = r[k].y + beta * sv.y;
                *reinterpret_cast<double2 *>(Sown + k * P) = sv;
            }
#pragma unroll
            for (int q = 0; q < kPerThread; ++q)
                if (h_src[q]) S[h_dst[q]] = hv[q] + beta * S[h_dst[q]];
            __syncthreads();
            stamp(1);   // s' update + barrier
        }
        {   // z = A s'  (dec_fluid.rs:100-119), z.s' (+ b.b and max|b| in iteration 0)
            double2 up = *reinterpret_cast<const double2 *>(Sown - P), cur = *reinterpret_cast<const double2 *>(Sown);
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                const double2 dn = *reinterpret_cast<const double2 *>(Sown + (k + 1) * P);
                const double wv = Sown[k * P - 1], ev = Sown[k * P + 2];
                double z0, z1;
                if (all_open) {
                    z0 = pano::laplacian_cell<double>(cur.x, up.x, dn.x, wv, cur.y, true, true, true, true, a.dt);
                    z1 = pano::laplacian_cell<double>(cur.y, up.y, dn.y, cur.x, ev, true, true, true, true, a.dt);
                } else {
                    z0 = pano::laplacian_cell<double>(cur.x, up.x, dn.x, wv, cur.y, (mN[0] >> k) & 1u, (mS[0] >> k) & 1u,
                                                      (mW[0] >> k) & 1u, (mE[0] >> k) & 1u, a.dt);
                    z1 = pano::laplacian_cell<double>(cur.y, up.y, dn.y, cur.x, ev, (mN[1] >> k) & 1u, (mS[1] >> k) & 1u,
                                                      (mW[1] >> k) & 1u, (mE[1] >> k) & 1u, a.dt);
                    if (!((vm[0] >> k) & 1u)) z0 = 0.0;            // cells outside the grid hold s = 0 and contribute nothing
                    if (!((vm[1] >> k) & 1u)) z1 = 0.0;
                }
                z[k].x = z0;
                z[k].y = z1;
                acc_zs = acc_zs + z0 * cur.x;
                acc_zs = acc_zs + z1 * cur.y;
                if (first) {
                    acc_bmax = fmax(acc_bmax, fmax(fabs(cur.x), fabs(cur.y)));
                    acc_bb = acc_bb + cur.x * cur.x;
                    acc_bb = acc_bb + cur.y * cur.y;
                }
                up = cur;
                cur = dn;
            }
        }
        stamp(2);       // stencil
        cta_reduce3<T>(acc_zs, acc_bb, acc_bmax, first ? 3 : 1, 0x4u, sh);
        stamp(3);       // CTA reduction
        if (!allreduce(first ? 3 : 1, acc_zs, acc_bb, acc_bmax, 0x4u)) { failed = true; break; }
        stamp(4);       // grid all-reduce #1
        ++nred;
        const double zs = red[0];
        if (first) {
            sigma = red[1];                                        // pcg.rs:46
            bmax = red[2];                                         // pcg.rs:35
            rmax = bmax;
            if (bmax < a.threshold) { early = true; break; }       // pcg.rs:35-38
        }
        ++applies;
        alpha = sigma / zs;                                        // pcg.rs:53
        const double nalpha = -alpha;
        // ------------------------------------------------------------------ P2
        double acc_rr = 0, acc_rmax = 0, unused = 0;
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            r[k].x = r[k].x + nalpha * z[k].x;                     // pcg.rs:56 (cells outside the grid: 0 + a*0)
            r[k].y = r[k].y + nalpha * z[k].y;
            acc_rmax = fmax(acc_rmax, fmax(fabs(r[k].x), fabs(r[k].y)));
            acc_rr = acc_rr + r[k].x * r[k].x;
            acc_rr = acc_rr + r[k].y * r[k].y;
        }
        {   // post the tile's boundary lines of r for the neighbours' next P1 (tag: next iteration index)
            const unsigned long long tag = a.seq_base + (unsigned long long)(it + 1);
            if (rg == 0 && has_n) {
                unit_store(my_mail + 2 * cp, r[0].x, tag);
                unit_store(my_mail + 2 * cp + 1, r[0].y, tag);
            }
            if (rg == RG - 1 && has_s) {
                unit_store(my_mail + TW + 2 * cp, r[KR - 1].x, tag);
                unit_store(my_mail + TW + 2 * cp + 1, r[KR - 1].y, tag);
            }
            if (cp == 0 && has_w) {
#pragma unroll
                for (int k = 0; k < KR; ++k) unit_store(my_mail + 2 * TW + rg * KR + k, r[k].x, tag);
            }
            if (cp == CP - 1 && has_e) {
#pragma unroll
                for (int k = 0; k < KR; ++k) unit_store(my_mail + 2 * TW + TH + rg * KR + k, r[k].y, tag);
            }
        }
        stamp(5);       // P2 update + mailbox post
        cta_reduce3<T>(acc_rr, acc_rmax, unused, 2, 0x2u, sh);
        stamp(6);       // CTA reduction
        // x += alpha s' (pcg.rs:55; x = 0 before iteration 0) is off the critical path: it runs inside the reduction's wait
        auto update_x = [&] {
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                const double2 sv = *reinterpret_cast<const double2 *>(Sown + k * P);
                double2 xv = *reinterpret_cast<const double2 *>(Xown + k * TW);
                xv.x = xv.x + alpha * sv.x;
                xv.y = xv.y + alpha * sv.y;
                *reinterpret_cast<double2 *>(Xown + k * TW) = xv;
            }
        };
        if (!grid_allreduce_units(a.units, a.seq_base + nred, nred, 2, acc_rr, acc_rmax, 0.0, 0x2u, sh->vals, sh->out, &sh->ok,
                                  &a.ctl->error, /*fenced=*/false, [] { __syncthreads(); }, red, nullptr, update_x, a.inbox)) {
            failed = true;
            break;
        }
        stamp(7);       // grid all-reduce #2 (+ x update)
        ++nred;
        const double rr = red[0];
        rmax = red[1];                                             // pcg.rs:58
        if (rmax < a.threshold) { converged = true; break; }       // pcg.rs:60-63
        beta = rr / sigma;                                         // pcg.rs:67-68
        sigma = rr;                                                // pcg.rs:79
    }
    if (failed) return;

    // ------------------------------------------------------------------ write the state back once
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        const size_t gi = (size_t)(gy0 + k) * w + gx;
        const double2 sv = *reinterpret_cast<const double2 *>(Sown + k * P);
        const double2 xv = *reinterpret_cast<const double2 *>(Xown + k * TW);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if ((vm[j] >> k) & 1u) {
                const double rj = j ? r[k].y : r[k].x, sj = j ? sv.y : sv.x, xj = j ? xv.y : xv.x;
                if (early) {
                    a.x[gi + j] = 0.0;                             // nothing else is touched (pcg.rs:35-38)
                } else {
                    a.x[gi + j] = xj;
                    a.r[gi + j] = rj;
                    a.s0[gi + j] = converged ? sj : rj + beta * sj;   // exhausted: trailing search update (pcg.rs:72-77)
                }
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

struct Cfg {
    int kr, tw, t;
};
// instantiated geometries, from big tiles to small
constexpr Cfg kCfgs[] = {{8, 256, 512}, {7, 256, 512}, {4, 256, 512}, {4, 128, 512}, {2, 128, 512}, {2, 64, 512}, {1, 64, 512}, {1, 32, 128}};

template <int KR, int TW, int T>
int launch_cfg(pano_ctx *ctx, Res2Args &a, int grid) {
    constexpr int CP = TW / 2, RG = T / CP, TH = RG * KR, P = TW + 4;
    const size_t smem_bytes = ((size_t)(TH + 2) * P + (size_t)TH * TW) * sizeof(double) + sizeof(Res2Shared) + 16;
    PANO_CUDA(cudaFuncSetAttribute(k_cg_resident2<KR, TW, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    void *kargs[] = {(void *)&a};
    PANO_CUDA(cudaLaunchCooperativeKernel((const void *)k_cg_resident2<KR, TW, T>, dim3((unsigned)grid), dim3(T), kargs, smem_bytes,
                                          ctx->stream));
    return pano_after_launch(ctx, "cg_resident2");
}

// the instantiated geometry with the fewest cells per CTA whose tiles fit on the SMs
bool plan2(size_t h, size_t w, int num_sms, Cfg *out, int *tiles_x, int *tiles_y) {
    bool found = false;
    long best = 0;
    for (const Cfg &c : kCfgs) {
        const int th = (c.t / (c.tw / 2)) * c.kr;
        const long tx = ((long)w + c.tw - 1) / c.tw, ty = ((long)h + th - 1) / th;
        if (tx * ty > num_sms || tx * ty > kMaxCtas || tx * ty > c.t) continue;   // the all-reduce polls one unit per thread
        const long cells = (long)th * c.tw;
        if (!found || cells < best) {
            found = true;
            best = cells;
            *out = c;
            *tiles_x = (int)tx;
            *tiles_y = (int)ty;
        }
    }
    return found;
}

}  // namespace

bool pano_cg_resident2_supported(pano_ctx *ctx, size_t h, size_t w) {
    Cfg c;
    int tx, ty;
    return h >= 1 && w >= 1 && plan2(h, w, ctx->num_sms, &c, &tx, &ty);
}

int pano_cg_resident2_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, size_t h, size_t w,
                             int max_iterations, double threshold, double timestep, RectI m) {
    Cfg c{0, 0, 0};
    int tx = 0, ty = 0;
    if (!plan2(h, w, ctx->num_sms, &c, &tx, &ty)) PANO_FAIL(PANO_ERR_INVALID, "cg_resident2: a %zux%zu grid does not fit on chip", h, w);
    Res2Args a;
    a.x = x; a.b = b; a.r = r; a.s0 = s0;
    a.h = (int)h; a.w = (int)w;
    a.dt = timestep; a.threshold = threshold; a.max_iter = max_iterations;
    a.m = m;
    a.tiles_x = tx; a.tiles_y = ty;
    const int grid = tx * ty;
    const int th = (c.t / (c.tw / 2)) * c.kr;
    const size_t mail_doubles = 2 * (size_t)grid * (2 * c.tw + 2 * th);   // 16-byte units
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
    a.units = (ReduceUnit *)ctx->d_units;
    a.inbox = pano_option(ctx, "cg_push", 0) ? (ReduceUnit *)ctx->d_inbox : nullptr;
    a.seq_base = (++ctx->launch_epoch) << 32;
    a.ctl = ctx->d_cg;
    a.dbg = pano_option(ctx, "cg_profile", 0) ? ctx->d_cg->prof : nullptr;   // device address of the 8 slots
    PANO_TRY(pano_cg_control_reset(ctx));
#define PANO_CFG(KR, TW, T) \
    if (c.kr == KR && c.tw == TW && c.t == T) return launch_cfg<KR, TW, T>(ctx, a, grid)
    PANO_CFG(8, 256, 512);
    PANO_CFG(7, 256, 512);
    PANO_CFG(4, 256, 512);
    PANO_CFG(4, 128, 512);
    PANO_CFG(2, 128, 512);
    PANO_CFG(2, 64, 512);
    PANO_CFG(1, 64, 512);
    PANO_CFG(1, 32, 128);
#undef PANO_CFG
    PANO_FAIL(PANO_ERR_INVALID, "cg_resident2: no kernel for KR=%d TW=%d T=%d", c.kr, c.tw, c.t);
}
