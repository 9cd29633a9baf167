// pano_cg_resident_sr.cu -- the SM-resident CG kernel with ONE grid-wide reduction per iteration.
//
// Same data placement idea as pano_cg_resident2.cu (one CTA per SM owns a fixed tile for the whole solve, HBM is read
// once and written once, tile boundary values travel through {value, sequence} mailboxes, fence-free root all-reduce), in
// the Chronopoulos-Gear arrangement of pano_cg_sr.cu: the two dependent reductions of pcg.rs:53 and :67 (z.s and r.r)
// become one reduction of {r.r, (A r).r, max|r|}.  At 1024^2 the two grid all-reduces and the two CTA reductions were
// 6.2 of the 9.7 us per iteration of k_cg_resident2; here an iteration is
//     s = w + beta s (registers), r -= alpha s (shared memory, boundary lines posted to the mailboxes first),
//     p = r_old + beta p, x += alpha p (shared memory; runs while the mailbox units are in flight),
//     poll the neighbours' boundary lines into the halo frame, w = A r (stencil), ONE reduction.
// r (with a two-column halo frame), p and x live in shared memory, w and s = A p in registers.
// Pressure solve of examples/dec_fluid.rs:91-119 = pcg.rs:14-82 with the closure of :100-119; the iterates agree with the
// reference's to ~1e-14 relative (scripts/cgcg_numerics.py).
#include "pano_cell_math.h"
#include "pano_sm100.cuh"

using namespace pano_sm100;

namespace {

struct ResSrArgs {
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
    unsigned long long seq_base;
    PanoCgControl *ctl;
    long long *dbg;         // optional: per-section clock64 totals of CTA 0 (option "cg_profile")
};

struct ResSrShared {
    double vals[3][kMaxCtas];
    double out[4];
    double wsum[3][32];
    int ok;
};

template <int T>
__device__ __forceinline__ void cta_reduce_ssm(double &v0, double &v1, double &v2, ResSrShared *sh) {   // sum, sum, max
    constexpr int kWarps = T / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v0 = warp_sum(v0);
    v1 = warp_sum(v1);
    v2 = warp_max(v2);
    __syncthreads();
    if (lane == 0) {
        sh->wsum[0][wid] = v0;
        sh->wsum[1][wid] = v1;
        sh->wsum[2][wid] = v2;
    }
    __syncthreads();
    double t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) {
        t0 = t0 + sh->wsum[0][i];
        t1 = t1 + sh->wsum[1][i];
        t2 = fmax(sh->wsum[2][i], t2);
    }
    v0 = t0; v1 = t1; v2 = t2;
}

// KR rows per thread, TW tile width (cells), T threads.  Column pairs per row: TW/2; row groups: RG = T/(TW/2);
// tile height TH = RG*KR.  Shared r tile: (TH+2) rows x P = TW+4 columns, cell (ty,tx) at (ty+1)*P + tx + 2.
template <int KR, int TW, int T>
__global__ void __launch_bounds__(T, 1) k_cg_resident_sr(const ResSrArgs a) {
    constexpr int CP = TW / 2, RG = T / CP, TH = RG * KR, P = TW + 4;
    static_assert(T % CP == 0 && RG >= 1, "thread layout");
    extern __shared__ __align__(16) double smem[];
    double *R = smem;                                  // (TH+2) x P
    double *X = smem + (TH + 2) * P;                   // TH x TW
    double *PM = X + TH * TW;                          // TH x TW: the search direction p
    ResSrShared *sh = reinterpret_cast<ResSrShared *>(PM + TH * TW);
    const int tid = threadIdx.x;
    const int cp = tid % CP, rg = tid / CP;            // column pair, row group
    const int tile = blockIdx.x, tcx = tile % a.tiles_x, tcy = tile / a.tiles_x;
    const int x0 = tcx * TW, y0 = tcy * TH;
    const int gx = x0 + 2 * cp, gy0 = y0 + rg * KR;
    const int h = a.h, w = a.w;
    double *Rown = R + (rg * KR + 1) * P + 2 * cp + 2; // this thread's first pair
    double *Xown = X + rg * KR * TW + 2 * cp;
    double *Pown = PM + rg * KR * TW + 2 * cp;
    constexpr int kMailStride = 2 * TW + 2 * TH;
    ReduceUnit *my_mail = a.mail + (size_t)tile * kMailStride;
    volatile unsigned int *err = &a.ctl->error;
    if (tid == 0) sh->ok = 1;

    // masks, KR <= 8 bits each, packed (registers are what this kernel is short of): per column j (0: column gx, 1: gx+1)
    // mk[j] = validity | north-open << 8 | south-open << 16 | west-open << 24, and me = east-open of column 0 | of column 1 << 8
    static_assert(KR <= 8, "mask packing");
    unsigned mk[2] = {0, 0}, me = 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            const int gy = gy0 + k, gxx = gx + j;
            if (gy < h && gxx < w) {
                mk[j] |= 1u << k;
                if (gy > 0 && !in_rect(a.m, gy, gxx)) mk[j] |= 1u << (8 + k);
                if (gy < h - 1 && !in_rect(a.m, gy + 1, gxx)) mk[j] |= 1u << (16 + k);
                if (gxx > 0 && !in_rect(a.m, gy, gxx)) mk[j] |= 1u << (24 + k);
                if (gxx < w - 1 && !in_rect(a.m, gy, gxx + 1)) me |= 1u << (8 * j + k);
            }
        }
    }
    constexpr unsigned kFull = (1u << KR) - 1u;
    constexpr unsigned kFull4 = kFull | (kFull << 8) | (kFull << 16) | (kFull << 24);
    const bool all_open = mk[0] == kFull4 && mk[1] == kFull4 && me == (kFull | (kFull << 8));
    const bool has_n = tcy > 0, has_s = tcy + 1 < a.tiles_y, has_w = tcx > 0, has_e = tcx + 1 < a.tiles_x;

    // halo frame entries served by this thread: mailbox unit of the neighbour -> position in R
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

    // ---- state: r = b (shared memory, halo frame from global b), p = x = 0 (shared memory), s = 0 and w (registers)
    double2 wv_[KR], s[KR];
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        const size_t gi = (size_t)(gy0 + k) * w + gx;
        double2 rv;
        rv.x = ((mk[0] >> k) & 1u) ? a.b[gi] : 0.0;
        rv.y = ((mk[1] >> k) & 1u) ? a.b[gi + 1] : 0.0;
        s[k] = make_double2(0.0, 0.0);
        *reinterpret_cast<double2 *>(Rown + k * P) = rv;
        *reinterpret_cast<double2 *>(Xown + k * TW) = make_double2(0.0, 0.0);
        *reinterpret_cast<double2 *>(Pown + k * TW) = make_double2(0.0, 0.0);
    }
    for (int i = tid; i < 2 * P + 2 * TH; i += T) {     // frame cells (the outermost frame columns are never read)
        int fy, fx;
        if (i < P) { fy = -1; fx = i - 2; }
        else if (i < 2 * P) { fy = TH; fx = i - P - 2; }
        else if (i < 2 * P + TH) { fy = i - 2 * P; fx = -1; }
        else { fy = i - 2 * P - TH; fx = TW; }
        const int gy = y0 + fy, gxx = x0 + fx;
        const bool inside = gy >= 0 && gy < h && gxx >= 0 && gxx < w;
        R[(fy + 1) * P + fx + 2] = inside ? a.b[(size_t)gy * w + gxx] : 0.0;
    }
    __syncthreads();

    unsigned long long nred = 0;
    double gamma = 0, alpha = 0, beta = 0, rmax = 0, bmax = 0;
    int it = -1, applies = 0;
    bool converged = false, early = false, failed = false;
    double red[3];

    const bool prof = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    long long tprev = prof ? clock64() : 0;
    auto stamp = [&](int slot) {
        if (prof) {
            const long long t = clock64();
            a.dbg[slot] += t - tprev;
            tprev = t;
        }
    };
    // w = A r from the shared r tile, and the three partial sums of the pass
    auto stencil = [&](double &acc_g, double &acc_d, double &acc_max) {
        double2 up = *reinterpret_cast<const double2 *>(Rown - P), cur = *reinterpret_cast<const double2 *>(Rown);
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            const double2 dn = *reinterpret_cast<const double2 *>(Rown + (k + 1) * P);
            const double wv = Rown[k * P - 1], ev = Rown[k * P + 2];
            double z0, z1;
            if (all_open) {
                z0 = pano::laplacian_cell<double>(cur.x, up.x, dn.x, wv, cur.y, true, true, true, true, a.dt);
                z1 = pano::laplacian_cell<double>(cur.y, up.y, dn.y, cur.x, ev, true, true, true, true, a.dt);
            } else {
                z0 = pano::laplacian_cell<double>(cur.x, up.x, dn.x, wv, cur.y, (mk[0] >> (8 + k)) & 1u, (mk[0] >> (16 + k)) & 1u,
                                                  (mk[0] >> (24 + k)) & 1u, (me >> k) & 1u, a.dt);
                z1 = pano::laplacian_cell<double>(cur.y, up.y, dn.y, cur.x, ev, (mk[1] >> (8 + k)) & 1u, (mk[1] >> (16 + k)) & 1u,
                                                  (mk[1] >> (24 + k)) & 1u, (me >> (8 + k)) & 1u, a.dt);
                if (!((mk[0] >> k) & 1u)) z0 = 0.0;            // cells outside the grid hold r = 0 and contribute nothing
                if (!((mk[1] >> k) & 1u)) z1 = 0.0;
            }
            wv_[k].x = z0;
            wv_[k].y = z1;
            acc_g = acc_g + cur.x * cur.x;
            acc_g = acc_g + cur.y * cur.y;
            acc_d = acc_d + z0 * cur.x;
            acc_d = acc_d + z1 * cur.y;
            acc_max = fmax(acc_max, fmax(fabs(cur.x), fabs(cur.y)));
            up = cur;
            cur = dn;
        }
    };
    auto reduce = [&](double g, double d, double mx) {
        cta_reduce_ssm<T>(g, d, mx, sh);
        const bool ok = grid_allreduce_units(a.units, a.seq_base + nred, nred, 3, g, d, mx, 0x4u, sh->vals, sh->out, &sh->ok,
                                             &a.ctl->error, /*fenced=*/false, [] { __syncthreads(); }, red, nullptr, NoWork(), nullptr);
        ++nred;
        return ok;
    };

    {   // opening pass: w_0 = A b, gamma_0 = b.b, delta_0 = (A b).b, max|b|
        double g = 0, d = 0, mx = 0;
        stencil(g, d, mx);
        if (!reduce(g, d, mx)) return;
        gamma = red[0];                                            // pcg.rs:46
        bmax = red[2];                                             // pcg.rs:35
        rmax = bmax;
        alpha = gamma / red[1];                                    // pcg.rs:53 with s = r
        beta = 0.0;
        if (bmax < a.threshold) early = true;                      // pcg.rs:35-38
    }
    stamp(0);
    if (!early) {
        for (it = 0; it < a.max_iter; ++it) {
            const bool first = it == 0;
            const double nalpha = -alpha;
            const unsigned long long tag = a.seq_base + (unsigned long long)(it + 1);
            // ---- own cells: s = w + beta s, r -= alpha s (boundary lines posted at once), p = r_old + beta p, x += alpha p
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                const double2 ro = *reinterpret_cast<const double2 *>(Rown + k * P);
                if (first) s[k] = wv_[k];
                else { s[k].x = wv_[k].x + beta * s[k].x; s[k].y = wv_[k].y + beta * s[k].y; }
                double2 rn;
                rn.x = ro.x + nalpha * s[k].x;                     // pcg.rs:56 (cells outside the grid: 0 + a*0)
                rn.y = ro.y + nalpha * s[k].y;
                *reinterpret_cast<double2 *>(Rown + k * P) = rn;
                if (k == 0 && rg == 0 && has_n) {
                    unit_store(my_mail + 2 * cp, rn.x, tag);
                    unit_store(my_mail + 2 * cp + 1, rn.y, tag);
                }
                if (k == KR - 1 && rg == RG - 1 && has_s) {
                    unit_store(my_mail + TW + 2 * cp, rn.x, tag);
                    unit_store(my_mail + TW + 2 * cp + 1, rn.y, tag);
                }
                if (cp == 0 && has_w) unit_store(my_mail + 2 * TW + rg * KR + k, rn.x, tag);
                if (cp == CP - 1 && has_e) unit_store(my_mail + 2 * TW + TH + rg * KR + k, rn.y, tag);
                double2 pv = *reinterpret_cast<const double2 *>(Pown + k * TW);
                pv.x = ro.x + beta * pv.x;                         // pcg.rs:72-77 (beta_0 = 0, p_(-1) = 0: p_0 = r_0)
                pv.y = ro.y + beta * pv.y;
                *reinterpret_cast<double2 *>(Pown + k * TW) = pv;
                double2 xv = *reinterpret_cast<const double2 *>(Xown + k * TW);
                xv.x = xv.x + alpha * pv.x;                        // pcg.rs:55
                xv.y = xv.y + alpha * pv.y;
                *reinterpret_cast<double2 *>(Xown + k * TW) = xv;
            }
            stamp(1);   // update + mailbox post
            // ---- the neighbours' boundary lines of the new r into the halo frame
            {
                double hv[kPerThread];
                bool ok = true;
#pragma unroll
                for (int q = 0; q < kPerThread; ++q) {
                    hv[q] = 0.0;
                    if (h_src[q]) ok = unit_poll(h_src[q], tag, hv[q], err) && ok;
                }
                if (!ok) sh->ok = 0;
#pragma unroll
                for (int q = 0; q < kPerThread; ++q)
                    if (h_src[q]) R[h_dst[q]] = hv[q];
            }
            __syncthreads();
            stamp(2);   // mailbox poll + barrier
            double g = 0, d = 0, mx = 0;
            stencil(g, d, mx);
            stamp(3);   // stencil
            if (!reduce(g, d, mx)) { failed = true; break; }
            stamp(4);   // CTA reduction + grid all-reduce
            ++applies;
            const double gamma_new = red[0], delta = red[1];
            rmax = red[2];                                         // pcg.rs:58
            if (rmax < a.threshold) { converged = true; break; }   // pcg.rs:60-63
            beta = gamma_new / gamma;                              // pcg.rs:67-68
            alpha = gamma_new / (delta - beta * gamma_new / alpha);   // = sigma' / (z'.s'): Chronopoulos-Gear
            gamma = gamma_new;                                     // pcg.rs:79
        }
    }
    if (failed) return;

    // ------------------------------------------------------------------ write the state back once
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        const size_t gi = (size_t)(gy0 + k) * w + gx;
        const double2 rv = *reinterpret_cast<const double2 *>(Rown + k * P);
        const double2 pv = *reinterpret_cast<const double2 *>(Pown + k * TW);
        const double2 xv = *reinterpret_cast<const double2 *>(Xown + k * TW);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if ((mk[j] >> k) & 1u) {
                const double rj = j ? rv.y : rv.x, pj = j ? pv.y : pv.x, xj = j ? xv.y : xv.x;
                if (early) {
                    a.x[gi + j] = 0.0;                             // nothing else is touched (pcg.rs:35-38)
                } else {
                    a.x[gi + j] = xj;
                    a.r[gi + j] = rj;
                    a.s0[gi + j] = converged ? pj : rj + beta * pj;   // exhausted: trailing search update (pcg.rs:72-77)
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
constexpr size_t smem_of() {
    constexpr int CP = TW / 2, RG = T / CP, TH = RG * KR, P = TW + 4;
    return ((size_t)(TH + 2) * P + 2 * (size_t)TH * TW) * sizeof(double) + sizeof(ResSrShared) + 16;
}

template <int KR, int TW, int T>
int launch_cfg(pano_ctx *ctx, ResSrArgs &a, int grid) {
    constexpr size_t smem_bytes = smem_of<KR, TW, T>();
    static_assert(smem_bytes <= 232448, "shared memory budget of one CTA");
    PANO_CUDA(cudaFuncSetAttribute(k_cg_resident_sr<KR, TW, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    void *kargs[] = {(void *)&a};
    PANO_CUDA(cudaLaunchCooperativeKernel((const void *)k_cg_resident_sr<KR, TW, T>, dim3((unsigned)grid), dim3(T), kargs, smem_bytes,
                                          ctx->stream));
    return pano_after_launch(ctx, "cg_resident_sr");
}

// the instantiated geometry with the fewest cells per CTA whose tiles fit on the SMs
bool plan_sr(size_t h, size_t w, int num_sms, Cfg *out, int *tiles_x, int *tiles_y) {
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

bool pano_cg_resident_sr_supported(pano_ctx *ctx, size_t h, size_t w) {
    Cfg c;
    int tx, ty;
    return h >= 1 && w >= 1 && plan_sr(h, w, ctx->num_sms, &c, &tx, &ty);
}

int pano_cg_resident_sr_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, size_t h, size_t w,
                               int max_iterations, double threshold, double timestep, RectI m) {
    Cfg c{0, 0, 0};
    int tx = 0, ty = 0;
    if (!plan_sr(h, w, ctx->num_sms, &c, &tx, &ty)) PANO_FAIL(PANO_ERR_INVALID, "cg_resident_sr: a %zux%zu grid does not fit on chip", h, w);
    ResSrArgs a;
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
    PANO_FAIL(PANO_ERR_INVALID, "cg_resident_sr: no kernel for KR=%d TW=%d T=%d", c.kr, c.tw, c.t);
}
