// pano_cg_cluster.cu -- the pressure solve for SMALL grids (up to ~40 k cells: the reference's shipped 128 x 128 example,
// BASELINE configs[0]) on ONE thread-block cluster.
//
// Same algorithm and per-cell arithmetic as the other CG kernels (pcg.rs:14-82 + the closure of dec_fluid.rs:100-119).
// What changes is where the two reductions of an iteration happen.  On a grid this small the work per iteration is a
// fraction of a microsecond, so the grid-wide all-reduce through L2 (~2 us each, two per iteration) is everything:
// the SM-resident kernel needs 5.3 us per iteration at 128^2.  Here eight CTAs form a cluster: x, r, z and the search
// direction live in shared memory for the whole solve, halo rows and reduction partials are written straight into the
// neighbours' shared memory (DSMEM), and the only synchronisation is the hardware cluster barrier -- TWO per
// iteration, no global memory traffic at all between the initial load of b and the final store of x, r, s.
//   slab decomposition along y over the 8 CTAs of the cluster, one halo row each way.  The halo of the search
//   direction is never exchanged: a CTA pushes its boundary rows of r together with its r.r partial (before beta is
//   known) and every CTA forms s' = r + beta s on its own rows AND on its two halo rows -- the same expression on the
//   same operands as the owner evaluates, hence the same bits -- so the search update needs no barrier of its own.
//   reductions: per-thread partials in a fixed order -> fixed shuffle tree -> one double per CTA, pushed into every
//   CTA's mailbox; every CTA adds the eight values in rank order, so all hold bit-identical alpha / beta / max|r|.
#include <cooperative_groups.h>

#include "pano_cell_math.h"
#include "pano_internal.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int CL = 8;                // CTAs per cluster (the portable maximum)
constexpr int CT = 512;              // threads per CTA
constexpr int kCap = 5120;           // cells per CTA
constexpr int kK = kCap / CT;        // cells per thread (10)
constexpr int kMaxW = 1024;          // widest grid (one halo row is w doubles)

struct ClArgs {
    double *x;
    const double *b;
    double *r, *s0;
    int h, w;
    double dt, threshold;
    int max_iter;
    RectI m;
    PanoCgControl *ctl;
};

struct ClSmem {
    double red[3][CL];     // mailbox: value k of CTA rank j (written by rank j through DSMEM)
    double wsum[3][CT / 32];
};
// S holds: halo row above | owned rows | halo row below.  RH holds the neighbours' boundary rows of r: [0] above, [1] below.

// deterministic block reductions: fixed shuffle tree per warp, fixed balanced tree over the 16 warp totals; one barrier
// (the scratch row is not reused before the next cluster barrier).  Result in every thread.
__device__ __forceinline__ double tree16_sum(const double *v) {
    return (((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]))) + (((v[8] + v[9]) + (v[10] + v[11])) + ((v[12] + v[13]) + (v[14] + v[15])));
}
__device__ __forceinline__ double tree16_max(const double *v) {
    double t = 0;
#pragma unroll
    for (int i = 0; i < CT / 32; ++i) t = v[i] > t ? v[i] : t;
    return t;
}
static_assert(CT / 32 == 16, "tree16_* assume 16 warps");
__device__ __forceinline__ double block_sum_cl(double v, double *wsum) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    return tree16_sum(wsum);
}
// sum of a and max of b in one pass (one barrier)
__device__ __forceinline__ void block_sum_max_cl(double a, double b, double *wa, double *wb, double &sa, double &mb) {
    a = warp_sum(a);
    b = warp_max(b);
    if ((threadIdx.x & 31) == 0) {
        wa[threadIdx.x >> 5] = a;
        wb[threadIdx.x >> 5] = b;
    }
    __syncthreads();
    sa = tree16_sum(wa);
    mb = tree16_max(wb);
}

// KA = cells per thread of the largest slab (compile time: the cell loops are fully unrolled, so the independent cells of a
// thread overlap their shared-memory and FP64 latencies, and no issue slot goes to cells a small grid does not have)
template <int KA>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CT, 1) k_cg_cluster(const ClArgs a) {
    extern __shared__ __align__(16) unsigned char raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), tid = threadIdx.x;
    const int w = a.w, h = a.h;
    const int y0 = (int)((long long)h * rank / CL), y1 = (int)((long long)h * (rank + 1) / CL);
    const int nloc = (y1 - y0) * w;

    ClSmem *sh = reinterpret_cast<ClSmem *>(raw);
    double *X = reinterpret_cast<double *>(raw + sizeof(ClSmem));
    double *R = X + kCap, *Z = R + kCap, *S = Z + kCap;   // S: halo row above | owned rows | halo row below
    double *So = S + w;                                   // owned cells
    double *RH = S + kCap + 2 * kMaxW;                    // r of the neighbours' boundary rows: [0, w) above, [w, 2w) below
    // neighbours that own rows (CTAs without rows are skipped when h < CL)
    int up = rank - 1, dn = rank + 1;
    while (up >= 0 && (long long)h * (up + 1) / CL == (long long)h * up / CL) --up;
    while (dn < CL && (long long)h * (dn + 1) / CL == (long long)h * dn / CL) ++dn;
    const bool has_up = nloc > 0 && up >= 0, has_dn = nloc > 0 && dn < CL;
    // where my first / last owned row goes: the upper neighbour's "below" slots, the lower neighbour's "above" slots
    double *up_s = nullptr, *up_r = nullptr, *dn_s = nullptr, *dn_r = nullptr;
    if (has_up) {
        const int urows = (int)((long long)h * (up + 1) / CL) - (int)((long long)h * up / CL);
        up_s = cluster.map_shared_rank(S, up) + (size_t)(urows + 1) * w;
        up_r = cluster.map_shared_rank(RH, up) + w;
    }
    if (has_dn) {
        dn_s = cluster.map_shared_rank(S, dn);
        dn_r = cluster.map_shared_rank(RH, dn);
    }

    // per-cell open-edge flags (bit 0 N, 1 S, 2 W, 3 E), fixed for the whole solve
    unsigned char open[KA];
#pragma unroll
    for (int k = 0; k < KA; ++k) {
        const int i = tid + k * CT;
        open[k] = 0;
        if (i < nloc) {
            const int y = y0 + i / w, x = i % w;
            const bool in = in_rect(a.m, y, x);
            open[k] = (unsigned char)((y > 0 && !in ? 1 : 0) | (y < h - 1 && !in_rect(a.m, y + 1, x) ? 2 : 0) | (x > 0 && !in ? 4 : 0) |
                                      (x < w - 1 && !in_rect(a.m, y, x + 1) ? 8 : 0));
        }
    }
    auto publish = [&](int slot, double t) {   // this CTA's total -> everybody's mailbox (read after the next cluster barrier)
        if (tid < CL) cluster.map_shared_rank(&sh->red[slot][0], tid)[rank] = t;
    };
    auto total = [&](int slot) {
        const double *v = sh->red[slot];
        return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    };
    auto total_max = [&](int slot) { double t = 0; for (int j = 0; j < CL; ++j) t = sh->red[slot][j] > t ? sh->red[slot][j] : t; return t; };
    static_assert(CL == 8, "total() is written for 8 CTAs");

    // ---- init: r = s = b, x = 0 (pcg.rs:32, 40-42); sigma = b.b, max|b| (:35, :46)
    for (int x = tid; x < w; x += CT) {          // halo rows nobody fills (grid ends); never used arithmetically (closed faces)
        S[x] = 0.0;
        So[nloc + x] = 0.0;
        RH[x] = 0.0;
        RH[w + x] = 0.0;
    }
    double acc0 = 0, acc1 = 0;
#pragma unroll
    for (int k = 0; k < KA; ++k) {
        const int i = tid + k * CT;
        if (i < nloc) {
            const double bv = a.b[(size_t)y0 * w + i];
            R[i] = bv;
            So[i] = bv;
            X[i] = 0.0;
            acc0 = acc0 + bv * bv;
            const double ab = bv < 0 ? -bv : bv;
            acc1 = ab > acc1 ? ab : acc1;
        }
    }
    double bb, bm;
    block_sum_max_cl(acc0, acc1, sh->wsum[1], sh->wsum[2], bb, bm);   // also: all of So is written (block barrier inside)
    cluster.sync();           // everybody's arrays exist and their end halos are zeroed before anybody pushes into them
    if (has_up) for (int x = tid; x < w; x += CT) up_s[x] = So[x];                    // s = b: first / last row to the neighbours
    if (has_dn) for (int x = tid; x < w; x += CT) dn_s[x] = So[nloc - w + x];
    publish(1, bb);           // slots 1 and 2: iteration 0 writes slot 0 first, possibly before a slow CTA has read these
    publish(2, bm);
    cluster.sync();
    double sigma = total(1);
    const double bmax = total_max(2);
    double rmax = bmax, beta = 0;
    int it = 0, applies = 0;
    bool converged = false;
    const bool early = bmax < a.threshold;                 // pcg.rs:35-38
    if (!early) {
        for (it = 0; it < a.max_iter; ++it) {
            if (it > 0) {
                // search = r + beta * search (:72-77) on my rows and on BOTH halo rows (from the neighbours' r rows pushed
                // before the last barrier): no exchange, no barrier of the cluster
#pragma unroll
                for (int k = 0; k < KA; ++k) {
                    const int i = tid + k * CT;
                    if (i < nloc) So[i] = R[i] + beta * So[i];
                }
                if (has_up) for (int x = tid; x < w; x += CT) S[x] = RH[x] + beta * S[x];
                if (has_dn) for (int x = tid; x < w; x += CT) So[nloc + x] = RH[w + x] + beta * So[nloc + x];
                __syncthreads();
            }
            // z = A s (:51), z.s
            double zs = 0;
#pragma unroll
            for (int k = 0; k < KA; ++k) {
                const int i = tid + k * CT;
                if (i < nloc) {
                    const unsigned o = open[k];
                    const double c = So[i];
                    const double z = pano::laplacian_cell<double>(c, So[i - w], So[i + w], So[i - 1], So[i + 1], (o & 1) != 0, (o & 2) != 0,
                                                                  (o & 4) != 0, (o & 8) != 0, a.dt);
                    Z[i] = z;
                    zs = zs + z * c;
                }
            }
            publish(0, block_sum_cl(zs, sh->wsum[0]));
            cluster.sync();
            const double alpha = sigma / total(0);         // :53
            ++applies;
            // x += alpha s, r -= alpha z (:55-56); r.r and max|r| (:58, :67)
            const double nalpha = -alpha;
            double rr = 0, rm = 0;
#pragma unroll
            for (int k = 0; k < KA; ++k) {
                const int i = tid + k * CT;
                if (i < nloc) {
                    X[i] = X[i] + alpha * So[i];
                    const double rn = R[i] + nalpha * Z[i];
                    R[i] = rn;
                    rr = rr + rn * rn;
                    const double ar = rn < 0 ? -rn : rn;
                    rm = ar > rm ? ar : rm;
                }
            }
            double rr_cta, rm_cta;
            block_sum_max_cl(rr, rm, sh->wsum[1], sh->wsum[2], rr_cta, rm_cta);   // block barrier inside: all of R is final
            if (has_up) for (int x = tid; x < w; x += CT) up_r[x] = R[x];                   // boundary rows of the new r
            if (has_dn) for (int x = tid; x < w; x += CT) dn_r[x] = R[nloc - w + x];
            publish(1, rr_cta);
            publish(2, rm_cta);
            cluster.sync();
            const double rr_all = total(1);
            rmax = total_max(2);
            if (rmax < a.threshold) {                      // :60-63
                converged = true;
                break;
            }
            beta = rr_all / sigma;                         // :68
            sigma = rr_all;                                // :79
        }
    }
    // ---- results: x always; residual and search untouched on the early-out, as in the reference
    if (early) {
        for (int i = tid; i < nloc; i += CT) a.x[(size_t)y0 * w + i] = 0.0;
    } else {
        for (int i = tid; i < nloc; i += CT) {
            const size_t g = (size_t)y0 * w + i;
            a.x[g] = X[i];
            a.r[g] = R[i];
            a.s0[g] = converged ? So[i] : R[i] + beta * So[i];   // exhausted: the loop's last pass still updates search (:72-77)
        }
    }
    if (rank == 0 && tid == 0) {
        a.ctl->iterations = early ? -1 : (converged ? it : a.max_iter);
        a.ctl->applies = applies;
        a.ctl->final_residual = rmax;
        a.ctl->rhs_max = bmax;
    }
    cluster.sync();   // nobody leaves while a peer may still write into its shared memory
}

size_t cluster_smem(size_t) { return sizeof(ClSmem) + (4 * (size_t)kCap + 4 * (size_t)kMaxW) * sizeof(double); }   // X, R, Z, S + 2 halo rows, 2 r rows

}  // namespace

bool pano_cg_cluster_supported(pano_ctx *ctx, size_t h, size_t w) {
    if (ctx->cc_major < 9 || h < 1 || w < 1 || w > kMaxW) return false;
    const size_t rows = (h + CL - 1) / CL;                 // the largest slab
    if (rows * w > (size_t)kCap) return false;
    return cluster_smem(w) <= ctx->smem_optin;
}

int pano_cg_cluster_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, size_t h, size_t w, int max_iterations,
                           double threshold, double timestep, RectI m) {
    if (!pano_cg_cluster_supported(ctx, h, w)) PANO_FAIL(PANO_ERR_INVALID, "cg_cluster: a %zux%zu grid does not fit one cluster", h, w);
    ClArgs a{x, b, r, s0, (int)h, (int)w, timestep, threshold, max_iterations, m, ctx->d_cg};
    const size_t smem = cluster_smem(w);
    const int ka = (int)((((h + CL - 1) / CL) * w + CT - 1) / CT);       // cells per thread of the largest slab, 1..kK
    PANO_TRY(pano_cg_control_reset(ctx));
#define PANO_CL_CASE(K)                                                                                              \
    case K:                                                                                                          \
        PANO_CUDA(cudaFuncSetAttribute(k_cg_cluster<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        k_cg_cluster<K><<<CL, CT, smem, ctx->stream>>>(a);                                                           \
        break;
    switch (ka) {
        PANO_CL_CASE(1) PANO_CL_CASE(2) PANO_CL_CASE(3) PANO_CL_CASE(4) PANO_CL_CASE(5)
        PANO_CL_CASE(6) PANO_CL_CASE(7) PANO_CL_CASE(8) PANO_CL_CASE(9) PANO_CL_CASE(10)
        default: PANO_FAIL(PANO_ERR_INVALID, "cg_cluster: %d cells per thread", ka);
    }
#undef PANO_CL_CASE
    static_assert(kK == 10, "the dispatch above lists 1..kK");
    return pano_after_launch(ctx, "cg_cluster");
}

int pano_preload_cg_cluster() {
    cudaFuncAttributes fa;
    PANO_CUDA(cudaFuncGetAttributes(&fa, k_cg_cluster<4>));
    return PANO_OK;
}
