// pano_cg_sr.cu -- the pressure solve with ONE grid-wide reduction per iteration: a persistent, warp-specialised,
// TMA-pipelined conjugate-gradient kernel in the Chronopoulos-Gear arrangement (one CTA per SM).
//
// The reference iteration (pcg.rs:48-80) has two dependent reductions, z.s (for alpha) and r.r (for beta): two grid-wide
// (and, on several GPUs, cross-GPU) exchanges and two pipeline drains per iteration -- k_cg_stream pays ~13 us per
// iteration for them on one GPU and ~25 us on eight (profiles/r01_v7_multi_gpu_cg.md), on 82 us of streaming work per
// 1024 x 8192 slab.  The same iterates follow from ONE reduction of three values (Chronopoulos & Gear 1989):
//
//     r_0 = b, w_0 = A r_0, gamma_0 = r_0.r_0, delta_0 = w_0.r_0, beta_0 = 0, alpha_0 = gamma_0 / delta_0
//     pass i:   p_i = r_i + beta_i p_(i-1)            (the reference's search direction s, pcg.rs:72-77)
//               s_i = w_i + beta_i s_(i-1)            (= A p_i, the reference's z, by recurrence instead of a product)
//               x_(i+1) = x_i + alpha_i p_i           (pcg.rs:55)
//               r_(i+1) = r_i - alpha_i s_i           (pcg.rs:56)
//               w_(i+1) = A r_(i+1);  gamma' = r_(i+1).r_(i+1),  delta' = w_(i+1).r_(i+1),  max|r_(i+1)|     -> ONE exchange
//               stop if max|r_(i+1)| < threshold      (pcg.rs:58-63)
//               beta_(i+1) = gamma'/gamma;  alpha_(i+1) = gamma' / (delta' - beta_(i+1) gamma' / alpha_i)
//
// In exact arithmetic p, x, r are the reference's; in f64 they agree to ~1e-14 relative after 100 iterations on the smoke
// plume (scripts/cgcg_numerics.py, against the CPU oracle) -- far inside the +-2 iterations / 1e-5 of BASELINE.json.
//
// One pass = one sweep over the tiles.  w_(i+1) = A r_(i+1) needs r_(i+1) on the tile's one-cell halo, i.e. s_i and hence
// w_i = A r_i there, i.e. r_i on a TWO-cell halo: every CTA recomputes the ring with the same expressions in the same order
// as its owner (bit-identical), so no second barrier is needed.  r and s are double-buffered (the halo recomputation reads the
// neighbours' OLD values); p and x are updated in place.  HBM traffic per cell and iteration: read r, s, p, x, write r, s, p,
// x = 64 B, the same as k_cg_stream's two phases.
//
//   * a producer lane streams (TH+4)x(TW+4) boxes of r, (TH+2)x(TW+4) boxes of s and TH x TW boxes of p, x through a 3-stage
//     shared-memory ring with cp.async.bulk.tensor (TMA); out-of-range elements are zero-filled
//   * 8 consumer warps: a thread owns 2 adjacent columns x 4 rows and evaluates the first stage on 4 columns x 6 rows
//     around them (128-bit shared loads, everything else in registers; no inter-warp exchange, no block barrier per tile)
//   * tiles are claimed dynamically in fixed batches, reductions stay deterministic (as in k_cg_stream<true>)
//   * on a slab of a multi-GPU grid the two first / last rows of r_(i+1) and the first / last row of s_i go straight into
//     the neighbours' ghost rows over NVLink, and the reduction gets its cross-rank stage (pano_sm100.cuh)
#include "pano_cell_math.h"
#include "pano_sm100.cuh"

using namespace pano_sm100;

namespace {

constexpr int TH = 32, TW = 64;                 // tile (cells)
constexpr int kHX = 2;                          // column halo of every halo box (even: TMA needs 16-byte aligned box starts)
constexpr int BW = TW + 2 * kHX;                // 68
constexpr int RH = TH + 4, SH = TH + 2;         // rows of the r box (two-cell halo) and of the s box (one-cell halo)
constexpr int kRBoxBytes = RH * BW * 8;         // 19584
constexpr int kSBoxBytes = SH * BW * 8;         // 18496
constexpr int kIntBoxBytes = TH * TW * 8;       // 16384
constexpr int kRSlot = 19712, kSSlot = 18560;   // rounded up to multiples of 128
constexpr int kStageBytes = kRSlot + kSSlot + 2 * kIntBoxBytes;   // 71040
constexpr int kStages = 3;
constexpr int kConsumers = 256, kConsumerWarps = 8;
constexpr int kThreads = kConsumers + 32;       // + one producer warp
constexpr int kTailBytes = 8192;
constexpr int kSmemBytes = kStages * kStageBytes + kTailBytes;   // 221312
constexpr int kRows = TH / kConsumerWarps;      // 4

struct SrArgs {
    CUtensorMap m_b, m_r[2], m_s[2], m_p, m_x;   // b and r: RH x BW boxes; s: SH x BW; p, x: TH x TW
    double *x, *p;
    const double *b;
    double *r[2], *s[2];
    int h, w;
    double dt, threshold;
    int max_iter;
    RectI m;
    int tiles_x, tiles_y;
    ReduceUnit *units;
    unsigned long long seq_base;
    PanoCgControl *ctl;
    int zigzag;
    int fence_mode;
    int dynamic;
    int sm_exchange;              // 1: tile_sr_sm (r_(i+1) exchanged through shared memory), 0: tile_sr (recomputed per thread)
    const int *order;             // tile order: position -> tile (see pano_cg_tile_order), or null: row-major
    int slow_lo, slow_hi;         // the positions [slow_lo, slow_hi) hold the select-path tiles and the slab-edge tile rows
    int halo_mid;                 // slab with neighbours: the halo flags go out as soon as a CTA is past those positions
    int early_load;               // slab with neighbours: the first tiles of a pass are loaded on the GPU's "local done" (pano_sm100.cuh)
    unsigned long long *claim;    // 4 claim counters, used round-robin by the passes (zeroed at launch)
    ReduceUnit *tparts;           // [3 values][nbatch * kConsumerWarps] per-(batch, warp) partials, {value, pass tag}
    int batch_len, nbatch_long, nbatch;
    // ---- slab of a larger grid (multi-GPU); single GPU: row0 = 0, gy0 = 0, gh = h, no peers
    int row0;                     // array row of the first owned row (ghost rows sit above it)
    int gy0, gh;                  // global row of the first owned row; global grid height (walls)
    double *up_r[2], *up_s[2];    // upper neighbour's arrays at the row that mirrors MY row 0 (its first ghost row below its slab), or null
    double *dn_r[2], *dn_s[2];    // lower neighbour's arrays at the row that mirrors MY row h-2 (r) / h-1 (s), or null
    double *dn_x;                 // lower neighbour's x at the row that mirrors MY row h-1 (its projection reads it as p[y-1]), or null
    XRank xr;
    long long *dbg;               // optional per-section clock64 totals of CTA 0 (option "cg_profile"): 0 wait for the pass's first tile,
                                  // 1 tile loop, 2 batch-unit poll, 3 CTA reduction, 4 grid (+ cross-GPU) all-reduce
    long long *dbg_cta;           // optional [2][gridDim.x]: per-CTA totals of sections 0 and 1
};

struct Tail {                     // small shared-memory area behind the stage ring
    uint64_t full[kStages], empty[kStages], go;
    double vals[3][kMaxCtas];
    double out[4];
    double wsum[3][kConsumerWarps];
    int cont;                     // 1: producer continues with the next pass, 0: stop
    int tile[kStages];            // the tile staged in each ring slot, -1 = no more tiles in this pass
    int batch[kStages];           // its batch index if the tile is the LAST of its batch, else -1
    int behind[kStages];          // 1: the tile lies behind the slow / halo positions of the order in this pass's direction
    int ok;
};

__device__ __forceinline__ void consumer_sync() { named_bar_sync(1, kConsumers); }

__device__ __forceinline__ int slot_word(const int *p) {   // read by lane 0 (the lane that later releases the slot), broadcast
    int v = 0;
    if ((threadIdx.x & 31) == 0) v = *(const volatile int *)p;
    return __shfl_sync(0xffffffffu, v, 0);
}

// deterministic block reduction of {sum, sum, max} among the 256 consumer threads (one pair of barriers for the three values);
// results in every consumer thread
__device__ __forceinline__ void consumer_reduce3(double &v0, double &v1, double &v2, double (*wsum)[kConsumerWarps]) {
    v0 = warp_sum(v0);
    v1 = warp_sum(v1);
    v2 = warp_max(v2);
    consumer_sync();
    if ((threadIdx.x & 31) == 0) {
        const int wd = threadIdx.x >> 5;
        wsum[0][wd] = v0;
        wsum[1][wd] = v1;
        wsum[2][wd] = v2;
    }
    consumer_sync();
    double t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; ++i) {
        t0 += wsum[0][i];
        t1 += wsum[1][i];
        t2 = wsum[2][i] > t2 ? wsum[2][i] : t2;
    }
    v0 = t0; v1 = t1; v2 = t2;
}

struct Open4 { bool n, s, w, e; };
__device__ __forceinline__ Open4 open_edges(const SrArgs &a, int gy, int gx) {   // gy, gx GLOBAL; a cell outside the grid has no open edge
    Open4 o;
    const bool in = gy >= 0 && gy < a.gh && gx >= 0 && gx < a.w;
    o.n = in && gy > 0 && !in_rect(a.m, gy, gx);
    o.s = in && gy < a.gh - 1 && !in_rect(a.m, gy + 1, gx);
    o.w = in && gx > 0 && !in_rect(a.m, gy, gx);
    o.e = in && gx < a.w - 1 && !in_rect(a.m, gy, gx + 1);
    return o;
}

// kMode 0: the opening pass (R = b: gamma_0 = b.b, delta_0 = (A b).b, max|b|; nothing stored)
// kMode 1: pass 0 (R = b = r_0, s_(-1) = p_(-1) = 0, x_0 = 0)
// kMode 2: pass i > 0
// Consumer warp `wid` owns rows 4*wid .. 4*wid+3 of the tile, the lane columns 2*lane, 2*lane+1.
template <bool kFast, int kMode>
__device__ __forceinline__ void tile_sr(const SrArgs &a, const double *R, const double *S, const double *Pb, const double *Xb,
                                        double *r_dst, double *s_dst, double *r_up, double *r_dn, double *s_up, double *s_dn,
                                        int ty0, int tx0, double alpha, double beta, double &acc_g, double &acc_d, double &acc_max) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = 2 * lane, row0 = wid * kRows;
    const int gx = tx0 + col;                                  // global column of the thread's first own column
    const double nalpha = -alpha;
    // r box: cell (ty, tx) of the tile at [(ty + 2) * BW + tx + kHX]; this thread reads columns col-2 .. col+3 = box columns col .. col+5
    const double *pr = R + (row0 + 2) * BW + col;
    // s box: cell (ty, tx) at [(ty + 1) * BW + tx + kHX]; columns col-2 .. col+3 as well (col-1 .. col+2 are used)
    const double *ps = S + (row0 + 1) * BW + col;
    double ri[3][6];                                           // r_i, rows j-1, j, j+1
    double rn[3][4];                                           // r_(i+1), rows jj-1, jj, jj+1, columns col-1 .. col+2
    auto load6 = [&](const double *q, double *dst) {
        const double2 u = *reinterpret_cast<const double2 *>(q), v = *reinterpret_cast<const double2 *>(q + 2),
                      t = *reinterpret_cast<const double2 *>(q + 4);
        dst[0] = u.x; dst[1] = u.y; dst[2] = v.x; dst[3] = v.y; dst[4] = t.x; dst[5] = t.y;
    };
    if (kMode == 0) {
        // only w_0 = A b on the own cells: rows 0..3, own columns
        load6(pr - BW, ri[0]);
        load6(pr, ri[1]);
#pragma unroll
        for (int j = 0; j < kRows; ++j) {
            load6(pr + (j + 1) * BW, ri[2]);
            const int ly = ty0 + row0 + j, gy = a.gy0 + ly;
            double z0, z1;
            bool valid = true;
            if (kFast) {
                z0 = pano::laplacian_cell<double>(ri[1][2], ri[0][2], ri[2][2], ri[1][1], ri[1][3], true, true, true, true, a.dt);
                z1 = pano::laplacian_cell<double>(ri[1][3], ri[0][3], ri[2][3], ri[1][2], ri[1][4], true, true, true, true, a.dt);
            } else {
                valid = ly < a.h && gx < a.w;
                const Open4 o0 = open_edges(a, gy, gx), o1 = open_edges(a, gy, gx + 1);
                z0 = pano::laplacian_cell<double>(ri[1][2], ri[0][2], ri[2][2], ri[1][1], ri[1][3], o0.n, o0.s, o0.w, o0.e, a.dt);
                z1 = pano::laplacian_cell<double>(ri[1][3], ri[0][3], ri[2][3], ri[1][2], ri[1][4], o1.n, o1.s, o1.w, o1.e, a.dt);
            }
            if (valid) {
                const double c0 = ri[1][2], c1 = ri[1][3];
                acc_g = acc_g + c0 * c0;
                acc_g = acc_g + c1 * c1;
                acc_d = acc_d + z0 * c0;
                acc_d = acc_d + z1 * c1;
                const double a0 = c0 < 0 ? -c0 : c0, a1 = c1 < 0 ? -c1 : c1;
                acc_max = a0 > acc_max ? a0 : acc_max;
                acc_max = a1 > acc_max ? a1 : acc_max;
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) { ri[0][k] = ri[1][k]; ri[1][k] = ri[2][k]; }
        }
        return;
    }
    load6(pr - 2 * BW, ri[0]);
    load6(pr - BW, ri[1]);
#pragma unroll
    for (int j = -1; j <= kRows; ++j) {
        load6(pr + (j + 1) * BW, ri[2]);
        const int ly = ty0 + row0 + j, gy = a.gy0 + ly;
        // ---- first stage on row j, columns col-1 .. col+2:  w_i = A r_i,  s_i = w_i + beta s_(i-1),  r_(i+1) = r_i - alpha s_i
        double sp[4] = {0.0, 0.0, 0.0, 0.0};
        if (kMode == 2) {
            const double *q = ps + j * BW;
            const double2 u = *reinterpret_cast<const double2 *>(q), v = *reinterpret_cast<const double2 *>(q + 2),
                          t = *reinterpret_cast<const double2 *>(q + 4);
            sp[0] = u.y; sp[1] = v.x; sp[2] = v.y; sp[3] = t.x;
        }
        double si[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double wv;
            if (kFast) {
                wv = pano::laplacian_cell<double>(ri[1][k + 1], ri[0][k + 1], ri[2][k + 1], ri[1][k], ri[1][k + 2], true, true, true, true, a.dt);
            } else {
                const Open4 o = open_edges(a, gy, gx - 1 + k);
                wv = pano::laplacian_cell<double>(ri[1][k + 1], ri[0][k + 1], ri[2][k + 1], ri[1][k], ri[1][k + 2], o.n, o.s, o.w, o.e, a.dt);
            }
            si[k] = kMode == 2 ? wv + beta * sp[k] : wv;
            rn[2][k] = ri[1][k + 1] + nalpha * si[k];                          // pcg.rs:56
        }
        const bool own_row = j >= 0 && j < kRows;
        if (own_row) {
            const bool valid = kFast || (ly < a.h && gx < a.w);                 // the width is even: both columns are valid together
            if (valid) {
                const size_t gi = (size_t)(a.row0 + ly) * a.w + gx;
                const double2 sv = make_double2(si[1], si[2]), rv = make_double2(rn[2][1], rn[2][2]);
                *reinterpret_cast<double2 *>(s_dst + gi) = sv;
                *reinterpret_cast<double2 *>(r_dst + gi) = rv;
                // halo rows go straight into the neighbours' HBM (NVLink): two rows of r, one of s
                if (r_up && ly < 2) *reinterpret_cast<double2 *>(r_up + (size_t)ly * a.w + gx) = rv;
                if (r_dn && ly >= a.h - 2) *reinterpret_cast<double2 *>(r_dn + (size_t)(ly - (a.h - 2)) * a.w + gx) = rv;
                if (s_up && ly == 0) *reinterpret_cast<double2 *>(s_up + gx) = sv;
                if (s_dn && ly == a.h - 1) *reinterpret_cast<double2 *>(s_dn + gx) = sv;
            }
        }
        // ---- second stage on row jj = j - 1 (own rows, own columns): w_(i+1) = A r_(i+1), the dots, p and x
        if (j >= 1) {
            const int jj = j - 1;
            const int ly2 = ly - 1, gy2 = gy - 1;
            double z0, z1;
            bool valid = true;
            if (kFast) {
                z0 = pano::laplacian_cell<double>(rn[1][1], rn[0][1], rn[2][1], rn[1][0], rn[1][2], true, true, true, true, a.dt);
                z1 = pano::laplacian_cell<double>(rn[1][2], rn[0][2], rn[2][2], rn[1][1], rn[1][3], true, true, true, true, a.dt);
            } else {
                valid = ly2 < a.h && gx < a.w;
                const Open4 o0 = open_edges(a, gy2, gx), o1 = open_edges(a, gy2, gx + 1);
                z0 = pano::laplacian_cell<double>(rn[1][1], rn[0][1], rn[2][1], rn[1][0], rn[1][2], o0.n, o0.s, o0.w, o0.e, a.dt);
                z1 = pano::laplacian_cell<double>(rn[1][2], rn[0][2], rn[2][2], rn[1][1], rn[1][3], o1.n, o1.s, o1.w, o1.e, a.dt);
            }
            if (valid) {
                const size_t gi = (size_t)(a.row0 + ly2) * a.w + gx;
                const int ti = (row0 + jj) * TW + col;
                const double r0 = ri[0][2], r1 = ri[0][3];                      // r_i of row jj (the window has moved on by one row)
                double2 pn, xn;
                if (kMode == 1) {
                    pn = make_double2(r0, r1);                                  // p_0 = r_0 (pcg.rs:40-42)
                    xn = make_double2(alpha * r0, alpha * r1);                  // x_1 = 0 + alpha_0 p_0
                } else {
                    const double2 po = *reinterpret_cast<const double2 *>(Pb + ti), xo = *reinterpret_cast<const double2 *>(Xb + ti);
                    pn.x = r0 + beta * po.x;                                    // pcg.rs:72-77
                    pn.y = r1 + beta * po.y;
                    xn.x = xo.x + alpha * pn.x;                                 // pcg.rs:55
                    xn.y = xo.y + alpha * pn.y;
                }
                *reinterpret_cast<double2 *>(a.p + gi) = pn;
                *reinterpret_cast<double2 *>(a.x + gi) = xn;
                if (a.dn_x && ly2 == a.h - 1) *reinterpret_cast<double2 *>(a.dn_x + gx) = xn;   // every pass; the last one counts
                const double c0 = rn[1][1], c1 = rn[1][2];
                acc_g = acc_g + c0 * c0;
                acc_g = acc_g + c1 * c1;
                acc_d = acc_d + z0 * c0;
                acc_d = acc_d + z1 * c1;
                const double a0 = c0 < 0 ? -c0 : c0, a1 = c1 < 0 ? -c1 : c1;
                acc_max = a0 > acc_max ? a0 : acc_max;
                acc_max = a1 > acc_max ? a1 : acc_max;
            }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) { ri[0][k] = ri[1][k]; ri[1][k] = ri[2][k]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) { rn[0][k] = rn[1][k]; rn[1][k] = rn[2][k]; }
    }
}

// The same pass with the new residual exchanged through SHARED MEMORY instead of being recomputed four columns wide in every
// thread (tile_sr evaluates 24 first-stage cells for 8 cells it owns; ncu: 6.2 G warp instructions per solve at 4096^2 against
// 5.7 G for the two-reduction kernel, both latency-bound at 27 % issue utilisation, so time follows the instruction count):
//   phase A  every thread evaluates the first stage (w_i, s_i, r_(i+1)) on its 8 own cells, 192 threads one cell of the tile's
//            ring each; r_(i+1) goes to global memory (own cells) and, for all of them, into the staged s box IN PLACE -- the
//            thread that read s_(i-1) of a cell is the one that overwrites it
//   barrier  among the 256 consumers
//   phase B  second stage (w_(i+1) = A r_(i+1), dots, p, x) with the neighbours' r_(i+1) from shared memory
// The stage is handed back after a fence.proxy.async.shared::cta (generic writes, then TMA writes, to the same shared memory).
template <bool kFast, int kMode>
__device__ __forceinline__ void tile_sr_sm(const SrArgs &a, const double *R, double *S, const double *Pb, const double *Xb,
                                           double *r_dst, double *s_dst, double *r_up, double *r_dn, double *s_up, double *s_dn,
                                           int ty0, int tx0, double alpha, double beta, double &acc_g, double &acc_d, double &acc_max) {
    static_assert(kMode == 1 || kMode == 2, "the opening pass has no first stage");
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int col = 2 * lane, row0 = wid * kRows;
    const int gx = tx0 + col;
    const double nalpha = -alpha;
    // first stage at one cell: box indices of r (two-cell halo) and s (one-cell halo) of tile cell (ty, tx)
    auto first_stage = [&](int ty, int tx, bool open_all, double &s_out) -> double {
        const double *pr = R + (ty + 2) * BW + tx + kHX;
        const double c = pr[0], n = pr[-BW], so = pr[BW], we = pr[-1], ea = pr[1];
        double wv;
        if (open_all) {
            wv = pano::laplacian_cell<double>(c, n, so, we, ea, true, true, true, true, a.dt);
        } else {
            const Open4 o = open_edges(a, a.gy0 + ty0 + ty, tx0 + tx);
            wv = pano::laplacian_cell<double>(c, n, so, we, ea, o.n, o.s, o.w, o.e, a.dt);
        }
        s_out = kMode == 2 ? wv + beta * S[(ty + 1) * BW + tx + kHX] : wv;
        return c + nalpha * s_out;                                             // pcg.rs:56
    };
    // ---- phase A, own cells: rows row0 .. row0+3, columns col, col+1 (vertical register window over the r box)
    {
        const double *pr = R + (row0 + 2) * BW + col + kHX;
        double2 up = *reinterpret_cast<const double2 *>(pr - BW), c = *reinterpret_cast<const double2 *>(pr);
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
            const double2 dn = *reinterpret_cast<const double2 *>(pr + (k + 1) * BW);
            const double wv = pr[k * BW - 1], ev = pr[k * BW + 2];
            const int ly = ty0 + row0 + k, gy = a.gy0 + ly;
            double z0, z1;
            if (kFast) {
                z0 = pano::laplacian_cell<double>(c.x, up.x, dn.x, wv, c.y, true, true, true, true, a.dt);
                z1 = pano::laplacian_cell<double>(c.y, up.y, dn.y, c.x, ev, true, true, true, true, a.dt);
            } else {
                const Open4 o0 = open_edges(a, gy, gx), o1 = open_edges(a, gy, gx + 1);
                z0 = pano::laplacian_cell<double>(c.x, up.x, dn.x, wv, c.y, o0.n, o0.s, o0.w, o0.e, a.dt);
                z1 = pano::laplacian_cell<double>(c.y, up.y, dn.y, c.x, ev, o1.n, o1.s, o1.w, o1.e, a.dt);
            }
            double2 *ps = reinterpret_cast<double2 *>(S + (row0 + k + 1) * BW + col + kHX);
            double2 sv;
            if (kMode == 2) {
                const double2 so = *ps;
                sv.x = z0 + beta * so.x;
                sv.y = z1 + beta * so.y;
            } else {
                sv = make_double2(z0, z1);
            }
            double2 rv;
            rv.x = c.x + nalpha * sv.x;                                        // pcg.rs:56
            rv.y = c.y + nalpha * sv.y;
            *ps = rv;                                                          // r_(i+1) replaces s_(i-1) in the staged box
            if (kFast || (ly < a.h && gx < a.w)) {
                const size_t gi = (size_t)(a.row0 + ly) * a.w + gx;
                *reinterpret_cast<double2 *>(s_dst + gi) = sv;
                *reinterpret_cast<double2 *>(r_dst + gi) = rv;
                // halo rows go straight into the neighbours' HBM (NVLink): two rows of r, one of s
                if (r_up && ly < 2) *reinterpret_cast<double2 *>(r_up + (size_t)ly * a.w + gx) = rv;
                if (r_dn && ly >= a.h - 2) *reinterpret_cast<double2 *>(r_dn + (size_t)(ly - (a.h - 2)) * a.w + gx) = rv;
                if (s_up && ly == 0) *reinterpret_cast<double2 *>(s_up + gx) = sv;
                if (s_dn && ly == a.h - 1) *reinterpret_cast<double2 *>(s_dn + gx) = sv;
            }
            up = c;
            c = dn;
        }
    }
    // ---- phase A, the ring: row -1 (threads 0..63), row TH (64..127), column -1 (128..159), column TW (160..191)
    if (tid < 2 * TW + 2 * TH) {
        int ty, tx;
        if (tid < TW) { ty = -1; tx = tid; }
        else if (tid < 2 * TW) { ty = TH; tx = tid - TW; }
        else if (tid < 2 * TW + TH) { ty = tid - 2 * TW; tx = -1; }
        else { ty = tid - 2 * TW - TH; tx = TW; }
        double s_unused;
        const double rv = first_stage(ty, tx, kFast, s_unused);
        S[(ty + 1) * BW + tx + kHX] = rv;
    }
    consumer_sync();
    // ---- phase B: w_(i+1) = A r_(i+1) from the shared r_(i+1), the dots, p and x on the own cells
    {
        const double *pn = S + (row0 + 1) * BW + col + kHX;
        const double *pr = R + (row0 + 2) * BW + col + kHX;
        double2 up = *reinterpret_cast<const double2 *>(pn - BW), c = *reinterpret_cast<const double2 *>(pn);
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
            const double2 dn = *reinterpret_cast<const double2 *>(pn + (k + 1) * BW);
            const double wv = pn[k * BW - 1], ev = pn[k * BW + 2];
            const int ly = ty0 + row0 + k, gy = a.gy0 + ly;
            double z0, z1;
            bool valid = true;
            if (kFast) {
                z0 = pano::laplacian_cell<double>(c.x, up.x, dn.x, wv, c.y, true, true, true, true, a.dt);
                z1 = pano::laplacian_cell<double>(c.y, up.y, dn.y, c.x, ev, true, true, true, true, a.dt);
            } else {
                valid = ly < a.h && gx < a.w;
                const Open4 o0 = open_edges(a, gy, gx), o1 = open_edges(a, gy, gx + 1);
                z0 = pano::laplacian_cell<double>(c.x, up.x, dn.x, wv, c.y, o0.n, o0.s, o0.w, o0.e, a.dt);
                z1 = pano::laplacian_cell<double>(c.y, up.y, dn.y, c.x, ev, o1.n, o1.s, o1.w, o1.e, a.dt);
            }
            if (valid) {
                const size_t gi = (size_t)(a.row0 + ly) * a.w + gx;
                const int ti = (row0 + k) * TW + col;
                const double2 ro = *reinterpret_cast<const double2 *>(pr + k * BW);   // r_i
                double2 pv, xv;
                if (kMode == 1) {
                    pv = ro;                                                   // p_0 = r_0 (pcg.rs:40-42)
                    xv = make_double2(alpha * ro.x, alpha * ro.y);             // x_1 = 0 + alpha_0 p_0
                } else {
                    const double2 po = *reinterpret_cast<const double2 *>(Pb + ti), xo = *reinterpret_cast<const double2 *>(Xb + ti);
                    pv.x = ro.x + beta * po.x;                                 // pcg.rs:72-77
                    pv.y = ro.y + beta * po.y;
                    xv.x = xo.x + alpha * pv.x;                                // pcg.rs:55
                    xv.y = xo.y + alpha * pv.y;
                }
                *reinterpret_cast<double2 *>(a.p + gi) = pv;
                *reinterpret_cast<double2 *>(a.x + gi) = xv;
                if (a.dn_x && ly == a.h - 1) *reinterpret_cast<double2 *>(a.dn_x + gx) = xv;   // every pass; the last one counts
                acc_g = acc_g + c.x * c.x;
                acc_g = acc_g + c.y * c.y;
                acc_d = acc_d + z0 * c.x;
                acc_d = acc_d + z1 * c.y;
                const double a0 = c.x < 0 ? -c.x : c.x, a1 = c.y < 0 ? -c.y : c.y;
                acc_max = a0 > acc_max ? a0 : acc_max;
                acc_max = a1 > acc_max ? a1 : acc_max;
            }
            up = c;
            c = dn;
        }
    }
    // generic-proxy writes to the stage, then (once the slot is handed back) TMA writes to it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Position in the (un-reversed) tile order -> tile.  The order (pano_cg_tile_order, host) is: an eighth of the all-open tiles in
// row-major order; then the bulk of them with every tile that takes the select path (walls, obstacle, ragged edge) and the first
// and last tile row of a slab that has neighbours spread evenly in between; then the last eighth of the all-open tiles.  Odd passes
// walk it backwards (L2 reuse), so in EITHER direction a pass ends with all-open tiles.  Two reasons:
//   * select-path tiles are 2-3x slower; with row-major order one wall row ends every pass and every reduction waited ~10 us for
//     the CTAs that drew its tiles (option cg_profile: tile loops balanced to 2 %, yet 11.8 us in the all-reduce section per pass
//     on a 1024 x 8192 slab)
//   * on a slab of a multi-GPU grid the first and last tile row are mirrored into the neighbours' memory: a CTA can fence at system
//     scope and raise its halo flags as soon as it is past them, long before the pass ends, instead of putting 3 us of MEMBAR.SYS
//     and the flags' NVLink latency on the critical path of every reduction (halo_mid).
__device__ __forceinline__ int tile_at(const SrArgs &a, int pos) { return a.order ? __ldg(a.order + pos) : pos; }

// all open: every cell of the tile AND of its one-cell ring is an interior cell away from the walls and the obstacle
__host__ __device__ __forceinline__ bool tile_is_fast_hd(int h, int w, int gy0, int gh, const RectI &m, int ty0, int tx0) {
    const int g0 = gy0 + ty0;
    if (ty0 + TH > h) return false;                                              // ragged tile at the end of the slab
    if (g0 < 2 || g0 + TH > gh - 2 || tx0 < 2 || tx0 + TW > w - 2) return false;
    if (m.y1 > m.y0 && m.x1 > m.x0 && g0 - 1 < m.y1 && g0 + TH + 1 > m.y0 - 1 && tx0 - 1 < m.x1 && tx0 + TW + 1 > m.x0 - 1)
        return false;
    return true;
}
__device__ __forceinline__ bool tile_is_fast(const SrArgs &a, int ty0, int tx0) { return tile_is_fast_hd(a.h, a.w, a.gy0, a.gh, a.m, ty0, tx0); }

__device__ __forceinline__ bool tile_stores_remote(const SrArgs &a, int ty0) {
    return (ty0 == 0 && a.up_r[0] != nullptr) || (ty0 + TH >= a.h - 1 && a.dn_r[0] != nullptr);
}

// Pass k = 0 is the opening pass, pass k = 1 + i is iteration i.  Tiles are claimed from a global counter in batches of a
// fixed list (dynamic) or taken round-robin (static); see k_cg_stream for the determinism argument, which carries over.
__global__ void __launch_bounds__(kThreads, 1) k_cg_sr(const __grid_constant__ SrArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    Tail *tl = reinterpret_cast<Tail *>(smem + kStages * kStageBytes);
    const int tid = threadIdx.x, wid = tid >> 5;
    const int G = gridDim.x;
    const int ntiles = a.tiles_x * a.tiles_y;
    volatile unsigned int *err = &a.ctl->error;
    const bool dyn = a.dynamic != 0;
    const int n_my = (ntiles - (int)blockIdx.x + G - 1) / G;      // static lists: tiles blockIdx.x, +G, +2G, ...

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&tl->full[s], 1);
            mbar_init(&tl->empty[s], kConsumerWarps);
        }
        mbar_init(&tl->go, 1);
        tl->cont = 1;
        tl->ok = 1;
        fence_mbar_init();
    }
    __syncthreads();

    if (wid == kConsumerWarps) {
        // ============================================================ producer warp (one lane)
        if ((tid & 31) != 0) return;
        tma_prefetch_desc(&a.m_b);
        tma_prefetch_desc(&a.m_r[0]);
        tma_prefetch_desc(&a.m_r[1]);
        tma_prefetch_desc(&a.m_s[0]);
        tma_prefetch_desc(&a.m_s[1]);
        tma_prefetch_desc(&a.m_p);
        tma_prefetch_desc(&a.m_x);
        unsigned n = 0, ngo = 0;
        const unsigned long long M = (unsigned long long)a.nbatch + (unsigned long long)G;   // claims per pass (dynamic)
        bool stop = false;
        for (int k = 0; k <= a.max_iter && !stop; ++k) {
            const int it = k - 1;                                      // iteration of this pass (-1: the opening pass)
            bool exhausted = false;
            int b_idx = -1, b_next = 0, b_end = 0, jj = 0;
            bool behind = false;      // set by next_tile: is the tile behind the slow / halo positions in this pass's direction?
            bool safe = false;        // ... and: does it lie outside them (an all-open tile that reads no other GPU's rows)?
            auto next_tile = [&](int &t, int &closes) -> bool {
                if (!dyn) {
                    if (jj >= n_my) return false;
                    const int j = ((k & 1) && a.zigzag) ? n_my - 1 - jj : jj;
                    ++jj;
                    const int q = blockIdx.x + j * G;
                    t = tile_at(a, q);
                    behind = ((k & 1) && a.zigzag) ? q < a.slow_lo : q >= a.slow_hi;
                    safe = q < a.slow_lo || q >= a.slow_hi;
                    closes = -1;
                    return true;
                }
                if (b_next == b_end) {
                    if (exhausted) return false;
                    const unsigned long long v = atomicAdd(&a.claim[k & 3], 1ULL) - (unsigned long long)(k >> 2) * M;
                    if (v >= (unsigned long long)a.nbatch) { exhausted = true; return false; }   // exactly once per pass
                    b_idx = (int)v;
                    if (b_idx < a.nbatch_long) { b_next = b_idx * a.batch_len; b_end = b_next + a.batch_len; }
                    else { b_next = a.nbatch_long * a.batch_len + (b_idx - a.nbatch_long); b_end = b_next + 1; }
                }
                const int pos = b_next++;
                const int q = ((k & 1) && a.zigzag) ? ntiles - 1 - pos : pos;
                t = tile_at(a, q);
                behind = ((k & 1) && a.zigzag) ? q < a.slow_lo : q >= a.slow_hi;
                safe = q < a.slow_lo || q >= a.slow_hi;
                closes = b_next == b_end ? b_idx : -1;
                return true;
            };
            auto issue = [&](int t, int closes, bool bh) -> bool {
                const int st = n % kStages;
                if (!mbar_wait(&tl->empty[st], ((n / kStages) & 1) ^ 1, err)) return false;
                tl->tile[st] = t;
                tl->batch[st] = closes;
                tl->behind[st] = bh ? 1 : 0;
                const int tx0 = (t % a.tiles_x) * TW, ty0 = (t / a.tiles_x) * TH;
                unsigned char *base = smem + st * kStageBytes;
                uint64_t *bar = &tl->full[st];
                if (it <= 0) {                                          // opening pass and pass 0 read b only
                    mbar_arrive_expect_tx(bar, kRBoxBytes);
                    tma_load_2d(base, &a.m_b, bar, tx0 - kHX, a.row0 + ty0 - 2);
                } else {
                    mbar_arrive_expect_tx(bar, kRBoxBytes + kSBoxBytes + 2 * kIntBoxBytes);
                    tma_load_2d(base, &a.m_r[it & 1], bar, tx0 - kHX, a.row0 + ty0 - 2);
                    tma_load_2d(base + kRSlot, &a.m_s[(it - 1) & 1], bar, tx0 - kHX, a.row0 + ty0 - 1);
                    tma_load_2d(base + kRSlot + kSSlot, &a.m_p, bar, tx0, a.row0 + ty0);
                    tma_load_2d(base + kRSlot + kSSlot + kIntBoxBytes, &a.m_x, bar, tx0, a.row0 + ty0);
                }
                ++n;
                return true;
            };
            // The first tiles are claimed (atomic round trips through L2) while the previous pass is still being reduced ...
            int pre_t[kStages], pre_c[kStages], npre = 0, nissued = 0;
            bool pre_b[kStages], pre_s[kStages];
            const int want_pre = (k != 0 && a.early_load) ? kStages : 1;
            while (npre < want_pre) {
                int t, closes;
                if (!next_tile(t, closes)) break;
                pre_t[npre] = t; pre_c[npre] = closes; pre_b[npre] = behind; pre_s[npre] = safe;
                ++npre;
            }
            // ... but everything a pass READS was written by the previous pass, possibly by other CTAs: wait for its reduction.
            if (k != 0) {
                const unsigned n0 = n;
                if (a.early_load && npre > 0 && pre_s[0]) {
                    // Slab of a multi-GPU grid: tiles that read no other GPU's rows may be loaded as soon as THIS GPU has finished the
                    // previous pass ("local done", published by the root while the totals are still crossing NVLink)
                    const unsigned long long nprev = (unsigned long long)(k - 1);
                    const ReduceUnit *done = a.units + (nprev & 1) * kUnitsPerBank + 3 * kMaxCtas + 3;
                    double dummy;
                    if (!unit_poll(done, a.seq_base + nprev, dummy, err)) return;
                    fence_gpu(false);
                    fence_proxy_async();
                    while (nissued < npre && pre_s[nissued]) {
                        if (!issue(pre_t[nissued], pre_c[nissued], pre_b[nissued])) return;
                        ++nissued;
                    }
                }
                if (!mbar_wait(&tl->go, ngo & 1, err)) return;
                ++ngo;
                stop = !*(volatile int *)&tl->cont;
                fence_proxy_async();
                if (stop) {                                             // let the early loads land before the CTA leaves
                    for (unsigned m = n0; m < n; ++m)
                        if (!mbar_wait(&tl->full[m % kStages], (m / kStages) & 1, err)) return;
                    return;
                }
            }
            for (; nissued < npre; ++nissued)
                if (!issue(pre_t[nissued], pre_c[nissued], pre_b[nissued])) return;
            for (;;) {
                int t, closes;
                if (!next_tile(t, closes)) break;
                if (!issue(t, closes, behind)) return;
            }
            // end-of-pass marker for the consumers: an empty slot
            if (!mbar_wait(&tl->empty[n % kStages], ((n / kStages) & 1) ^ 1, err)) return;
            tl->tile[n % kStages] = -1;
            mbar_arrive(&tl->full[n % kStages]);
            ++n;
        }
        // the reduction that ends the last pass (nothing follows it)
        if (!stop) mbar_wait(&tl->go, ngo & 1, err);
        return;
    }

    // ================================================================ consumer warps
    unsigned n = 0;
    double gamma = 0, alpha = 0, beta = 0, rmax = 0, bmax = 0;
    int it = -1, applies = 0;
    bool converged = false, early = false;
    double red[3];

    const bool prof = a.dbg != nullptr && tid == 0;
    long long tprev = prof ? clock64() : 0;
    auto stamp = [&](int slot) {
        if (prof) {
            const long long t = clock64();
            if (blockIdx.x == 0) a.dbg[slot] += t - tprev;
            if (slot < 2) a.dbg_cta[slot * gridDim.x + blockIdx.x] += t - tprev;
            tprev = t;
        }
    };
    for (int k = 0; k <= a.max_iter; ++k) {
        it = k - 1;
        bool first_tile = true;
        double acc_g = 0, acc_d = 0, acc_max = 0;
        bool remote = false;      // uniform over the CTA: every consumer warp walks the same tiles
        const unsigned long long tag = a.seq_base + (unsigned long long)k + 1;
        double *r_dst = a.r[(it + 1) & 1], *s_dst = a.s[it & 1];
        double *r_up = a.up_r[(it + 1) & 1], *r_dn = a.dn_r[(it + 1) & 1], *s_up = a.up_s[it & 1], *s_dn = a.dn_s[it & 1];
        bool flags_pending = a.halo_mid != 0, flags_sent = false;
        for (;; ++n) {
            const int st = n % kStages;
            if (!mbar_wait(&tl->full[st], (n / kStages) & 1, err)) return;
            if (first_tile) { stamp(0); first_tile = false; }
            const int t = slot_word(&tl->tile[st]);
            if (t < 0) {                                    // end-of-pass marker: hand the slot back and leave
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&tl->empty[st]);
                ++n;
                break;
            }
            const int tx0 = (t % a.tiles_x) * TW, ty0 = (t / a.tiles_x) * TH;
            if (flags_pending && slot_word(&tl->behind[st]) != 0) {
                // this CTA's halo tiles of the pass are behind it (tiles are claimed in increasing order): vouch for them now
                consumer_sync();                   // every consumer warp's halo-row stores happen-before thread 0's fence
                if (tid == 0) {
                    fence_sys((a.fence_mode & kFenceLight) != 0);
                    send_halo_flags(&a.xr, (unsigned long long)k);
                }
                flags_pending = false;
                flags_sent = true;
                remote = false;
            }
            const double *R = reinterpret_cast<const double *>(smem + st * kStageBytes);
            double *S = reinterpret_cast<double *>(smem + st * kStageBytes + kRSlot);
            const double *Pb = reinterpret_cast<const double *>(smem + st * kStageBytes + kRSlot + kSSlot);
            const double *Xb = Pb + TH * TW;
            const bool fast = tile_is_fast(a, ty0, tx0);
            if (k == 0) {
                if (fast) tile_sr<true, 0>(a, R, S, Pb, Xb, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ty0, tx0, 0.0, 0.0, acc_g, acc_d, acc_max);
                else tile_sr<false, 0>(a, R, S, Pb, Xb, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ty0, tx0, 0.0, 0.0, acc_g, acc_d, acc_max);
            } else {
                remote = remote || tile_stores_remote(a, ty0);
                if (a.sm_exchange) {
                    if (k == 1) {
                        if (fast) tile_sr_sm<true, 1>(a, R, S, Pb, Xb, r_dst, s_dst, r_up, r_dn, s_up, s_dn, ty0, tx0, alpha, 0.0, acc_g, acc_d, acc_max);
                        else tile_sr_sm<false, 1>(a, R, S, Pb, Xb, r_dst, s_dst, r_up, r_dn, s_up, s_dn, ty0, tx0, alpha, 0.0, acc_g, acc_d, acc_max);
                    } else {
                        if (fast) tile_sr_sm<true, 2>(a, R, S, Pb, Xb, r_dst, s_dst, r_up, r_dn, s_up, s_dn, ty0, tx0, alpha, beta, acc_g, acc_d, acc_max);
                        else tile_sr_sm<false, 2>(a, R, S, Pb, Xb, r_dst, s_dst, r_up, r_dn, s_up, s_dn, ty0, tx0, alpha, beta, acc_g, acc_d, acc_max);
                    }
                } else if (k == 1) {
                    if (fast) tile_sr<true, 1>(a, R, S, Pb, Xb, r_dst, s_dst, r_up, r_dn, s_up, s_dn, ty0, tx0, alpha, 0.0, acc_g, acc_d, acc_max);
                    else tile_sr<false, 1>(a, R, S, Pb, Xb, r_dst, s_dst, r_up, r_dn, s_up, s_dn, ty0, tx0, alpha, 0.0, acc_g, acc_d, acc_max);
                } else {
                    if (fast) tile_sr<true, 2>(a, R, S, Pb, Xb, r_dst, s_dst, r_up, r_dn, s_up, s_dn, ty0, tx0, alpha, beta, acc_g, acc_d, acc_max);
                    else tile_sr<false, 2>(a, R, S, Pb, Xb, r_dst, s_dst, r_up, r_dn, s_up, s_dn, ty0, tx0, alpha, beta, acc_g, acc_d, acc_max);
                }
            }
            const int closes = dyn ? slot_word(&tl->batch[st]) : -1;
            if (closes >= 0) {                              // this warp's partials of the batch that ends here, then afresh
                const size_t plane = (size_t)a.nbatch * kConsumerWarps, u = (size_t)closes * kConsumerWarps + wid;
                const double p0 = warp_sum(acc_g), p1 = warp_sum(acc_d), p2 = warp_max(acc_max);
                if ((tid & 31) == 0) {
                    unit_store(a.tparts + u, p0, tag);
                    unit_store(a.tparts + plane + u, p1, tag);
                    unit_store(a.tparts + 2 * plane + u, p2, tag);
                }
                acc_g = 0; acc_d = 0; acc_max = 0;
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&tl->empty[st]);
        }
        stamp(1);
        if (dyn) {   // the units of this CTA's FIXED batch range, whoever computed them, in a fixed order
            const size_t plane = (size_t)a.nbatch * kConsumerWarps;
            const int c0 = (int)((long long)blockIdx.x * a.nbatch / G), c1 = (int)((long long)(blockIdx.x + 1) * a.nbatch / G);
            const ReduceUnit *base = a.tparts + (size_t)c0 * kConsumerWarps;
            for (int e = tid; e < (c1 - c0) * kConsumerWarps; e += kConsumers) {
                const ReduceUnit *us[3] = {base + e, base + plane + e, base + 2 * plane + e};
                double v[3] = {0.0, 0.0, 0.0};
                unit_poll_n(us, 3, tag, v, err);             // the three values' units in flight together
                acc_g = acc_g + v[0];
                acc_d = acc_d + v[1];
                acc_max = v[2] > acc_max ? v[2] : acc_max;
            }
        }
        stamp(2);
        double v0 = acc_g, v1 = acc_d, v2 = acc_max;
        consumer_reduce3(v0, v1, v2, tl->wsum);
        stamp(3);
        if (prof) {   // diagnostic: how long does the fence that the all-reduce starts with take on its own?
            fence_gpu(false);
            stamp(5);
        }
        if (!grid_allreduce_units(a.units, a.seq_base + (unsigned long long)k, (unsigned long long)k, 3, v0, v1, v2, 0x4u, tl->vals, tl->out,
                                  &tl->ok, &a.ctl->error, /*fenced=*/true, [] { consumer_sync(); }, red, &a.xr, NoWork(), nullptr,
                                  a.halo_mid ? (a.fence_mode | kFenceSysIfRemote) : a.fence_mode, remote, flags_sent)) {
            if (tid == 0) { tl->cont = 0; mbar_arrive(&tl->go); }
            return;
        }
        stamp(4);
        const double gamma_new = red[0], delta = red[1];
        rmax = red[2];
        if (k == 0) {
            bmax = rmax;                                       // pcg.rs:35
            if (bmax < a.threshold) early = true;              // pcg.rs:35-38
            gamma = gamma_new;                                 // pcg.rs:46
            alpha = gamma / delta;                             // pcg.rs:53 (s = r: z.s = (A r).r)
            beta = 0.0;
            if (early || a.max_iter == 0) {
                if (tid == 0) { tl->cont = 0; mbar_arrive(&tl->go); }
                break;
            }
        } else {
            ++applies;
            if (rmax < a.threshold) {                          // pcg.rs:58-63
                converged = true;
                if (tid == 0) { tl->cont = 0; mbar_arrive(&tl->go); }
                break;
            }
            beta = gamma_new / gamma;                          // pcg.rs:67-68
            alpha = gamma_new / (delta - beta * gamma_new / alpha);   // = sigma' / (z'.s'), by the Chronopoulos-Gear recurrence
            gamma = gamma_new;                                 // pcg.rs:79
            if (k == a.max_iter) {                             // the loop ran out
                if (tid == 0) { tl->cont = 0; mbar_arrive(&tl->go); }
                break;
            }
        }
        if (tid == 0) mbar_arrive(&tl->go);
    }

    // ------------------------------------------------------------------ epilogue (flat, once per solve)
    // after iteration `it` the residual r_(it+1) sits in buffer (it+1)&1; the caller's residual field is buffer 0
    const size_t ncell = (size_t)a.h * a.w, base = (size_t)a.row0 * a.w;
    const size_t stride = (size_t)G * kConsumers, i0 = base + (size_t)blockIdx.x * kConsumers + tid;
    if (early) {
        // x = 0 (pcg.rs:32) -- including the ghost row above, which the upper neighbour (same decision) would have mirrored
        const size_t first = (a.up_r[0] != nullptr && a.row0 > 0) ? base - (size_t)a.w : base;
        for (size_t i = first + (size_t)blockIdx.x * kConsumers + tid; i < base + ncell; i += stride) a.x[i] = 0.0;
    } else {
        const double *r_fin = a.r[(it + 1) & 1];
        const bool copy_r = r_fin != a.r[0];
        if (!converged || copy_r) {
            for (size_t i = i0; i < base + ncell; i += stride) {
                const double rv = __ldcg(r_fin + i);
                if (copy_r) a.r[0][i] = rv;
                // exhausted: the reference still performs the search update (pcg.rs:72-77) before leaving
                if (!converged) a.p[i] = rv + beta * __ldcg(a.p + i);
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

}  // namespace

// The tile order of tile_at: cached in the context (it depends only on the geometry), rebuilt and uploaded when that changes.
// margin: cells around a tile that must be all-open for its branch-free path (2 here, 1 in k_cg_stream, which shares the order).
int pano_cg_tile_order(pano_ctx *ctx, int h, int w, int gy0, int gh, RectI m, int tiles_x, int tiles_y, int th, int tw, int margin,
                       bool has_up, bool has_dn, const int **order_out, int *lo_out, int *hi_out) {
    const int ntiles = tiles_x * tiles_y;
    const long long key[12] = {h, w, gy0, gh, m.y0, m.y1, m.x0, m.x1, has_up ? 1 : 0, has_dn ? 1 : 0, margin, (long long)th * 65536 + tw};
    if (ctx->sr_order_n != ntiles || memcmp(ctx->sr_order_key, key, sizeof(key)) != 0) {
        auto all_open = [&](int ty0, int tx0) {
            const int g0 = gy0 + ty0;
            if (ty0 + th > h) return false;                                      // ragged tile at the end of the slab
            if (g0 < margin || g0 + th > gh - margin || tx0 < margin || tx0 + tw > w - margin) return false;
            if (m.y1 > m.y0 && m.x1 > m.x0 && g0 - (margin - 1) < m.y1 && g0 + th + (margin - 1) > m.y0 - 1 && tx0 - (margin - 1) < m.x1 &&
                tx0 + tw + (margin - 1) > m.x0 - 1)
                return false;
            return true;
        };
        std::vector<int> fast, slow;
        for (int t = 0; t < ntiles; ++t) {
            const int ty = t / tiles_x, ty0 = ty * th, tx0 = (t % tiles_x) * tw;
            const bool edge_row = (ty == 0 && has_up) || (ty == tiles_y - 1 && has_dn);
            (all_open(ty0, tx0) && !edge_row ? fast : slow).push_back(t);
        }
        // head and tail: an eighth of the all-open tiles each; in between the rest of them with the slow tiles spread evenly (all
        // slow tiles at once would leave HBM idle while every SM computes: measured, 8192^2 2.6 % slower than spread out)
        std::vector<int> order;
        order.reserve((size_t)ntiles);
        const size_t head = fast.size() / 8, mid_fast = fast.size() - 2 * head, mid = mid_fast + slow.size();
        order.insert(order.end(), fast.begin(), fast.begin() + head);
        size_t fi = head, si = 0;
        for (size_t i = 0; i < mid; ++i) {
            const bool take_slow = slow.size() && ((i + 1) * slow.size() / mid > i * slow.size() / mid);
            if (take_slow && si < slow.size()) order.push_back(slow[si++]);
            else if (fi < head + mid_fast) order.push_back(fast[fi++]);
            else order.push_back(slow[si++]);
        }
        order.insert(order.end(), fast.begin() + head + mid_fast, fast.end());
        if ((size_t)ntiles > ctx->sr_order_cap) {
            if (ctx->d_sr_order) {
                PANO_CUDA(cudaStreamSynchronize(ctx->stream));
                PANO_CUDA(cudaFree(ctx->d_sr_order));
                ctx->d_sr_order = nullptr;
                ctx->sr_order_cap = 0;
            }
            PANO_CUDA(cudaMalloc((void **)&ctx->d_sr_order, (size_t)ntiles * sizeof(int)));
            ctx->sr_order_cap = (size_t)ntiles;
        }
        // pageable host memory: the copy is staged before the call returns
        PANO_CUDA(cudaMemcpyAsync(ctx->d_sr_order, order.data(), (size_t)ntiles * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        memcpy(ctx->sr_order_key, key, sizeof(key));
        ctx->sr_order_n = ntiles;
        ctx->sr_order_lo = (int)head;
        ctx->sr_order_hi = (int)(head + mid);
    }
    *order_out = ctx->d_sr_order;
    *lo_out = ctx->sr_order_lo;
    *hi_out = ctx->sr_order_hi;
    return PANO_OK;
}

int pano_preload_cg_sr() {
    cudaFuncAttributes fa;
    PANO_CUDA(cudaFuncGetAttributes(&fa, k_cg_sr));
    PANO_CUDA(cudaFuncSetAttribute(k_cg_sr, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    return PANO_OK;
}

// x, b, r (caller's residual), p (caller's search), s0 (caller's auxiliary) plus two scratch arrays r1, s1 of the same
// shape.  Slab: all seven live in the rank's window with ghost rows (>= 2) around the owned rows.
int pano_cg_sr_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *p, double *s0, double *r1, double *s1, size_t h,
                      size_t w, int max_iterations, double threshold, double timestep, RectI m, const PanoCgSrSlab *slab) {
    if (max_iterations <= 0) PANO_FAIL(PANO_ERR_INVALID, "pano_cg_sr_launch: max_iterations = %d (callers handle the empty loop on the host)", max_iterations);
    PANO_CUDA(cudaFuncSetAttribute(k_cg_sr, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    static_assert(sizeof(Tail) <= kTailBytes, "Tail does not fit");
    static_assert(kSmemBytes <= 232448, "shared memory budget of one CTA");
    SrArgs a;
    memset(&a, 0, sizeof(a));
    const uint64_t pitch = (uint64_t)w * 8;
    const uint64_t rows = slab ? (uint64_t)slab->rows_total : (uint64_t)h;   // h = owned rows, rows = stored rows
    PANO_TRY(pano_make_tensor_map_2d(&a.m_b, b, 8, w, rows, pitch, BW, RH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_r[0], r, 8, w, rows, pitch, BW, RH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_r[1], r1, 8, w, rows, pitch, BW, RH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_s[0], s0, 8, w, rows, pitch, BW, SH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_s[1], s1, 8, w, rows, pitch, BW, SH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_p, p, 8, w, rows, pitch, TW, TH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_x, x, 8, w, rows, pitch, TW, TH));
    a.x = x; a.b = b; a.p = p;
    a.r[0] = r; a.r[1] = r1; a.s[0] = s0; a.s[1] = s1;
    a.h = (int)h; a.w = (int)w;
    a.dt = timestep; a.threshold = threshold; a.max_iter = max_iterations;
    a.m = m;
    a.tiles_x = ((int)w + TW - 1) / TW;
    a.tiles_y = ((int)h + TH - 1) / TH;
    a.ctl = ctx->d_cg;
    a.zigzag = pano_option(ctx, "cg_zigzag", 1) != 0;
    a.fence_mode = (int)pano_option(ctx, "cg_fence", 0);
    a.dbg = pano_option(ctx, "cg_profile", 0) ? ctx->d_cg->prof : nullptr;
    a.dbg_cta = reinterpret_cast<long long *>(ctx->d_partials);   // >= 12288 doubles (pano_ctx_create); zeroed below when profiling
    if (a.dbg) PANO_CUDA(cudaMemsetAsync(ctx->d_partials, 0, 2 * kMaxCtas * sizeof(long long), ctx->stream));
    a.row0 = 0; a.gy0 = 0; a.gh = (int)h;
    a.xr.rank = 0; a.xr.nranks = 1;
    int max_ctas = 0;
    if (slab) {
        a.row0 = slab->row0; a.gy0 = slab->gy0; a.gh = slab->gh;
        for (int i = 0; i < 2; ++i) {
            a.up_r[i] = slab->up_r[i]; a.up_s[i] = slab->up_s[i];
            a.dn_r[i] = slab->dn_r[i]; a.dn_s[i] = slab->dn_s[i];
        }
        a.dn_x = slab->dn_x;
        a.xr.rank = slab->rank; a.xr.nranks = slab->nranks;
        a.xr.seq_base = slab->xseq_base;
        a.xr.local = (ReduceUnit *)slab->xunits_local;
        for (int i = 0; i < kMaxRanks; ++i) a.xr.peer[i] = (ReduceUnit *)slab->xunits_peer[i];
        if (slab->nranks > 1 && pano_option(ctx, "cg_xflags", 1) != 0) {
            // halo flags (pano_sm100.cuh): the neighbours run the same launch code on their slabs, so their CTA counts
            // follow from the slab split
            auto ctas_of = [&](int rk) {
                const long long rws = (long long)slab->gh * (rk + 1) / slab->nranks - (long long)slab->gh * rk / slab->nranks;
                long long g = ctx->num_sms, nt = (long long)a.tiles_x * ((rws + TH - 1) / TH);
                if (slab->max_ctas > 0 && g > slab->max_ctas) g = slab->max_ctas;
                if (g > nt) g = nt;
                if (g > kMaxCtas) g = kMaxCtas;
                return (int)g;
            };
            a.xr.hflags = a.xr.local + kXUnitsTotal;
            if (slab->rank > 0) { a.xr.hflags_up = a.xr.peer[slab->rank - 1] + kXUnitsTotal; a.xr.g_up = ctas_of(slab->rank - 1); }
            if (slab->rank + 1 < slab->nranks) { a.xr.hflags_dn = a.xr.peer[slab->rank + 1] + kXUnitsTotal; a.xr.g_dn = ctas_of(slab->rank + 1); }
        }
        max_ctas = slab->max_ctas;
    }
    a.units = (ReduceUnit *)ctx->d_units;
    a.seq_base = (++ctx->launch_epoch) << 32;
    int G = ctx->num_sms;
    if (max_ctas > 0 && G > max_ctas) G = max_ctas;
    const int ntiles = a.tiles_x * a.tiles_y;
    if (G > ntiles) G = ntiles;
    if (G > kMaxCtas) G = kMaxCtas;
    PANO_TRY(pano_cg_control_reset(ctx));
    // dynamic tile scheduling ("cg_dynamic": 0 off, 1 on, -1 auto: from 24 tiles per CTA, as measured for k_cg_stream)
    const int64_t dyn_opt = pano_option(ctx, "cg_dynamic", -1);
    const bool dynamic = dyn_opt > 0 || (dyn_opt < 0 && ntiles >= 24 * G);
    a.dynamic = dynamic ? 1 : 0;
    a.sm_exchange = pano_option(ctx, "cg_sr_exchange", 1) != 0 ? 1 : 0;
    a.order = nullptr;
    a.slow_lo = a.slow_hi = 0;
    const int64_t om = pano_option(ctx, "cg_order_mid", -1);      // -1 auto (up to 96 tiles per CTA, see pano_cg_stream.cu), 0 off, 1 on
    const bool multi_gpu = slab && slab->nranks > 1;
    if (om > 0 || (om < 0 && (multi_gpu || (long long)a.tiles_x * a.tiles_y <= 96LL * ctx->num_sms))) {
        const bool multi = slab && slab->nranks > 1;
        PANO_TRY(pano_cg_tile_order(ctx, a.h, a.w, a.gy0, a.gh, a.m, a.tiles_x, a.tiles_y, TH, TW, 2, multi && slab->rank > 0,
                                    multi && slab->rank + 1 < slab->nranks, &a.order, &a.slow_lo, &a.slow_hi));
    }
    a.halo_mid = (a.order && slab && slab->nranks > 1 && a.xr.hflags != nullptr && pano_option(ctx, "cg_halo_mid", 1) != 0) ? 1 : 0;
    a.early_load = (a.order && slab && slab->nranks > 1 && a.xr.hflags != nullptr && pano_option(ctx, "cg_early_load", 1) != 0) ? 1 : 0;
    if (dynamic) {
        int bl = (int)pano_option(ctx, "cg_batch", 0);
        if (bl <= 0) bl = ntiles / (6 * G);
        if (bl > 8) bl = 8;
        if (bl < 1) bl = 1;
        a.batch_len = bl;
        a.nbatch_long = bl > 1 ? (int)((long long)ntiles * 4 / 5 / bl) : 0;
        a.nbatch = a.nbatch_long + (ntiles - a.nbatch_long * bl);
        const size_t units = 3 * (size_t)ntiles * kConsumerWarps;
        if (units > ctx->tparts_cap) {
            if (ctx->d_tparts) {
                PANO_CUDA(cudaStreamSynchronize(ctx->stream));
                PANO_CUDA(cudaFree(ctx->d_tparts));
                ctx->d_tparts = nullptr;
                ctx->tparts_cap = 0;
            }
            PANO_CUDA(cudaMalloc(&ctx->d_tparts, units * sizeof(ReduceUnit)));
            PANO_CUDA(cudaMemsetAsync(ctx->d_tparts, 0, units * sizeof(ReduceUnit), ctx->stream));   // tag 0 never matches
            ctx->tparts_cap = units;
        }
        if (!ctx->d_claim) PANO_CUDA(cudaMalloc((void **)&ctx->d_claim, 4 * sizeof(unsigned long long)));
        PANO_CUDA(cudaMemsetAsync(ctx->d_claim, 0, 4 * sizeof(unsigned long long), ctx->stream));
        a.tparts = (ReduceUnit *)ctx->d_tparts;
        a.claim = ctx->d_claim;
    }
    void *kargs[] = {(void *)&a};
    PANO_CUDA(cudaLaunchCooperativeKernel((const void *)k_cg_sr, dim3((unsigned)G), dim3(kThreads), kargs, kSmemBytes, ctx->stream));
    return pano_after_launch(ctx, dynamic ? "cg_sr(dynamic)" : "cg_sr");
}
