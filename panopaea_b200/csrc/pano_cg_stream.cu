// pano_cg_stream.cu -- the pressure solve for grids that do not fit on chip: a persistent,
// warp-specialised, TMA-pipelined conjugate-gradient kernel (one CTA per SM).
//
// Same algorithm and arithmetic as k_cg_generic (pano_cg.cu; pcg.rs:14-82 + dec_fluid.rs:100-119):
//   P1: s' = r + beta*s on every tile INCLUDING its one-cell halo, z = A s' (not stored), z.s'
//   P2: z recomputed from s', x += alpha s', r -= alpha z, r.r, max|r|
// What changes is the data movement:
//   * a producer warp streams (TH+2)x(TW+4) halo boxes of r / s and TH x TW boxes of r / x into a
//     4-stage shared-memory ring with cp.async.bulk.tensor (TMA), completion on mbarriers;
//     out-of-range box elements are zero-filled by the TMA unit, which is exactly what a closed
//     (Neumann) wall edge needs
//   * 8 consumer warps run the stencil out of shared memory: each thread owns 2 adjacent columns
//     x 4 rows with a vertical register window, 128-bit shared loads and 128-bit global stores;
//     s' is formed in registers (the staged boxes are never written, so no proxy fence is needed
//     before the TMA unit refills a stage)
//   * boxes that do not depend on the other CTAs' work of the current phase (s_old in P1, r and x
//     in P2) are prefetched BEFORE the grid-wide reduction completes, so the pipeline does not
//     drain at phase boundaries
//   * tiles that touch neither a wall nor the obstacle take a branch-free path
//   * the two grid-wide reductions per iteration use the publish+poll all-reduce of
//     pano_sm100.cuh: one L2 round trip, deterministic, no atomics
//   * P2 walks the CTA's tiles in reverse order, so the tail of P1's footprint is still in the
//     126 MB L2 when P2 starts (and vice versa)
// HBM traffic per cell and iteration: 24 B (P1) + 40 B (P2) = 64 B.
#include "pano_cell_math.h"
#include "pano_sm100.cuh"

using namespace pano_sm100;

namespace {

constexpr int TH = 32, TW = 64;                 // tile (cells)
// Halo box: one row above/below, TWO columns left/right.  Measured on B200: cp.async.bulk.tensor
// traps ("illegal instruction") unless coordinate0 * sizeof(element) is a multiple of 16 bytes, so
// an f64 box must start on an even column; tx0 - 2 is even, tx0 - 1 is not.  The interior then
// starts at box column kHX = 2, which also keeps it 16-byte aligned in shared memory.
constexpr int kHX = 2;
constexpr int BH = TH + 2, BW = TW + 2 * kHX;   // 34 x 68
constexpr int kBoxElems = BH * BW;              // 2312 (even)
constexpr int kHaloBoxBytes = kBoxElems * 8;    // 18496 = TMA transaction size
constexpr int kHaloSlot = 18560;                // rounded up to a multiple of 128
constexpr int kIntBoxBytes = TH * TW * 8;       // 16384
constexpr int kStageBytes = 2 * kHaloSlot + kIntBoxBytes;   // 53504
constexpr int kStages = 4;
constexpr int kConsumers = 256, kConsumerWarps = 8;
constexpr int kThreads = kConsumers + 32;       // + one producer warp
constexpr int kTailBytes = 8192;
constexpr int kSmemBytes = kStages * kStageBytes + kTailBytes;

struct StreamArgs {
    CUtensorMap m_b_halo, m_r_halo, m_r_int, m_s0_halo, m_s1_halo, m_x_int;
    double *x;
    const double *b;
    double *r, *s0, *s1;
    int h, w;
    double dt, threshold;
    int max_iter;
    RectI m;
    int tiles_x, tiles_y;
    ReduceUnit *units;            // kUnitsTotal units (pano_sm100.cuh)
    unsigned long long seq_base;
    PanoCgControl *ctl;
    int zigzag;
    const int *order;             // tile order: position -> tile (pano_cg_sr.cu: pano_cg_tile_order), or null: row-major
    int halo_first;               // slab with neighbours: halo tile rows first in every phase (see slot_tile)
    int fence_mode;               // kFenceLight | kFenceSysIfRemote (pano_sm100.cuh), option "cg_fence"
    // ---- dynamic tile scheduling (k_cg_stream<true>)
    unsigned long long *claim;    // 4 claim counters, used round-robin by the phases (zeroed at launch)
    ReduceUnit *tparts;           // [3 values][nbatch * kConsumerWarps] per-(batch, warp) partials, {value, phase tag}
    int batch_len, nbatch_long, nbatch;   // claim unit: the first nbatch_long batches have batch_len tiles, the rest one tile
    // ---- slab of a larger grid (multi-GPU); single GPU: row0 = 0, gy0 = 0, gh = h, no peers
    int row0;                     // array row of the first owned row (ghost rows sit above it)
    int gy0, gh;                  // global row of the first owned row; global grid height (walls)
    double *up_r, *up_s0, *up_s1; // upper neighbour's ghost row BELOW its slab (receives my first row), or null
    double *dn_r, *dn_s0, *dn_s1; // lower neighbour's ghost row ABOVE its slab (receives my last row), or null
    XRank xr;
    long long *dbg;               // optional per-section clock64 totals of CTA 0 (option "cg_profile")
    long long *dbg_cta;           // optional [2][gridDim.x]: per-CTA clock64 totals spent in the P1 / P2 tile loops
};

struct Tail {                     // small shared-memory area behind the stage ring
    uint64_t full[kStages], empty[kStages], go;
    double vals[3][kMaxCtas];
    double out[4];
    double wsum[3][kConsumerWarps];
    int cont;                     // 1: producer continues with the next phase, 0: stop
    int tile[kStages];            // dynamic scheduling: the tile staged in each ring slot, -1 = no more tiles in this phase
    int batch[kStages];           // its batch index if the tile is the LAST of its batch (partials are published then), else -1
    int ok;
};

__device__ __forceinline__ void consumer_sync() { named_bar_sync(1, kConsumers); }

// A ring-slot word written by the producer lane (dynamic scheduling): read by lane 0 and broadcast, so that the thread that
// reads it is the thread that later hands the slot back with mbarrier.arrive on `empty` -- the write-after-read edge then
// needs no cumulativity argument (and compute-sanitizer's racecheck can follow it).
__device__ __forceinline__ int slot_word(const int *p) {
    int v = 0;
    if ((threadIdx.x & 31) == 0) v = *(const volatile int *)p;
    return __shfl_sync(0xffffffffu, v, 0);
}

// deterministic block reduction among the 256 consumer threads; result in every consumer thread
__device__ __forceinline__ double consumer_sum(double v, double *wsum) {
    v = warp_sum(v);
    consumer_sync();
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    consumer_sync();
    double t = 0;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; ++i) t += wsum[i];
    return t;
}
__device__ __forceinline__ double consumer_max(double v, double *wsum) {
    v = warp_max(v);
    consumer_sync();
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    consumer_sync();
    double t = 0;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; ++i) t = wsum[i] > t ? wsum[i] : t;
    return t;
}

// Grid-wide all-reduce of up to three block totals, called by all consumer threads (pano_sm100.cuh).
// fenced: the tiles this CTA stored to global memory must be visible to the others afterwards.
// remote: this CTA stored rows into a neighbour GPU's memory since the previous exchange.
__device__ __forceinline__ bool grid_allreduce(const StreamArgs &a, Tail *tl, unsigned long long n, int nvals, double v0,
                                               double v1, double v2, unsigned max_mask, double *out, bool remote, bool flags_sent) {
    return grid_allreduce_units(a.units, a.seq_base + n, n, nvals, v0, v1, v2, max_mask, tl->vals, tl->out,
                                &tl->ok, &a.ctl->error, /*fenced=*/true, [] { consumer_sync(); }, out, &a.xr, NoWork(), nullptr,
                                a.fence_mode, remote && !flags_sent, flags_sent);
}
// Static tile lists with halo_first: a CTA's halo tiles are its first `nh` tiles of every phase.  Right after the last of
// them the CTA fences at system scope and raises its halo flags for the reduction that ends the phase (exchange number
// `nred`), ~3 us of one warp in the middle of the phase instead of all warps' at its end.
__device__ __forceinline__ void early_halo_flags(const StreamArgs &a, unsigned long long nred) {
    consumer_sync();                       // every consumer warp's halo-row stores happen-before thread 0's fence
    if (threadIdx.x == 0) {
        fence_sys((a.fence_mode & kFenceLight) != 0);
        send_halo_flags(&a.xr, nred);
    }
}
// does a tile touch a slab edge whose rows are mirrored into a neighbour's ghost row?
__device__ __forceinline__ bool tile_stores_remote(const StreamArgs &a, int ty0) {
    return (ty0 == 0 && a.up_r != nullptr) || (ty0 + TH >= a.h && a.dn_r != nullptr);
}

// ------------------------------------------------------------------------------ tile kernels
// Consumer warp `wid` owns rows kRows*wid .. +kRows-1 of the tile; lane owns columns 2*lane, 2*lane+1.
// S / R: halo boxes (BH x BW), cell (ty, tx) at [(ty+1)*BW + tx + kHX]; 16-byte aligned pairs.
constexpr int kRows = TH / kConsumerWarps;   // 4

struct Open4 { bool n, s, w, e; };
__device__ __forceinline__ Open4 open_edges(const StreamArgs &a, int gy, int gx) {
    Open4 o;
    o.n = gy > 0 && !in_rect(a.m, gy, gx);          // gy is the GLOBAL row
    o.s = gy < a.gh - 1 && !in_rect(a.m, gy + 1, gx);
    o.w = gx > 0 && !in_rect(a.m, gy, gx);
    o.e = gx < a.w - 1 && !in_rect(a.m, gy, gx + 1);
    return o;
}

// P1: s' = r + beta*s (kFirst: s' = b, nothing stored), z = A s', accumulate z.s' (+ b.b, max|b|)
template <bool kFast, bool kFirst>
__device__ __forceinline__ void tile_p1(const StreamArgs &a, const double *S, const double *R, double *s_dst, double *s_up,
                                        double *s_dn, int ty0, int tx0, double beta, double &acc_zs, double &acc_bb,
                                        double &acc_bmax) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = 2 * lane, row0 = wid * kRows;
    const int gx = tx0 + col;
    const double *ps = S + (row0 + 1) * BW + col + kHX;
    const double *pr = R + (row0 + 1) * BW + col + kHX;
    auto sp2 = [&](int off) -> double2 {
        double2 sv = *reinterpret_cast<const double2 *>(ps + off);
        if (kFirst) return sv;
        const double2 rv = *reinterpret_cast<const double2 *>(pr + off);
        sv.x = rv.x + beta * sv.x;
        sv.y = rv.y + beta * sv.y;
        return sv;
    };
    auto sp1 = [&](int off) -> double {
        const double sv = ps[off];
        if (kFirst) return sv;
        return pr[off] + beta * sv;
    };
    double2 up = sp2(-BW), c = sp2(0);
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const double2 dn = sp2((k + 1) * BW);
        const double wv = sp1(k * BW - 1), ev = sp1(k * BW + 2);
        const int ly = ty0 + row0 + k, gy = a.gy0 + ly;   // local (owned) row, global row
        double z0, z1;
        bool valid = true;
        if (kFast) {
            z0 = pano::laplacian_cell<double>(c.x, up.x, dn.x, wv, c.y, true, true, true, true, a.dt);
            z1 = pano::laplacian_cell<double>(c.y, up.y, dn.y, c.x, ev, true, true, true, true, a.dt);
        } else {
            valid = ly < a.h && gx < a.w;          // the width is even: both columns are valid together
            const Open4 o0 = open_edges(a, gy, gx), o1 = open_edges(a, gy, gx + 1);
            z0 = pano::laplacian_cell<double>(c.x, up.x, dn.x, wv, c.y, o0.n, o0.s, o0.w, o0.e, a.dt);
            z1 = pano::laplacian_cell<double>(c.y, up.y, dn.y, c.x, ev, o1.n, o1.s, o1.w, o1.e, a.dt);
        }
        if (valid) {
            acc_zs = acc_zs + z0 * c.x;
            acc_zs = acc_zs + z1 * c.y;
            if (kFirst) {
                const double a0 = c.x < 0 ? -c.x : c.x, a1 = c.y < 0 ? -c.y : c.y;
                acc_bmax = a0 > acc_bmax ? a0 : acc_bmax;
                acc_bmax = a1 > acc_bmax ? a1 : acc_bmax;
                acc_bb = acc_bb + c.x * c.x;
                acc_bb = acc_bb + c.y * c.y;
            } else {
                *reinterpret_cast<double2 *>(s_dst + (size_t)(a.row0 + ly) * a.w + gx) = c;
                if (ly == 0 && s_up) *reinterpret_cast<double2 *>(s_up + gx) = c;          // halo rows go straight into
                if (ly == a.h - 1 && s_dn) *reinterpret_cast<double2 *>(s_dn + gx) = c;    // the neighbours' HBM (NVLink)
            }
        }
        up = c;
        c = dn;
    }
}

// P2: z recomputed from s', x += alpha s', r -= alpha z, accumulate r.r and max|r|
template <bool kFast, bool kFirst>
__device__ __forceinline__ void tile_p2(const StreamArgs &a, const double *S, const double *R, const double *X, int ty0,
                                        int tx0, double alpha, double &acc_rr, double &acc_rmax) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = 2 * lane, row0 = wid * kRows;
    const int gx = tx0 + col;
    const double nalpha = -alpha;
    const double *ps = S + (row0 + 1) * BW + col + kHX;
    double2 up = *reinterpret_cast<const double2 *>(ps - BW), c = *reinterpret_cast<const double2 *>(ps);
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const double2 dn = *reinterpret_cast<const double2 *>(ps + (k + 1) * BW);
        const double wv = ps[k * BW - 1], ev = ps[k * BW + 2];
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
            double2 xo, ro;
            if (kFirst) {
                xo = make_double2(0.0, 0.0);
                ro = c;                                          // iteration 0: r = s = b (pcg.rs:40-42)
            } else {
                xo = *reinterpret_cast<const double2 *>(X + ti);
                ro = *reinterpret_cast<const double2 *>(R + ti);
            }
            double2 xn, rn;
            xn.x = xo.x + alpha * c.x;                           // pcg.rs:55
            xn.y = xo.y + alpha * c.y;
            rn.x = ro.x + nalpha * z0;                           // pcg.rs:56
            rn.y = ro.y + nalpha * z1;
            *reinterpret_cast<double2 *>(a.x + gi) = xn;
            *reinterpret_cast<double2 *>(a.r + gi) = rn;
            if (ly == 0 && a.up_r) *reinterpret_cast<double2 *>(a.up_r + gx) = rn;
            if (ly == a.h - 1 && a.dn_r) *reinterpret_cast<double2 *>(a.dn_r + gx) = rn;
            if (kFirst) *reinterpret_cast<double2 *>(a.s0 + gi) = c;
            const double a0 = rn.x < 0 ? -rn.x : rn.x, a1 = rn.y < 0 ? -rn.y : rn.y;
            acc_rmax = a0 > acc_rmax ? a0 : acc_rmax;
            acc_rmax = a1 > acc_rmax ? a1 : acc_rmax;
            acc_rr = acc_rr + rn.x * rn.x;
            acc_rr = acc_rr + rn.y * rn.y;
        }
        up = c;
        c = dn;
    }
}

__device__ __forceinline__ bool tile_is_fast(const StreamArgs &a, int ty0, int tx0) {
    const int g0 = a.gy0 + ty0;   // global row of the tile's first row
    if (ty0 + TH > a.h) return false;                                            // ragged tile at the end of the slab
    if (g0 < 1 || g0 + TH > a.gh - 1 || tx0 < 1 || tx0 + TW > a.w - 1) return false;   // touches a wall
    if (a.m.y1 > a.m.y0 && a.m.x1 > a.m.x0 && g0 < a.m.y1 && g0 + TH > a.m.y0 - 1 && tx0 < a.m.x1 && tx0 + TW > a.m.x0 - 1)
        return false;                                                            // touches the obstacle
    return true;
}

// Which tile a CTA works on in its jj-th step of a phase.  Single GPU: tiles blockIdx.x, +G, +2G, ... in P1 and the
// same list backwards in P2 (zig-zag: the tail of one phase is still in L2 when the next starts).
// Option "cg_halo_first" (slab of a multi-GPU grid, default OFF): the first and the last tile row are renumbered to the
// FRONT of the list and taken first in BOTH phases, so that the rows stored into the neighbours' memory over NVLink
// are long acknowledged when the phase's system-scope fence and reduction come.  Measured on B200s, 8192^2: no gain
// (2 GPUs: 1605 vs 1602-1612 Mcell-steps/s; 8 GPUs: 5487 vs 5634) -- the posted NVLink stores are not what the
// cross-GPU reductions wait for; kept as a switch for other topologies.
__device__ __forceinline__ int slot_tile(const StreamArgs &a, int phase, int jj, int n_my, int G, int ntiles) {
    const bool halo_first = a.halo_first != 0;
    int j = jj;
    if (phase == 1 && a.zigzag) {
        int nh = 0;                                   // leading steps of this CTA that are halo tiles
        if (halo_first) {
            const int nhalo = a.tiles_y > 1 ? 2 * a.tiles_x : a.tiles_x;
            nh = nhalo > (int)blockIdx.x ? (nhalo - (int)blockIdx.x + G - 1) / G : 0;
            if (nh > n_my) nh = n_my;
        }
        j = jj < nh ? jj : n_my - 1 - (jj - nh);
    }
    const int q = blockIdx.x + j * G;
    if (!halo_first && a.order) return __ldg(a.order + q);
    if (!halo_first || a.tiles_y <= 2) return q;
    if (q < a.tiles_x) return q;                                              // first tile row
    if (q < 2 * a.tiles_x) return ntiles - 2 * a.tiles_x + q;                 // last tile row
    return q - a.tiles_x;                                                     // rows 1 .. tiles_y - 2
}

// kDyn = false: every CTA owns a fixed list of tiles (slot_tile).
// kDyn = true:  tiles are claimed from a global counter as ring slots free up, so SMs that stream faster take more
//   tiles (measured with the fixed lists at 4096^2: the slowest CTA needs 15-20 % longer than the average one in
//   every phase, and everybody waits for it at the reduction).  The reductions stay deterministic and independent of
//   who computed what: the unit of claiming is a BATCH of consecutive tiles from a fixed list (long batches first, single
//   tiles for the last fifth of a phase, so that the finish is balanced to one tile while only a few claims and
//   reductions are paid per CTA); every consumer warp publishes its partial of every batch as a 16-byte
//   {value, phase tag} unit, and CTA c adds the units of the FIXED batch range c*nbatch/G .. (c+1)*nbatch/G in a fixed
//   order before the usual grid all-reduce.  Each phase instance k uses claim counter k mod 4; a counter only grows, by
//   nbatch + G per use (every CTA makes exactly one failing claim), so it never has to be reset while the kernel runs.
template <bool kDyn>
__global__ void __launch_bounds__(kThreads, 1) k_cg_stream(const __grid_constant__ StreamArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    Tail *tl = reinterpret_cast<Tail *>(smem + kStages * kStageBytes);
    const int tid = threadIdx.x, wid = tid >> 5;
    const int G = gridDim.x;
    const int ntiles = a.tiles_x * a.tiles_y;
    const int n_my = (ntiles - (int)blockIdx.x + G - 1) / G;      // tiles blockIdx.x, +G, +2G, ...
    volatile unsigned int *err = &a.ctl->error;

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
        tma_prefetch_desc(&a.m_b_halo);
        tma_prefetch_desc(&a.m_r_halo);
        tma_prefetch_desc(&a.m_r_int);
        tma_prefetch_desc(&a.m_s0_halo);
        tma_prefetch_desc(&a.m_s1_halo);
        tma_prefetch_desc(&a.m_x_int);
        unsigned n = 0, ngo = 0;
        int cur = 0;   // index of the s buffer that RECEIVES s' in P1 (0: s0, 1: s1)
        // Loads of one tile are split in two groups: `indep` boxes do not depend on what other CTAs
        // write in the phase that is just ending and may be issued before its reduction completes;
        // `dep` boxes may only be issued after it (mbarrier `go`).
        auto tile_of = [&](int phase, int jj, int &tx0, int &ty0) {
            const int t = slot_tile(a, phase, jj, n_my, G, ntiles);
            tx0 = (t % a.tiles_x) * TW;
            ty0 = (t / a.tiles_x) * TH;
        };
        auto issue = [&](int it_, int phase, int cur_, unsigned nn, int tx0, int ty0, bool indep, bool dep) {
            const bool first = it_ == 0;
            const int st = nn % kStages;
            unsigned char *base = smem + st * kStageBytes;
            uint64_t *bar = &tl->full[st];
            if (first) {                                   // b never changes: everything is independent
                if (indep) {
                    mbar_arrive_expect_tx(bar, kHaloBoxBytes);
                    tma_load_2d(base, &a.m_b_halo, bar, tx0 - kHX, a.row0 + ty0 - 1);
                }
            } else if (phase == 0) {
                if (indep) {
                    mbar_arrive_expect_tx(bar, 2 * kHaloBoxBytes);
                    // s (old) is not touched by P2 -- except in iteration 0, whose P2 materialises s0 = b;
                    // iteration 1 therefore reads b itself (s = b, pcg.rs:40-42), which never changes
                    tma_load_2d(base, it_ == 1 ? &a.m_b_halo : (cur_ ? &a.m_s0_halo : &a.m_s1_halo), bar, tx0 - kHX, a.row0 + ty0 - 1);
                }
                if (dep) tma_load_2d(base + kHaloSlot, &a.m_r_halo, bar, tx0 - kHX, a.row0 + ty0 - 1);        // r: ring written by neighbours in P2
            } else {
                if (indep) {
                    mbar_arrive_expect_tx(bar, kHaloBoxBytes + 2 * kIntBoxBytes);
                    tma_load_2d(base + kHaloSlot, &a.m_r_int, bar, tx0, a.row0 + ty0);                        // own tiles, untouched by P1
                    tma_load_2d(base + 2 * kHaloSlot, &a.m_x_int, bar, tx0, a.row0 + ty0);
                }
                if (dep) tma_load_2d(base, cur_ ? &a.m_s1_halo : &a.m_s0_halo, bar, tx0 - kHX, a.row0 + ty0 - 1);  // s': ring written by neighbours in P1
            }
        };
        bool stop = false;
        if constexpr (kDyn) {
            const unsigned long long M = (unsigned long long)a.nbatch + (unsigned long long)G;   // claims per phase instance
            for (int it = 0; it < a.max_iter && !stop; ++it) {
                for (int phase = 0; phase < 2 && !stop; ++phase) {
                    const int k = 2 * it + phase;
                    bool exhausted = false;
                    int b_idx = -1, b_next = 0, b_end = 0;    // current batch: index, next claim-order position, end position
                    // next tile of this phase (and whether it closes its batch), or false when the list is exhausted
                    auto next_tile = [&](int &t, int &closes) -> bool {
                        if (b_next == b_end) {
                            if (exhausted) return false;
                            const unsigned long long v = atomicAdd(&a.claim[k & 3], 1ULL) - (unsigned long long)(k >> 2) * M;
                            if (v >= (unsigned long long)a.nbatch) { exhausted = true; return false; }   // exactly once per phase
                            b_idx = (int)v;
                            if (b_idx < a.nbatch_long) { b_next = b_idx * a.batch_len; b_end = b_next + a.batch_len; }
                            else { b_next = a.nbatch_long * a.batch_len + (b_idx - a.nbatch_long); b_end = b_next + 1; }
                        }
                        const int pos = b_next++;
                        t = (phase == 1 && a.zigzag) ? ntiles - 1 - pos : pos;
                        if (a.order) t = __ldg(a.order + t);
                        closes = b_next == b_end ? b_idx : -1;
                        return true;
                    };
                    const bool need_go = k != 0;
                    int pre[kStages], npre = 0;
                    // (1) claim the first tiles and stage their independent boxes while the previous phase is still finishing
                    if (need_go) {
                        while (npre < kStages) {
                            int t, closes;
                            if (!next_tile(t, closes)) break;
                            const unsigned nn = n + npre;
                            if (!mbar_wait(&tl->empty[nn % kStages], ((nn / kStages) & 1) ^ 1, err)) return;
                            tl->tile[nn % kStages] = t;
                            tl->batch[nn % kStages] = closes;
                            issue(it, phase, cur, nn, (t % a.tiles_x) * TW, (t / a.tiles_x) * TH, true, false);
                            pre[npre++] = t;
                        }
                        // (2) the previous phase's grid-wide reduction
                        if (!mbar_wait(&tl->go, ngo & 1, err)) return;
                        ++ngo;
                        stop = !*(volatile int *)&tl->cont;
                        fence_proxy_async();
                    }
                    // (3) their dependent boxes (also when stopping: every armed barrier must complete)
                    for (int jj = 0; jj < npre; ++jj)
                        issue(it, phase, cur, n + jj, (pre[jj] % a.tiles_x) * TW, (pre[jj] / a.tiles_x) * TH, false, true);
                    if (stop) {
                        for (int jj = 0; jj < npre; ++jj)
                            if (!mbar_wait(&tl->full[(n + jj) % kStages], ((n + jj) / kStages) & 1, err)) return;
                        return;
                    }
                    n += npre;
                    // (4) the rest of the phase
                    for (;;) {
                        int t, closes;
                        if (!next_tile(t, closes)) break;
                        if (!mbar_wait(&tl->empty[n % kStages], ((n / kStages) & 1) ^ 1, err)) return;
                        tl->tile[n % kStages] = t;
                        tl->batch[n % kStages] = closes;
                        issue(it, phase, cur, n, (t % a.tiles_x) * TW, (t / a.tiles_x) * TH, true, true);
                        ++n;
                    }
                    // (5) end-of-phase marker for the consumers: an empty slot
                    if (!mbar_wait(&tl->empty[n % kStages], ((n / kStages) & 1) ^ 1, err)) return;
                    tl->tile[n % kStages] = -1;
                    mbar_arrive(&tl->full[n % kStages]);
                    ++n;
                }
                cur ^= 1;
            }
            if (!stop) mbar_wait(&tl->go, ngo & 1, err);
            return;
        }
        for (int it = 0; it < a.max_iter && !stop; ++it) {
            for (int phase = 0; phase < 2 && !stop; ++phase) {
                const bool need_go = !(it == 0 && phase == 0);
                const int npre = need_go ? (n_my < kStages ? n_my : kStages) : 0;
                int tx0, ty0;
                // (1) independent boxes of the first tiles, while the previous phase is still finishing
                for (int jj = 0; jj < npre; ++jj) {
                    tile_of(phase, jj, tx0, ty0);
                    if (!mbar_wait(&tl->empty[(n + jj) % kStages], (((n + jj) / kStages) & 1) ^ 1, err)) return;
                    issue(it, phase, cur, n + jj, tx0, ty0, true, false);
                }
                // (2) the previous phase's grid-wide reduction
                if (need_go) {
                    if (!mbar_wait(&tl->go, ngo & 1, err)) return;
                    ++ngo;
                    stop = !*(volatile int *)&tl->cont;
                    fence_proxy_async();
                }
                // (3) dependent boxes of those tiles (issued even when stopping, so that every armed
                //     barrier completes and no bulk copy is left in flight when the CTA exits)
                for (int jj = 0; jj < npre; ++jj) {
                    tile_of(phase, jj, tx0, ty0);
                    issue(it, phase, cur, n + jj, tx0, ty0, false, true);
                }
                if (stop) {
                    for (int jj = 0; jj < npre; ++jj)
                        if (!mbar_wait(&tl->full[(n + jj) % kStages], ((n + jj) / kStages) & 1, err)) return;
                    return;
                }
                // (4) the rest of the phase
                for (int jj = npre; jj < n_my; ++jj) {
                    tile_of(phase, jj, tx0, ty0);
                    if (!mbar_wait(&tl->empty[(n + jj) % kStages], (((n + jj) / kStages) & 1) ^ 1, err)) return;
                    issue(it, phase, cur, n + jj, tx0, ty0, true, true);
                }
                n += n_my;
            }
            cur ^= 1;
        }
        // the reduction that ends the last phase (nothing follows it)
        if (!stop) mbar_wait(&tl->go, ngo & 1, err);
        return;
    }

    // ================================================================ consumer warps
    unsigned n = 0;
    unsigned long long nred = 0;
    double sigma = 0, alpha = 0, beta = 0, rmax = 0, bmax = 0;
    int it = 0, applies = 0;
    bool converged = false, early = false;
    double *s_cur = a.s0, *s_old = a.s1;
    double red[3];

    const bool prof = a.dbg != nullptr && tid == 0;
    long long tprev = prof ? clock64() : 0;
    auto stamp = [&](int slot) {
        if (prof) {
            const long long t = clock64();
            if (blockIdx.x == 0) a.dbg[slot] += t - tprev;
            if ((slot & 1) == 0) a.dbg_cta[(slot >> 1) * gridDim.x + blockIdx.x] += t - tprev;   // tile loops of every CTA
            tprev = t;
        }
    };
    // leading halo tiles of this CTA in every phase (slot_tile), when they are flagged early
    int nh_early = 0;
    if (!kDyn && a.halo_first && a.tiles_y > 2 && a.xr.nranks > 1 && a.xr.hflags != nullptr) {
        const int nhalo = 2 * a.tiles_x;
        nh_early = nhalo > (int)blockIdx.x ? (nhalo - (int)blockIdx.x + G - 1) / G : 0;
        if (nh_early > n_my) nh_early = n_my;
    }
    for (it = 0; it < a.max_iter; ++it) {
        const bool first = it == 0;
        // ------------------------------------------------------------------ P1
        double acc_zs = 0, acc_bb = 0, acc_bmax = 0;
        bool remote = false;      // uniform over the CTA: every consumer warp walks the same tiles
        const unsigned long long tag1 = a.seq_base + (unsigned long long)(2 * it) + 1;   // phase tag of the per-tile units
        for (int jj = 0;; ++jj, ++n) {
            const int st = n % kStages;
            int t;
            if constexpr (kDyn) {
                if (!mbar_wait(&tl->full[st], (n / kStages) & 1, err)) return;
                t = slot_word(&tl->tile[st]);
                if (t < 0) {                                    // end-of-phase marker: hand the slot back and leave
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(&tl->empty[st]);
                    ++n;
                    break;
                }
            } else {
                if (jj >= n_my) break;
                t = slot_tile(a, 0, jj, n_my, G, ntiles);
                if (!mbar_wait(&tl->full[st], (n / kStages) & 1, err)) return;
            }
            const int tx0 = (t % a.tiles_x) * TW, ty0 = (t / a.tiles_x) * TH;
            remote = remote || tile_stores_remote(a, ty0);
            const double *S = reinterpret_cast<const double *>(smem + st * kStageBytes);
            const double *R = reinterpret_cast<const double *>(smem + st * kStageBytes + kHaloSlot);
            const bool fast = tile_is_fast(a, ty0, tx0);
            if (first) {
                if (fast) tile_p1<true, true>(a, S, R, nullptr, nullptr, nullptr, ty0, tx0, 0.0, acc_zs, acc_bb, acc_bmax);
                else tile_p1<false, true>(a, S, R, nullptr, nullptr, nullptr, ty0, tx0, 0.0, acc_zs, acc_bb, acc_bmax);
            } else {
                double *s_up = s_cur == a.s0 ? a.up_s0 : a.up_s1, *s_dn = s_cur == a.s0 ? a.dn_s0 : a.dn_s1;
                if (fast) tile_p1<true, false>(a, S, R, s_cur, s_up, s_dn, ty0, tx0, beta, acc_zs, acc_bb, acc_bmax);
                else tile_p1<false, false>(a, S, R, s_cur, s_up, s_dn, ty0, tx0, beta, acc_zs, acc_bb, acc_bmax);
            }
            const int closes = kDyn ? slot_word(&tl->batch[st]) : -1;
            if (kDyn && closes >= 0) {                          // this warp's partials of the batch that ends here, then afresh
                const size_t plane = (size_t)a.nbatch * kConsumerWarps, u = (size_t)closes * kConsumerWarps + wid;
                const double p0 = warp_sum(acc_zs);
                if ((tid & 31) == 0) unit_store(a.tparts + u, p0, tag1);
                if (first) {
                    const double p1 = warp_sum(acc_bb), p2 = warp_max(acc_bmax);
                    if ((tid & 31) == 0) {
                        unit_store(a.tparts + plane + u, p1, tag1);
                        unit_store(a.tparts + 2 * plane + u, p2, tag1);
                    }
                }
                acc_zs = 0; acc_bb = 0; acc_bmax = 0;
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&tl->empty[st]);
            if (!kDyn && jj == nh_early - 1) early_halo_flags(a, nred);
        }
        stamp(0);   // P1 tiles
        {
            if constexpr (kDyn) {   // the units of this CTA's FIXED batch range, whoever computed them, in a fixed order
                const size_t plane = (size_t)a.nbatch * kConsumerWarps;
                const int c0 = (int)((long long)blockIdx.x * a.nbatch / G), c1 = (int)((long long)(blockIdx.x + 1) * a.nbatch / G);
                const ReduceUnit *base = a.tparts + (size_t)c0 * kConsumerWarps;
                for (int e = tid; e < (c1 - c0) * kConsumerWarps; e += kConsumers) {
                    double v;
                    unit_poll(base + e, tag1, v, err);
                    acc_zs = acc_zs + v;
                    if (first) {
                        unit_poll(base + plane + e, tag1, v, err);
                        acc_bb = acc_bb + v;
                        unit_poll(base + 2 * plane + e, tag1, v, err);
                        acc_bmax = v > acc_bmax ? v : acc_bmax;
                    }
                }
            }
            const double v0 = consumer_sum(acc_zs, tl->wsum[0]);
            double v1 = 0, v2 = 0;
            if (first) {
                v1 = consumer_sum(acc_bb, tl->wsum[1]);
                v2 = consumer_max(acc_bmax, tl->wsum[2]);
            }
            if (!grid_allreduce(a, tl, nred++, first ? 3 : 1, v0, v1, v2, 0x4u, red, remote, nh_early > 0)) {
                if (tid == 0) { tl->cont = 0; mbar_arrive(&tl->go); }
                return;
            }
        }
        stamp(1);   // CTA reduction + grid all-reduce #1
        const double zs = red[0];
        if (first) {
            sigma = red[1];                                    // pcg.rs:46
            bmax = red[2];                                     // pcg.rs:35
            rmax = bmax;
            if (bmax < a.threshold) early = true;              // pcg.rs:35-38
        }
        if (early) {
            if (tid == 0) { tl->cont = 0; mbar_arrive(&tl->go); }
            break;
        }
        if (tid == 0) mbar_arrive(&tl->go);
        ++applies;
        alpha = sigma / zs;                                    // pcg.rs:53
        // ------------------------------------------------------------------ P2
        double acc_rr = 0, acc_rmax = 0;
        remote = false;
        const unsigned long long tag2 = a.seq_base + (unsigned long long)(2 * it + 1) + 1;
        for (int jj = 0;; ++jj, ++n) {
            const int st = n % kStages;
            int t;
            if constexpr (kDyn) {
                if (!mbar_wait(&tl->full[st], (n / kStages) & 1, err)) return;
                t = slot_word(&tl->tile[st]);
                if (t < 0) {
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(&tl->empty[st]);
                    ++n;
                    break;
                }
            } else {
                if (jj >= n_my) break;
                t = slot_tile(a, 1, jj, n_my, G, ntiles);
                if (!mbar_wait(&tl->full[st], (n / kStages) & 1, err)) return;
            }
            const int tx0 = (t % a.tiles_x) * TW, ty0 = (t / a.tiles_x) * TH;
            remote = remote || tile_stores_remote(a, ty0);
            const double *S = reinterpret_cast<const double *>(smem + st * kStageBytes);
            const double *R = reinterpret_cast<const double *>(smem + st * kStageBytes + kHaloSlot);
            const double *X = reinterpret_cast<const double *>(smem + st * kStageBytes + 2 * kHaloSlot);
            const bool fast = tile_is_fast(a, ty0, tx0);
            if (first) {
                if (fast) tile_p2<true, true>(a, S, R, X, ty0, tx0, alpha, acc_rr, acc_rmax);
                else tile_p2<false, true>(a, S, R, X, ty0, tx0, alpha, acc_rr, acc_rmax);
            } else {
                if (fast) tile_p2<true, false>(a, S, R, X, ty0, tx0, alpha, acc_rr, acc_rmax);
                else tile_p2<false, false>(a, S, R, X, ty0, tx0, alpha, acc_rr, acc_rmax);
            }
            const int closes = kDyn ? slot_word(&tl->batch[st]) : -1;
            if (kDyn && closes >= 0) {
                const size_t plane = (size_t)a.nbatch * kConsumerWarps, u = (size_t)closes * kConsumerWarps + wid;
                const double p0 = warp_sum(acc_rr), p1 = warp_max(acc_rmax);
                if ((tid & 31) == 0) {
                    unit_store(a.tparts + u, p0, tag2);
                    unit_store(a.tparts + plane + u, p1, tag2);
                }
                acc_rr = 0; acc_rmax = 0;
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&tl->empty[st]);
            if (!kDyn && jj == nh_early - 1) early_halo_flags(a, nred);
        }
        stamp(2);   // P2 tiles
        {
            if constexpr (kDyn) {
                const size_t plane = (size_t)a.nbatch * kConsumerWarps;
                const int c0 = (int)((long long)blockIdx.x * a.nbatch / G), c1 = (int)((long long)(blockIdx.x + 1) * a.nbatch / G);
                const ReduceUnit *base = a.tparts + (size_t)c0 * kConsumerWarps;
                for (int e = tid; e < (c1 - c0) * kConsumerWarps; e += kConsumers) {
                    double v;
                    unit_poll(base + e, tag2, v, err);
                    acc_rr = acc_rr + v;
                    unit_poll(base + plane + e, tag2, v, err);
                    acc_rmax = v > acc_rmax ? v : acc_rmax;
                }
            }
            const double v0 = consumer_sum(acc_rr, tl->wsum[0]);
            const double v1 = consumer_max(acc_rmax, tl->wsum[1]);
            if (!grid_allreduce(a, tl, nred++, 2, v0, v1, 0.0, 0x2u, red, remote, nh_early > 0)) {
                if (tid == 0) { tl->cont = 0; mbar_arrive(&tl->go); }
                return;
            }
        }
        stamp(3);   // CTA reduction + grid all-reduce #2
        const double rr = red[0];
        rmax = red[1];                                         // pcg.rs:58
        if (rmax < a.threshold) {                              // pcg.rs:60-63
            converged = true;
            if (tid == 0) { tl->cont = 0; mbar_arrive(&tl->go); }
            break;
        }
        if (tid == 0) mbar_arrive(&tl->go);
        beta = rr / sigma;                                     // pcg.rs:67-68
        sigma = rr;                                            // pcg.rs:79
        double *tmp = s_cur;
        s_cur = s_old;
        s_old = tmp;
    }

    // ------------------------------------------------------------------ epilogue (flat, once per solve)
    const size_t ncell = (size_t)a.h * a.w, base = (size_t)a.row0 * a.w;
    const size_t stride = (size_t)G * kConsumers, i0 = base + (size_t)blockIdx.x * kConsumers + tid;
    if (early) {
        for (size_t i = i0; i < base + ncell; i += stride) a.x[i] = 0.0;
    } else {
        // converged: the last applied direction is in s_cur; exhausted: the swap already happened, it is
        // in s_old, and the reference still performs the search update (pcg.rs:72-77) before leaving
        const double *s_fin = converged ? s_cur : s_old;
        if (!converged || s_fin != a.s0) {
            for (size_t i = i0; i < base + ncell; i += stride) {
                const double sv = __ldcg(s_fin + i);
                a.s0[i] = converged ? sv : __ldcg(a.r + i) + beta * sv;
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

// ---------------------------------------------------------------------------------- host side
int pano_make_tensor_map_2d(CUtensorMap *map, const void *base, size_t elem_bytes, uint64_t width, uint64_t height,
                            uint64_t row_pitch_bytes, uint32_t box_w, uint32_t box_h) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        PANO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) PANO_FAIL(PANO_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = (EncodeFn)fn;
    }
    if (elem_bytes != 8) PANO_FAIL(PANO_ERR_INVALID, "tensor maps are built for f64 fields only");
    const cuuint64_t dims[2] = {width, height};
    const cuuint64_t strides[1] = {row_pitch_bytes};
    const cuuint32_t box[2] = {box_w, box_h};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) PANO_FAIL(PANO_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (w=%llu h=%llu pitch=%llu box=%ux%u)",
                                      (int)rc, (unsigned long long)width, (unsigned long long)height,
                                      (unsigned long long)row_pitch_bytes, box_w, box_h);
    return PANO_OK;
}

int pano_cg_tile_order(pano_ctx *ctx, int h, int w, int gy0, int gh, RectI m, int tiles_x, int tiles_y, int th, int tw, int margin,
                       bool has_up, bool has_dn, const int **order_out, int *lo_out, int *hi_out);

int pano_preload_cg_stream() {
    cudaFuncAttributes fa;
    PANO_CUDA(cudaFuncGetAttributes(&fa, k_cg_stream<false>));
    PANO_CUDA(cudaFuncSetAttribute(k_cg_stream<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    PANO_CUDA(cudaFuncGetAttributes(&fa, k_cg_stream<true>));
    PANO_CUDA(cudaFuncSetAttribute(k_cg_stream<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    return PANO_OK;
}

// TMA needs 16-byte aligned rows: even width, 16-byte aligned base pointers.
bool pano_cg_stream_supported(size_t h, size_t w, const void *x, const void *b, const void *r, const void *s0, const void *s1) {
    if (w % 2 != 0 || w < 2 || h < 1) return false;
    const void *ps[] = {x, b, r, s0, s1};
    for (const void *p : ps)
        if (((uintptr_t)p & 15u) != 0) return false;
    return true;
}

int pano_cg_stream_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, double *s1, size_t h, size_t w,
                          int max_iterations, double threshold, double timestep, RectI m, const PanoCgSlab *slab) {
    // the producer of a launch with no iterations would wait on `go` for an arrival that never comes
    if (max_iterations <= 0) PANO_FAIL(PANO_ERR_INVALID, "pano_cg_stream_launch: max_iterations = %d (callers handle the empty loop on the host)", max_iterations);
    PANO_CUDA(cudaFuncSetAttribute(k_cg_stream<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    PANO_CUDA(cudaFuncSetAttribute(k_cg_stream<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    static_assert(sizeof(Tail) <= kTailBytes, "Tail does not fit");
    StreamArgs a;
    memset(&a, 0, sizeof(a));
    const uint64_t pitch = (uint64_t)w * 8;
    const uint64_t rows = slab ? (uint64_t)slab->rows_total : (uint64_t)h;   // h = owned rows, rows = stored rows
    PANO_TRY(pano_make_tensor_map_2d(&a.m_b_halo, b, 8, w, rows, pitch, BW, BH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_r_halo, r, 8, w, rows, pitch, BW, BH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_r_int, r, 8, w, rows, pitch, TW, TH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_s0_halo, s0, 8, w, rows, pitch, BW, BH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_s1_halo, s1, 8, w, rows, pitch, BW, BH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_x_int, x, 8, w, rows, pitch, TW, TH));
    a.x = x; a.b = b; a.r = r; a.s0 = s0; a.s1 = s1;
    a.h = (int)h; a.w = (int)w;
    a.dt = timestep; a.threshold = threshold; a.max_iter = max_iterations;
    a.m = m;
    a.tiles_x = ((int)w + TW - 1) / TW;
    a.tiles_y = ((int)h + TH - 1) / TH;
    a.ctl = ctx->d_cg;
    a.dbg = pano_option(ctx, "cg_profile", 0) ? ctx->d_cg->prof : nullptr;
    a.dbg_cta = reinterpret_cast<long long *>(ctx->d_partials);   // >= 12288 doubles (pano_ctx_create); zeroed below when profiling
    if (a.dbg) PANO_CUDA(cudaMemsetAsync(ctx->d_partials, 0, 2 * kMaxCtas * sizeof(long long), ctx->stream));
    a.zigzag = pano_option(ctx, "cg_zigzag", 1) != 0;
    a.fence_mode = (int)pano_option(ctx, "cg_fence", 0);
    a.order = nullptr;
    a.row0 = 0; a.gy0 = 0; a.gh = (int)h;
    a.xr.rank = 0; a.xr.nranks = 1;
    int max_ctas = 0;
    if (slab) {
        a.row0 = slab->row0; a.gy0 = slab->gy0; a.gh = slab->gh;
        a.up_r = slab->up_r; a.up_s0 = slab->up_s0; a.up_s1 = slab->up_s1;
        a.dn_r = slab->dn_r; a.dn_s0 = slab->dn_s0; a.dn_s1 = slab->dn_s1;
        a.xr.rank = slab->rank; a.xr.nranks = slab->nranks;
        a.xr.seq_base = slab->xseq_base;
        a.xr.local = (ReduceUnit *)slab->xunits_local;
        for (int i = 0; i < kMaxRanks; ++i) a.xr.peer[i] = (ReduceUnit *)slab->xunits_peer[i];
        if (slab->nranks > 1 && pano_option(ctx, "cg_xflags", 1) != 0) {
            // halo flags (pano_sm100.cuh): behind the cross-rank totals in every rank's unit array.  The neighbours run
            // the same launch code on their slabs, so their CTA counts follow from the slab split.
            auto ctas_of = [&](int r) {
                const long long rows = (long long)slab->gh * (r + 1) / slab->nranks - (long long)slab->gh * r / slab->nranks;
                long long g = ctx->num_sms, nt = (long long)a.tiles_x * ((rows + TH - 1) / TH);
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
        a.halo_first = (slab->nranks > 1 && pano_option(ctx, "cg_halo_first", 0) != 0) ? 1 : 0;
    }
    // "cg_order_mid": -1 auto, 0 off, 1 on.  Measured (two-reduction kernel, one GPU): 4096^2 (55 tiles per CTA) 17.70 -> 17.49 ms per
    // solve, 8192^2 (221) 67.01 -> 67.62: the shorter a pass, the more the slow tiles at its end cost; auto = up to 96 tiles per CTA.
    const int64_t om = pano_option(ctx, "cg_order_mid", -1);
    if ((om > 0 || (om < 0 && (long long)a.tiles_x * a.tiles_y <= 96LL * ctx->num_sms)) && !a.halo_first) {
        // select-path tiles (walls, obstacle) away from both ends of the tile order: see pano_cg_sr.cu, tile_at
        const bool multi = slab && slab->nranks > 1;
        int lo, hi;
        PANO_TRY(pano_cg_tile_order(ctx, a.h, a.w, a.gy0, a.gh, a.m, a.tiles_x, a.tiles_y, TH, TW, 1, multi && slab->rank > 0,
                                    multi && slab->rank + 1 < slab->nranks, &a.order, &lo, &hi));
    }
    static_assert(kUnitsTotal * sizeof(ReduceUnit) <= 4096 * 16, "d_units (allocated in pano_ctx_create) is too small");
    a.units = (ReduceUnit *)ctx->d_units;
    a.seq_base = (++ctx->launch_epoch) << 32;
    int G = ctx->num_sms;
    if (max_ctas > 0 && G > max_ctas) G = max_ctas;
    const int ntiles = a.tiles_x * a.tiles_y;
    if (G > ntiles) G = ntiles;
    if (G > kMaxCtas) G = kMaxCtas;
    PANO_TRY(pano_cg_control_reset(ctx));
    // dynamic tile scheduling pays once a CTA has enough tiles for the imbalance to matter ("cg_dynamic": 0 off, 1 on, -1 auto)
    const int64_t dyn_opt = pano_option(ctx, "cg_dynamic", -1);
    // measured on B200 (Mcell-steps/s, static -> dynamic): 2048^2 (14 tiles per CTA) 787 -> 758; 8192^2 over 8 GPUs (28) 5488 ->
    // 5549, i.e. noise; 4096^2 (55) 893 -> 944; 8192^2 (221) 842 -> 996.  The second reduction stage costs ~1 us per phase.
    // 1024 x 8192 slab (28): one GPU 96.2 -> 94.9 us per iteration, 2 GPUs 105.1 -> 104.0: the auto threshold is 24 tiles per CTA.
    const bool dynamic = dyn_opt > 0 || (dyn_opt < 0 && ntiles >= 24 * G);
    if (dynamic) {
        // the fixed batch list: ~80 % of the tiles in batches of up to 8 (about six long batches per CTA), then single tiles
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
    if (dynamic)
        PANO_CUDA(cudaLaunchCooperativeKernel((const void *)k_cg_stream<true>, dim3((unsigned)G), dim3(kThreads), kargs, kSmemBytes, ctx->stream));
    else
        PANO_CUDA(cudaLaunchCooperativeKernel((const void *)k_cg_stream<false>, dim3((unsigned)G), dim3(kThreads), kargs, kSmemBytes, ctx->stream));
    return pano_after_launch(ctx, dynamic ? "cg_stream(dynamic)" : "cg_stream");
}
