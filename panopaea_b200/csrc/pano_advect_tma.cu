// pano_advect_tma.cu -- K1, fourth generation: the fused advection (advect + both loops of advect_mac,
// examples/dec_fluid.rs:59-60, 173-291) as a persistent, warp-specialised, TMA-staged kernel (one CTA per SM).
//
// k_advect_march3 (pano_fused.cu) gathers through L1/L2 and is bound by the latency of those gathers (ncu, 4096^2:
// DRAM 47 %, long-scoreboard stalls 7.7 per issue, no waste: 0.95 x the algorithmic bytes).  Here the loads are decoupled
// from the arithmetic:
//   * a producer lane claims 32 x 64-cell tiles from a counter and streams, per tile, the (32+4) x (64+4) boxes of q and vy
//     and the box of vx into a 3-stage shared-memory ring with cp.async.bulk.tensor (TMA, mbarrier completion).  vx rows
//     are w+1 doubles long -- an odd multiple of 8 bytes, which no tensor map accepts as a stride -- so vx is described as
//     h/2 "super-rows" of 2(w+1) doubles and fetched as two boxes per tile: its even rows, and its odd rows one column to
//     the left (TMA box starts must be 16-byte aligned: measured, scripts/probe/tma_probe.cu)
//   * 16 consumer warps compute from shared memory: a thread marches down 4 rows of one column carrying the eight
//     velocity samples around its cell in registers; the twelve gathers of a row are shared-memory loads
//   * a backtrace that leaves the box (more than two cells: |v| dt >= 2) falls back to global memory for that cell; tiles
//     within one tile of the domain border, where the reference's index / coordinate clamps bite, run the marching body
//     with its clamps on the same staged boxes (first version: on global memory -- 4.6 % of the tiles took 17 % of the
//     kernel's stall samples, all long-scoreboard); the x = w / y = h strips run it on global memory (pano_advect_body.cuh)
// Arithmetic is pano_cell_math.h's in every path: bit-identical to the reference (tests/test_gpu_fused.py).
// HBM traffic: 48 B per cell (halo re-reads hit the 126 MB L2: neighbouring tiles are in flight together).
#include "pano_advect_body.cuh"
#include "pano_sm100.cuh"

using namespace pano_sm100;
using pano_adv::V32;
using pano_adv::V32W;

namespace {

#ifndef PANO_ADV_WARPS
#define PANO_ADV_WARPS 16                       // consumer warps (measured at 4096^2: 16 -> 157 us, 18 -> 168 us with 36-row tiles; 20 leave 80 registers: spills)
#endif
#ifndef PANO_ADV_TW
#define PANO_ADV_TW 64                          // tile width (cells): 64 -> 32-row tiles; 128 -> 16-row tiles with rows twice as long measured 166 vs 157 us at 4096^2
#endif
constexpr int kConsumerWarps = PANO_ADV_WARPS, kConsumers = 32 * kConsumerWarps;
constexpr int TW = PANO_ADV_TW, kColGroups = TW / 32;          // a warp covers 32 columns x 4 rows
constexpr int TH = 4 * (kConsumerWarps / kColGroups);          // tile (cells)
constexpr int kG = 2;                           // halo: every gather of a cell whose backtrace is shorter than 2 cells stays inside
constexpr int QW = TW + 2 * kG, QH = TH + 2 * kG;   // 68 x 36: q and vy boxes
constexpr int XW = QW + 2, XH = QH / 2;         // 70 x 18: each of the two vx boxes (even rows; odd rows shifted one column left)
constexpr int kQBytes = QH * QW * 8;            // 19584 (a multiple of 128)
constexpr int kXBytes = XH * XW * 8;            // 10080
constexpr int kXSlot = (kXBytes + 127) / 128 * 128;   // rounded up to a multiple of 128
constexpr int kXOdd = kXSlot / 8 + 1;           // offset (doubles) from an even-row element to the element one row down: odd box + its column shift
constexpr int kStageBytes = 2 * kQBytes + 2 * kXSlot;   // 59392
constexpr int kStages = 3;
constexpr int kThreads = kConsumers + 32;       // + one producer warp
constexpr int kRows = TH / (kConsumerWarps / kColGroups);   // 4 rows per thread
constexpr int kTailBytes = 1024;
constexpr int kSmemBytes = kStages * kStageBytes + kTailBytes;   // 179200
static_assert(kQBytes % 128 == 0 && kRows == 4 && TH % 2 == 0 && kSmemBytes <= 232448, "stage layout");

struct AdvArgs {
    CUtensorMap m_q, m_vy, m_vx;  // m_vx: the super-row view (2(w+1) wide, rows/2 high)
    double *q_dst, *vy_dst, *vx_dst;              // all pointers are the (virtual) addresses of GLOBAL row 0
    const double *q_src, *vy_src, *vx_src;
    int h, w;                     // the whole grid (walls, clamps)
    double dt;
    int ya, yb;                   // cell rows produced here (0, h on one GPU; a slab otherwise)
    int ylo;                      // global row held by row 0 of the stored source arrays (0 on one GPU)
    int whi_q, whi_vy;            // slab: stored rows are [ylo, whi_q) for q and vx, [ylo, whi_vy) for vy
    unsigned int *err;            // slab: raised (= 2) when a gather leaves the stored rows
    int tiles_x, tiles_y;
    unsigned int *claim;          // two tile counters, used alternately by successive launches
    int parity;
    int dynamic;
    int border_first;             // claim order: border tiles first (tile_at)
};

struct Tail {
    uint64_t full[kStages], empty[kStages];
    int tile[kStages];
};

__device__ __forceinline__ bool tile_interior(const AdvArgs &a, int ty0, int tx0) {
    // no clamp of the reference can bite within two cells of the tile, and the tile is complete
    return tx0 >= TW && tx0 + 2 * TW <= a.w && ty0 >= TH && ty0 + 2 * TH <= a.h && ty0 + TH <= a.yb;
}

// advect_mac's gather coordinates without the index clamps (the caller checks that the corner lies inside the staged box,
// which lies inside the grid): same bits as pano::mac_coord_fast there
// A negative coordinate makes floor_nonneg's guard word differ (v + 2^52 drops below 2^52), so it shows up in `bad` like a
// coordinate beyond 2^32 does, and the cell takes the full forms; for v >= 0 max(v, 0) = v and the weight v - floor(v) is the
// reference's (pano_cell_math.h).
struct MacCoordI {
    int x0, y0;
    double s, t;
    unsigned bad;
};
__device__ __forceinline__ MacCoordI mac_coord_open(double relx, double rely) {
    MacCoordI c;
    const pano::FloorNN fx = pano::floor_nonneg(relx), fy = pano::floor_nonneg(rely);
    c.bad = (fx.hi ^ pano::kFloorHi) | (fy.hi ^ pano::kFloorHi);
    c.x0 = (int)fx.i; c.y0 = (int)fy.i;
    c.s = relx - fx.f; c.t = rely - fy.f;
    return c;
}

// Border tiles: the marching body with the reference's clamps, reading the staged boxes where the (clamped, hence in-grid)
// index falls inside them and global memory otherwise.  Out-of-grid parts of a box are zero-filled by TMA and never addressed.
template <class A>
struct BoxQ {      // q or vy: box origin (by0, bx0), QH x QW; rows from `rh` on lie beyond the STORED rows of a slab (zero-filled
    const double *sm;   // like out-of-grid ones, but a long backtrace may address them: the global accessor then reports it)
    int by0, bx0;
    unsigned rh;
    A g;
    __device__ __forceinline__ double operator()(int y, int x) const {
        const unsigned ry = (unsigned)(y - by0), rx = (unsigned)(x - bx0);
        if (ry < rh && rx < (unsigned)QW) return sm[ry * QW + rx];
        return g(y, x);
    }
};
template <class A>
struct BoxX {      // vx: even rows at [(r >> 1) * XW + c], odd rows at [kXOdd + (r >> 1) * XW + c]
    const double *sm;
    int by0, bx0;
    unsigned rh;
    A g;
    __device__ __forceinline__ double operator()(int y, int x) const {
        const unsigned ry = (unsigned)(y - by0), rx = (unsigned)(x - bx0);
        if (ry < rh && rx < (unsigned)QW) return sm[(ry >> 1) * XW + rx + (ry & 1u) * kXOdd];
        return g(y, x);
    }
};

// one cell whose backtrace left the staged boxes: the full forms on global memory
// the global-memory accessors of a launch (slab: windowed, a gather beyond the stored rows raises the error word)
template <bool kSlab>
struct GlobalAcc {
    using Acc = typename std::conditional<kSlab, V32W, V32<double>>::type;
    Acc q, vy, vx;
    __device__ __forceinline__ explicit GlobalAcc(const AdvArgs &a) {
        if constexpr (kSlab) {
            q = V32W{a.q_src, a.w, a.ylo, a.whi_q, a.err};
            vy = V32W{a.vy_src, a.w, a.ylo, a.whi_vy, a.err};
            vx = V32W{a.vx_src, a.w + 1, a.ylo, a.whi_q, a.err};
        } else {
            q = V32<double>{a.q_src, a.w};
            vy = V32<double>{a.vy_src, a.w};
            vx = V32<double>{a.vx_src, a.w + 1};
        }
    }
};

template <bool kSlab>
__device__ __noinline__ void cell_from_global(const AdvArgs &a, int y, int x, double ucx, double ucy, double rxx, double rxy, double ryx,
                                              double ryy) {
    const GlobalAcc<kSlab> g(a);
    const int h = a.h, w = a.w;
    const pano::CellCoord cq = pano::advect_coord_fast((double)x + 0.5, (double)y + 0.5, (double)w - 1.00001, (double)h - 1.00001, -a.dt, ucx, ucy);
    a.q_dst[y * w + x] = pano::advect_gather_at(cq, g.q);
    const pano::MacCoord cx = pano::mac_coord_fast(rxx, rxy, h, w + 1), cy = pano::mac_coord_fast(ryx, ryy, h + 1, w);
    a.vx_dst[y * (w + 1) + x] = cx.bad ? pano_adv::mac_gather_far(rxx, rxy, h, w + 1, g.vx) : pano::mac_gather_at(cx, g.vx);
    a.vy_dst[y * w + x] = cy.bad ? pano_adv::mac_gather_far(ryx, ryy, h + 1, w, g.vy) : pano::mac_gather_at(cy, g.vy);
}

// Interior tile: column x = tx0 + lx, rows ty0 + ly0 .. + kRows - 1 (ly0 a multiple of 4), everything from shared memory.
//   Q, VY: boxes with origin (ty0 - 2, tx0 - 2), pitch QW.   VX: even rows of the same box at VX[(r >> 1) * XW + c],
//   odd rows at VX[kXOdd + (r >> 1) * XW + c]  (r, c box-relative).
// kBorder: the same path for the threads of a BORDER tile whose four cells are at least one cell away from every wall (no
// border select in the velocity averages); a cell then also checks that its gathers stay two cells inside the grid, where no
// clamp of the reference bites, and within the box rows that hold stored data (`rows_ok`, slabs); otherwise it takes the
// full forms on global memory like a cell whose backtrace leaves the box.
template <bool kBorder, bool kSlab>
__device__ __forceinline__ void advect_tile(const AdvArgs &a, const double *__restrict__ Q, const double *__restrict__ VY,
                                            const double *__restrict__ VX, int ty0, int tx0, int ly0, int lx, unsigned rows_ok) {
    const int h = a.h, w = a.w;
    const int x = tx0 + lx, ys = ty0 + ly0;
    const int bx0 = tx0 - kG, by0 = ty0 - kG;
    const int c0 = lx + kG, r0 = ly0 + kG;                     // box column / first box row (even) of this thread
    const double ndt = -a.dt, xd = (double)x, xh = xd + 0.5;
    double yd = (double)ys;
    const double *pvy = VY + r0 * QW + c0;
    const double *pvx = VX + (r0 >> 1) * XW + c0;              // vx(ys, x): even box row
    double C = pvy[0], E = pvy[-1];
    double G = pvx[kXOdd - XW], H = pvx[kXOdd - XW + 1];       // vx(ys - 1, .): the odd row above
    double *qo = a.q_dst + ys * w + x, *vyo = a.vy_dst + ys * w + x, *vxo = a.vx_dst + ys * (w + 1) + x;
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const int y = ys + k;
        const double yh = yd + 0.5;
        // vx(y, x), vx(y, x+1): rows alternate between the even and the odd box; vy(y+1, x), vy(y+1, x-1)
        const double *px = pvx + (k >> 1) * XW + ((k & 1) ? kXOdd : 0);
        const double A_ = px[0], B = px[1];
        const double D = pvy[(k + 1) * QW], F = pvy[(k + 1) * QW - 1];
        const double vvy = (C + D + E + F) / 4.0;               // dec_fluid.rs:220-225 with xc = x, xm = x - 1
        const double vvx = (A_ + B + G + H) / 4.0;              // :257-263 with yc = y, ym = y - 1
        // advect's coordinates (dec_fluid.rs:184-192): inside the box, which lies >= 62 cells inside the grid, the clamps
        // max(., 0) and min(., w - 1.00001) are identities, so they are skipped here and the box test below stands in for them
        const double ucx = (A_ + B) / 2.0, ucy = (C + D) / 2.0;
        const double pqx = (xh + ndt * ucx) - 0.5, pqy = (yh + ndt * ucy) - 0.5;
        const MacCoordI cq = mac_coord_open(pqx, pqy);
        double rxx, rxy, ryx, ryy;
        pano::mac_x_rel(xd, yh, ndt, A_, vvy, rxx, rxy);
        pano::mac_y_rel(xh, yd, ndt, vvx, C, ryx, ryy);
        const MacCoordI cx = mac_coord_open(rxx, rxy), cy = mac_coord_open(ryx, ryy);
        // box-relative corners; a corner at (u, v) needs u + 1 and v + 1 as well
        const unsigned qx = (unsigned)(cq.x0 - bx0), qy = (unsigned)(cq.y0 - by0);
        const unsigned xx = (unsigned)(cx.x0 - bx0), xy = (unsigned)(cx.y0 - by0);
        const unsigned yx = (unsigned)(cy.x0 - bx0), yy = (unsigned)(cy.y0 - by0);
        const unsigned mx = max(max(qx, xx), yx), my = max(max(qy, xy), yy);
        bool inside = (cq.bad | cx.bad | cy.bad) == 0u && mx <= (unsigned)(QW - 2) && my <= (kBorder ? rows_ok : (unsigned)(QH - 2));
        if (kBorder)   // corners at most (h-3, w-3): x0 + 1 <= w - 2 and y0 + 1 <= h - 2 in every array, and advect's min(., w - 1.00001) is idle
            inside = inside && max(max(cq.x0, cx.x0), cy.x0) <= w - 3 && max(max(cq.y0, cx.y0), cy.y0) <= h - 3;
        if (inside) {
            // one straight-line block: all twelve gathers in flight together
            const double *gq_ = Q + qy * QW + qx;
            const double *gy_ = VY + yy * QW + yx;
            const unsigned par = xy & 1u;
            const double *gx0 = VX + (xy >> 1) * XW + xx + par * kXOdd;                  // row y0
            const double *gx1 = VX + (xy >> 1) * XW + xx + kXOdd - par * (kXOdd - XW);   // row y0 + 1
            const double q00 = gq_[0], q01 = gq_[1], q10 = gq_[QW], q11 = gq_[QW + 1];
            const double x00 = gx0[0], x01 = gx0[1], x10 = gx1[0], x11 = gx1[1];
            const double y00 = gy_[0], y01 = gy_[1], y10 = gy_[QW], y11 = gy_[QW + 1];
            qo[k * w] = pano::bilinear(q00, q01, q10, q11, cq.s, cq.t);
            vxo[k * (w + 1)] = pano::bilinear(x00, x01, x10, x11, cx.s, cx.t);
            vyo[k * w] = pano::bilinear(y00, y01, y10, y11, cy.s, cy.t);
        } else {
            cell_from_global<kSlab>(a, y, x, ucx, ucy, rxx, rxy, ryx, ryy);
        }
        C = D; E = F; G = A_; H = B;
        yd += 1.0;
    }
}

template <class AQ, class AX>
__device__ __noinline__ void advect_edge_tile(const AdvArgs &a, const AQ &q, const AQ &vy, const AX &vx, int x, int ys) {
    pano_adv::advect_march3_body<true, kRows>(a.q_dst, a.vy_dst, a.vx_dst, q, vy, vx, a.h, a.w, a.dt, x, ys, a.yb);
}

// Position in the claim order -> tile.  Border tiles first (first and last tile row, then the first and last tile of every other
// row), interior tiles after them in row-major order: border tiles take the select path and cost 2-3x an interior tile; in plain
// row-major order the last tile ROW of the grid -- all border tiles -- ends the kernel and the SMs that drew them finish long
// after the others (the kernel's fixed cost, measured from a + b x cells over grid sizes, was ~30 us).
__device__ __forceinline__ int tile_at(const AdvArgs &a, int p) {
    const int nx = a.tiles_x, ny = a.tiles_y;
    if (nx < 3 || ny < 3) return p;
    if (p < nx) return p;                                            // first tile row
    if (p < 2 * nx) return (ny - 1) * nx + (p - nx);                 // last tile row
    const int nb = 2 * nx + 2 * (ny - 2);
    if (p < nb) {
        const int q = p - 2 * nx;
        return (1 + (q >> 1)) * nx + ((q & 1) ? nx - 1 : 0);         // first / last tile of rows 1 .. ny-2
    }
    const int q = p - nb;
    return (1 + q / (nx - 2)) * nx + 1 + q % (nx - 2);               // the interior
}

template <bool kSlab>
__global__ void __launch_bounds__(kThreads, 1) k_advect_tma(const __grid_constant__ AdvArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    Tail *tl = reinterpret_cast<Tail *>(smem + kStages * kStageBytes);
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int G = gridDim.x;
    const int ntiles = a.tiles_x * a.tiles_y;
    // no bounded wait here can expire unless the device is broken; the flag exists so that one does not hang a box
    __shared__ unsigned int s_err;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&tl->full[s], 1);
            mbar_init(&tl->empty[s], kConsumerWarps);
        }
        s_err = 0;
        fence_mbar_init();
    }
    __syncthreads();
    volatile unsigned int *err = &s_err;

    if (wid == kConsumerWarps) {
        // ============================================================ producer warp (one lane)
        if (lane != 0) return;
        tma_prefetch_desc(&a.m_q);
        tma_prefetch_desc(&a.m_vy);
        tma_prefetch_desc(&a.m_vx);
        if (blockIdx.x == 0) a.claim[a.parity ^ 1] = 0u;        // the NEXT launch's counter (this one was zeroed by the previous launch)
        unsigned n = 0;
        for (int jj = 0;; ++jj) {
            int t;
            if (a.dynamic) t = (int)atomicAdd(&a.claim[a.parity], 1u);
            else t = blockIdx.x + jj * G;
            if (t < ntiles && a.border_first) t = tile_at(a, t);
            const int st = n % kStages;
            if (!mbar_wait(&tl->empty[st], ((n / kStages) & 1) ^ 1, err)) return;
            if (t >= ntiles) {                                  // end marker
                tl->tile[st] = -1;
                mbar_arrive(&tl->full[st]);
                return;
            }
            tl->tile[st] = t;
            const int tx0 = (t % a.tiles_x) * TW, ty0 = a.ya + (t / a.tiles_x) * TH;
            {   // border tiles are staged as well: TMA zero-fills what lies outside the stored arrays, the clamped indices never go there
                unsigned char *base = smem + st * kStageBytes;
                uint64_t *bar = &tl->full[st];
                mbar_arrive_expect_tx(bar, 2 * kQBytes + 2 * kXBytes);
                const int by = ty0 - kG - a.ylo;                // box row in the stored arrays (even; -2 in the first tile row of the grid)
                tma_load_2d(base, &a.m_q, bar, tx0 - kG, by);
                tma_load_2d(base + kQBytes, &a.m_vy, bar, tx0 - kG, by);
                tma_load_2d(base + 2 * kQBytes, &a.m_vx, bar, tx0 - kG, by >> 1);                       // even rows
                tma_load_2d(base + 2 * kQBytes + kXSlot, &a.m_vx, bar, a.w + 1 + tx0 - kG - 1, by >> 1);   // odd rows, one column to the left
            }
            ++n;
        }
    }

    // ================================================================ consumer warps
    using Acc = typename GlobalAcc<kSlab>::Acc;
    const GlobalAcc<kSlab> ga(a);
    const Acc &gq = ga.q, &gvy = ga.vy, &gvx = ga.vx;
    const int lx = (wid % kColGroups) * 32 + lane, ly0 = (wid / kColGroups) * kRows;
    // First, while the first boxes are in flight: the strips the cell tiles do not cover -- column x = w (vx only) and, where this launch owns it, face row y = h (vy only)
    const int ncol = a.yb - a.ya, nrow = a.yb == a.h ? a.w : 0;
    for (int i = blockIdx.x * kConsumers + tid; i < ncol + nrow; i += G * kConsumers) {
        if (i < ncol) pano_adv::advect_march3_body<true, 1>(a.q_dst, a.vy_dst, a.vx_dst, gq, gvy, gvx, a.h, a.w, a.dt, a.w, a.ya + i, a.ya + i + 1);
        else pano_adv::advect_march3_body<true, 1>(a.q_dst, a.vy_dst, a.vx_dst, gq, gvy, gvx, a.h, a.w, a.dt, i - ncol, a.h, a.h + 1);
    }
    for (unsigned n = 0;; ++n) {
        const int st = n % kStages;
        if (!mbar_wait(&tl->full[st], (n / kStages) & 1, err)) return;
        int t = 0;
        if (lane == 0) t = *(const volatile int *)&tl->tile[st];
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t < 0) break;
        const int tx0 = (t % a.tiles_x) * TW, ty0 = a.ya + (t / a.tiles_x) * TH;
        if (tile_interior(a, ty0, tx0)) {
            const double *Q = reinterpret_cast<const double *>(smem + st * kStageBytes);
            advect_tile<false, kSlab>(a, Q, Q + kQBytes / 8, Q + 2 * (kQBytes / 8), ty0, tx0, ly0, lx, 0u);
        } else {
            const int x = tx0 + lx, ys = ty0 + ly0;
            if (x < a.w && ys < a.yb) {
                const double *Q = reinterpret_cast<const double *>(smem + st * kStageBytes);
                const int by0 = ty0 - kG;
                const unsigned rhq = kSlab ? (unsigned)min(QH, a.whi_q - by0) : (unsigned)QH;
                const unsigned rhvy = kSlab ? (unsigned)min(QH, a.whi_vy - by0) : (unsigned)QH;
                const int ylast = min(a.h, a.yb) - 1;             // last row this launch produces
                if (x >= 1 && x <= a.w - 1 && ys >= 1 && ys + kRows - 1 <= min(a.h - 1, ylast)) {
                    // away from the walls: the shared-memory path; rows_ok = last box row a gather corner may use
                    advect_tile<true, kSlab>(a, Q, Q + kQBytes / 8, Q + 2 * (kQBytes / 8), ty0, tx0, ly0, lx, rhq - 2u);
                } else {
                    const BoxQ<Acc> bq{Q, by0, tx0 - kG, rhq, gq}, bvy{Q + kQBytes / 8, by0, tx0 - kG, rhvy, gvy};
                    const BoxX<Acc> bvx{Q + 2 * (kQBytes / 8), by0, tx0 - kG, rhq, gvx};
                    advect_edge_tile(a, bq, bvy, bvx, x, ys);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&tl->empty[st]);
    }
}

}  // namespace

int pano_preload_advect_tma() {
    cudaFuncAttributes fa;
    PANO_CUDA(cudaFuncGetAttributes(&fa, k_advect_tma<false>));
    PANO_CUDA(cudaFuncSetAttribute(k_advect_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    PANO_CUDA(cudaFuncGetAttributes(&fa, k_advect_tma<true>));
    PANO_CUDA(cudaFuncSetAttribute(k_advect_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    return PANO_OK;
}

// Even sizes (the vx super-rows pair rows; TMA rows are 16-byte aligned), 16-byte aligned arrays, 32-bit indices, and a
// grid large enough for interior tiles to exist.  rows_* = rows stored from global row ylo on.
bool pano_advect_tma_supported(size_t h, size_t w, int ya, int ylo, size_t rows_q, const void *q, const void *vy, const void *vx) {
    if (h % 2 || w % 2 || rows_q % 2 || ((ya - ylo) % 2) != 0) return false;
    if (h < 64 || w < 128) return false;   // smaller grids: every tile is a border tile (correct, but the marching kernel is the better fit)
    if ((h + 1) * (w + 1) >= ((size_t)1 << 31)) return false;
    const void *ps[] = {q, vy, vx};
    for (const void *p : ps)
        if (((uintptr_t)p & 15u) != 0) return false;
    return true;
}

// Sources are self-advected (src == vel).  Pointers are virtual global-row-0 addresses; *_stored point at stored row 0
// (= global row ylo).  err: slab error word or null (one GPU: every row is stored).
int pano_advect_tma_launch(pano_ctx *ctx, double *q_dst, double *vy_dst, double *vx_dst, const double *q_src, const double *vy_src,
                           const double *vx_src, size_t h, size_t w, double dt, int ya, int yb, int ylo, size_t rows_q, size_t rows_vy,
                           unsigned int *err) {
    static_assert(sizeof(Tail) <= kTailBytes, "Tail does not fit");
    static_assert(kQBytes % 128 == 0 && kXSlot % 128 == 0 && kXSlot >= kXBytes, "stage layout");
    PANO_CUDA(cudaFuncSetAttribute(k_advect_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    PANO_CUDA(cudaFuncSetAttribute(k_advect_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    AdvArgs a;
    memset(&a, 0, sizeof(a));
    const double *q_st = q_src + (ptrdiff_t)ylo * (ptrdiff_t)w, *vy_st = vy_src + (ptrdiff_t)ylo * (ptrdiff_t)w,
                 *vx_st = vx_src + (ptrdiff_t)ylo * (ptrdiff_t)(w + 1);
    PANO_TRY(pano_make_tensor_map_2d(&a.m_q, q_st, 8, w, rows_q, (uint64_t)w * 8, QW, QH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_vy, vy_st, 8, w, rows_vy, (uint64_t)w * 8, QW, QH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_vx, vx_st, 8, 2 * (w + 1), rows_q / 2, (uint64_t)(w + 1) * 16, XW, XH));
    a.q_dst = q_dst; a.vy_dst = vy_dst; a.vx_dst = vx_dst;
    a.q_src = q_src; a.vy_src = vy_src; a.vx_src = vx_src;
    a.h = (int)h; a.w = (int)w; a.dt = dt;
    a.ya = ya; a.yb = yb; a.ylo = ylo;
    a.whi_q = ylo + (int)rows_q; a.whi_vy = ylo + (int)rows_vy;
    if (a.whi_q > (int)h) a.whi_q = (int)h;                    // rows beyond the grid are storage, not data
    if (a.whi_vy > (int)h + 1) a.whi_vy = (int)h + 1;
    a.err = err;
    a.tiles_x = ((int)w + TW - 1) / TW;
    a.tiles_y = (yb - ya + TH - 1) / TH;
    if (!ctx->d_adv_claim) {
        PANO_CUDA(cudaMalloc((void **)&ctx->d_adv_claim, 2 * sizeof(unsigned int)));
        PANO_CUDA(cudaMemsetAsync(ctx->d_adv_claim, 0, 2 * sizeof(unsigned int), ctx->stream));
    }
    a.claim = ctx->d_adv_claim;
    a.parity = (int)(ctx->adv_epoch++ & 1u);
    a.dynamic = pano_option(ctx, "advect_dynamic", 1) != 0;
    a.border_first = pano_option(ctx, "advect_border_first", 1) != 0;
    int G = ctx->num_sms;
    const int64_t cap = pano_option(ctx, "advect_ctas", 0);
    if (cap > 0 && cap < G) G = (int)cap;
    const int ntiles = a.tiles_x * a.tiles_y;
    if (G > ntiles) G = ntiles;
    if (err) k_advect_tma<true><<<G, kThreads, kSmemBytes, ctx->stream>>>(a);
    else k_advect_tma<false><<<G, kThreads, kSmemBytes, ctx->stream>>>(a);
    return pano_after_launch(ctx, "advect_tma");
}
