// pano_advect_body.cuh -- the marching form of the fused advection (advect + both loops of advect_mac,
// examples/dec_fluid.rs:173-291), shared by k_advect_march3 (pano_fused.cu: gathers through L1/L2) and by the
// TMA-staged persistent kernel (pano_advect_tma.cu: border tiles, the x = w / y = h strips and backtraces that leave a tile).
#pragma once

#include <type_traits>

#include "pano_cell_math.h"
#include "pano_internal.cuh"

namespace pano_adv {

template <class T>
struct V32 {   // row-major array addressed with 32-bit indices (the host checks that every index fits)
    const T *p;
    int pitch;
    __device__ __forceinline__ T operator()(int y, int x) const { return p[y * pitch + x]; }
};

// The same for a slab of a larger grid: `p` is the VIRTUAL address of global row 0 and only rows [lo, hi) are stored.
// A gather that leaves the stored window (backtrace longer than the ghost zone) raises *err and is clamped into the
// window so that it cannot fault.
struct V32W {
    const double *p;
    int pitch, lo, hi;
    unsigned int *err;
    __device__ __forceinline__ double operator()(int y, int x) const {
        if (y < lo || y >= hi) {
            *err = 2u;
            y = y < lo ? lo : hi - 1;
        }
        return p[y * pitch + x];
    }
};

// a backtrace beyond 2^32 cells: the general form (never in practice)
template <class A>
__device__ __noinline__ double mac_gather_far(double relx, double rely, int H, int W, A q) {
    return pano::mac_gather<double>(relx, rely, H, W, q);
}

// One thread, one column x, rows ys .. ys+kRows-1 of the index space (h+1) x (w+1): the eight velocity samples around
// (y, x) are carried in registers (4 new loads per row), indices are 32-bit.  Arithmetic: the exact fast forms of
// pano_cell_math.h (floor by a round-down add, one clamp per axis instead of three, no 64-bit conversions), positions
// carried as doubles and advanced by +1.0 (exact).  kEdge = false: the caller guarantees 1 <= x <= w-1 and
// 1 <= ys, ys+kRows-1 <= h-1, so every border select disappears.  kEdge = true: rows from yend on are skipped.
template <bool kEdge, int kRows, class A, class AX>
__device__ __forceinline__ void advect_march3_body(double *__restrict__ q_dst, double *__restrict__ vy_dst, double *__restrict__ vx_dst,
                                                   const A &q, const A &vy, const AX &vx, int h, int w, double dt, int x, int ys, int yend) {
    const bool xin = !kEdge || x < w, xpos = !kEdge || x > 0;
    const double ndt = -dt, xd = (double)x, xh = xd + 0.5;
    const double wlim = (double)w - 1.00001, hlim = (double)h - 1.00001;
    double yd = (double)ys;
    double C = xin ? vy(ys, x) : 0.0, E = xpos ? vy(ys, x - 1) : 0.0;
    double G = 0.0, H = 0.0;
    if (!kEdge || ys > 0) {
        G = vx(ys - 1, x);
        H = xin ? vx(ys - 1, x + 1) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const int y = ys + k;
        if (kEdge && (y > h || y >= yend)) break;
        const bool yin = !kEdge || y < h;
        const double yh = yd + 0.5;
        double A_ = 0.0, B = 0.0, D = 0.0, F = 0.0;
        if (yin) {
            A_ = vx(y, x);
            if (xin) { B = vx(y, x + 1); D = vy(y + 1, x); }
            if (xpos) F = vy(y + 1, x - 1);
        }
        // the three backtraced positions of this row: advect (dec_fluid.rs:180-183), advect_mac x (:220-225) and y (:257-263)
        double vvy, vvx;
        if (kEdge) {
            const bool ypos = y > 0;
            const double t0 = xin ? C : E, t1 = xin ? D : F, t2 = xpos ? E : C, t3 = xpos ? F : D;
            vvy = (t0 + t1 + t2 + t3) / 4.0;
            const double u0 = yin ? A_ : G, u1 = yin ? B : H, u2 = ypos ? G : A_, u3 = ypos ? H : B;
            vvx = (u0 + u1 + u2 + u3) / 4.0;
        } else {
            vvy = (C + D + E + F) / 4.0;
            vvx = (A_ + B + G + H) / 4.0;
        }
        const pano::CellCoord cq = pano::advect_coord_fast(xh, yh, wlim, hlim, ndt, (A_ + B) / 2.0, (C + D) / 2.0);
        double rxx, rxy, ryx, ryy;
        pano::mac_x_rel(xd, yh, ndt, A_, vvy, rxx, rxy);
        pano::mac_y_rel(xh, yd, ndt, vvx, C, ryx, ryy);
        const pano::MacCoord cx = pano::mac_coord_fast(rxx, rxy, h, w + 1), cy = pano::mac_coord_fast(ryx, ryy, h + 1, w);
        if ((cx.bad | cy.bad) == 0u) {
            // one straight-line block: all twelve gathers can be in flight together
            const double vq = (yin && xin) ? pano::advect_gather_at(cq, q) : 0.0;
            const double vxn = yin ? pano::mac_gather_at(cx, vx) : 0.0;
            const double vyn = xin ? pano::mac_gather_at(cy, vy) : 0.0;
            if (yin && xin) q_dst[y * w + x] = vq;
            if (yin) vx_dst[y * (w + 1) + x] = vxn;
            if (xin) vy_dst[y * w + x] = vyn;
        } else {                                             // a backtrace beyond 2^32 cells: the general form
            if (yin && xin) q_dst[y * w + x] = pano::advect_gather_at(cq, q);
            if (yin) vx_dst[y * (w + 1) + x] = mac_gather_far(rxx, rxy, h, w + 1, vx);
            if (xin) vy_dst[y * w + x] = mac_gather_far(ryx, ryy, h + 1, w, vy);
        }
        C = D; E = F; G = A_; H = B;
        yd += 1.0;
    }
}

}  // namespace pano_adv
