// pano_fused.cu -- the non-solver passes of the step, each fused into one kernel:
//   K1 advect_all      examples/dec_fluid.rs:59-66   (advect + both advect_mac loops, one pass)
//   K3 neg_divergence  examples/dec_fluid.rs:69-83   (hodge_1_dual -> box zero -> d1 -> negate, + max|b|)
//   K5 laplacian_apply examples/dec_fluid.rs:100-119 (stand-alone form of the CG operator)
//   K8 project         examples/dec_fluid.rs:124-141 (hodge_2 -> d0_dual -> scaled_add -> walls)
// All are HBM-bound; see DESIGN.md for bytes per cell.
#include "pano_advect_body.cuh"
#include "pano_cell_math.h"
#include "pano_internal.cuh"

bool pano_advect_tma_supported(size_t h, size_t w, int ya, int ylo, size_t rows_q, const void *q, const void *vy, const void *vx);
int pano_advect_tma_launch(pano_ctx *ctx, double *q_dst, double *vy_dst, double *vx_dst, const double *q_src, const double *vy_src,
                           const double *vx_src, size_t h, size_t w, double dt, int ya, int yb, int ylo, size_t rows_q, size_t rows_vy,
                           unsigned int *err);
int pano_preload_advect_tma();

namespace {

constexpr int kThreads = 256;

template <class T>
struct View2 {   // row-major (rows, pitch) array in global memory
    const T *p;
    int pitch;
    __device__ __forceinline__ T operator()(int y, int x) const { return p[(size_t)y * pitch + x]; }
};

// Same, for a slab of a larger grid: `p` is the VIRTUAL address of global row 0 and only rows
// [lo, hi) are stored.  A gather that leaves the stored window (backtrace longer than the ghost
// zone) raises *err and is clamped into the window so that it cannot fault.
template <class T>
struct View2W {
    const T *p;
    int pitch, lo, hi;
    unsigned int *err;
    __device__ __forceinline__ T operator()(int y, int x) const {
        if (y < lo || y >= hi) {
            *err = 2u;
            y = y < lo ? lo : hi - 1;
        }
        return p[(size_t)y * pitch + x];
    }
};

// ------------------------------------------------------------------ K1: advection
// Tile = 32 x 8 cells per 256-thread block; index space (h+1) x (w+1) covers q (h,w),
// vx (h,w+1) and vy (h+1,w) in one sweep.  Gathers go through L1/L2 (the backtrace lands
// within ~2 cells of the thread's own cell for CFL-bounded flow).
// Rows [ya, yb) of the cell grid are produced (the whole grid on one GPU, a slab otherwise);
// the vy face row y belongs to the slab that owns cell row y, and face row h to the last slab.
template <class T, bool kScalar, bool kMac, class VQ, class VY, class VX>
__device__ __forceinline__ void advect_rows(T *q_dst, T *vy_dst, T *vx_dst, const VQ &q, const VY &qy, const VX &qx,
                                            const VY &vy, const VX &vx, int h, int w, T dt, int ya, int yb) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    if (x > w) return;
    const int yf = yb == h ? h + 1 : yb;   // end of the vy face rows of this slab
    for (int y = ya + blockIdx.y * 8 + (threadIdx.x >> 5); y < yf; y += gridDim.y * 8) {
        if (kScalar && y < yb && x < w) q_dst[(size_t)y * w + x] = pano::advect_cell<T>(y, x, h, w, dt, q, vy, vx);
        if (kMac) {
            if (y < yb) vx_dst[(size_t)y * (w + 1) + x] = pano::advect_mac_x<T>(y, x, h, w, dt, qx, vy, vx);
            if (x < w) vy_dst[(size_t)y * w + x] = pano::advect_mac_y<T>(y, x, h, w, dt, qy, vy, vx);
        }
    }
}

template <class T, bool kScalar, bool kMac>
__global__ void __launch_bounds__(kThreads)
k_advect(T *__restrict__ q_dst, T *__restrict__ vel_dst, const T *__restrict__ q_src, const T *__restrict__ mac_src,
         const T *__restrict__ vel, int h, int w, T dt) {
    const size_t off = (size_t)w * (h + 1);
    const View2<T> vy{vel, w}, vx{vel + off, w + 1};
    const View2<T> qy{mac_src, w}, qx{mac_src + off, w + 1};
    const View2<T> q{q_src, w};
    advect_rows<T, kScalar, kMac>(q_dst, vel_dst, vel_dst + off, q, qy, qx, vy, vx, h, w, dt, 0, h);
}

// slab form (self-advection): separate component pointers, all of them virtual global-row-0 addresses;
// cell rows [wlo, whi) and face rows [wlo, whi + 1) are stored
__global__ void __launch_bounds__(kThreads)
k_advect_slab(double *q_dst, double *vy_dst, double *vx_dst, const double *q_src, const double *vy_src, const double *vx_src,
              int h, int w, double dt, int ya, int yb, int wlo, int whi, unsigned int *err) {
    // face row whi is stored but only rows below it are ever received from the neighbour; face row h exists on the last slab
    const View2W<double> vy{vy_src, w, wlo, whi > h + 1 ? h + 1 : whi, err}, vx{vx_src, w + 1, wlo, whi > h ? h : whi, err};
    const View2W<double> q{q_src, w, wlo, whi > h ? h : whi, err};
    advect_rows<double, true, true>(q_dst, vy_dst, vx_dst, q, vy, vx, vy, vx, h, w, dt, ya, yb);
}

// ------------------------------------------------------------------ K1, second generation
// The first kernel above is issue-bound (ncu: 13 warp instructions per cell, 71 % issue-active, DRAM at 44 %):
// 64-bit index arithmetic for 26 loads per cell and no reuse of the velocity samples shared by the three
// advected quantities.  Here a thread owns one column and marches down kRows rows: the eight velocity samples
// around (y, x) are carried in registers (4 new loads per row), indices are 32-bit, clamps are fmin/fmax.
// Arithmetic is pano_cell_math.h's, so results stay bit-identical.  (Self-advection only: src == vel.)
using pano_adv::V32;

constexpr int kAdvRows = 4;   // rows per thread; a 256-thread block covers 32 columns x 32 rows

template <class T>
__global__ void __launch_bounds__(kThreads)
k_advect_march(T *__restrict__ q_dst, T *__restrict__ vy_dst, T *__restrict__ vx_dst, const T *__restrict__ q_src,
               const T *__restrict__ vy_src, const T *__restrict__ vx_src, int h, int w, T dt) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ys = (blockIdx.y * 8 + (threadIdx.x >> 5)) * kAdvRows;
    if (x > w || ys > h) return;
    const V32<T> q{q_src, w}, vy{vy_src, w}, vx{vx_src, w + 1};
    const bool xin = x < w, xpos = x > 0;
    // rows of vy at y (C at x, E at x-1) and of vx at y-1 (G at x, H at x+1)
    T C = xin ? vy(ys, x) : (T)0, E = xpos ? vy(ys, x - 1) : (T)0;
    T G = (T)0, H = (T)0;
    if (ys > 0) {
        G = vx(ys - 1, x);
        H = xin ? vx(ys - 1, x + 1) : (T)0;
    }
#pragma unroll
    for (int k = 0; k < kAdvRows; ++k) {
        const int y = ys + k;
        if (y > h) break;
        const bool yin = y < h;
        T A = (T)0, B = (T)0, D = (T)0, F = (T)0;
        if (yin) {
            A = vx(y, x);
            if (xin) { B = vx(y, x + 1); D = vy(y + 1, x); }
            if (xpos) F = vy(y + 1, x - 1);
        }
        if (yin && xin) {                                   // advect (dec_fluid.rs:180-183)
            const T ucx = (A + B) / (T)2, ucy = (C + D) / (T)2;
            q_dst[y * w + x] = pano::advect_cell_uv<T>(y, x, h, w, dt, ucx, ucy, q);
        }
        if (yin) {                                          // advect_mac, x component (:220-225): xc = min(x, w-1), xm = max(x-1, 0)
            const T t0 = xin ? C : E, t1 = xin ? D : F, t2 = xpos ? E : C, t3 = xpos ? F : D;
            const T vvy = (t0 + t1 + t2 + t3) / (T)4;
            vx_dst[y * (w + 1) + x] = pano::advect_mac_x_uv<T>(y, x, h, w, dt, A, vvy, vx);
        }
        if (xin) {                                          // advect_mac, y component (:257-263): yc = min(y, h-1), ym = max(y-1, 0)
            const bool ypos = y > 0;
            const T t0 = yin ? A : G, t1 = yin ? B : H, t2 = ypos ? G : A, t3 = ypos ? H : B;
            const T vvx = (t0 + t1 + t2 + t3) / (T)4;
            vy_dst[y * w + x] = pano::advect_mac_y_uv<T>(y, x, h, w, dt, vvx, C, vy);
        }
        C = D; E = F; G = A; H = B;
    }
}

// ------------------------------------------------------------------ K1, third generation (f64)
// k_advect_march is issue-bound at ~370 instructions per cell (ncu: 63 % issue-active, DRAM 42 % of the copy peak).
// Same marching scheme with the exact fast forms: pano_advect_body.cuh.
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// kPrefetch: 0 none, 1 L1, 2 L2, 3 L1 + the tile of the block `ahead` launches further on into L2 -- every row of q, vy, vx
// the block will touch first is requested at kernel entry, so the DRAM latency of the velocity loads AND of the
// first-touch gathers (two dependent round trips per row otherwise) overlaps across the whole tile instead of being
// paid row by row.
template <int kRows, int kPrefetch, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_advect_march3(double *__restrict__ q_dst, double *__restrict__ vy_dst, double *__restrict__ vx_dst, const double *__restrict__ q_src,
                const double *__restrict__ vy_src, const double *__restrict__ vx_src, int h, int w, double dt, int ahead) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ys = (blockIdx.y * 8 + (threadIdx.x >> 5)) * kRows;
    // block-uniform: columns [bx0, bx0+32) within [1, w-1] and rows [by0, by0 + 8*kRows) within [1, h-1]
    const int bx0 = blockIdx.x * 32, by0 = blockIdx.y * 8 * kRows;
    if (kPrefetch == 3) {
        const int nb = blockIdx.y * gridDim.x + blockIdx.x + ahead;
        const int px = (nb % gridDim.x) * 32 + (threadIdx.x & 31), py = ((nb / gridDim.x) * 8 + (threadIdx.x >> 5)) * kRows;
        if (px < w && py + kRows < h) {
#pragma unroll
            for (int k = 0; k < kRows; ++k) {
                prefetch_l2(q_src + (py + k) * w + px);
                prefetch_l2(vy_src + (py + k + 1) * w + px);
                prefetch_l2(vx_src + (py + k) * (w + 1) + px);
            }
        }
    }
    if (bx0 >= 1 && bx0 + 31 <= w - 1 && by0 >= 1 && by0 + 8 * kRows - 1 <= h - 1) {
        if (kPrefetch) {
#pragma unroll
            for (int k = 0; k < kRows; ++k) {
                const double *a = q_src + (ys + k) * w + x, *b = vy_src + (ys + k + 1) * w + x, *c = vx_src + (ys + k) * (w + 1) + x;
                if (kPrefetch == 2) { prefetch_l2(a); prefetch_l2(b); prefetch_l2(c); }
                else { prefetch_l1(a); prefetch_l1(b); prefetch_l1(c); }
            }
        }
        pano_adv::advect_march3_body<false, kRows>(q_dst, vy_dst, vx_dst, V32<double>{q_src, w}, V32<double>{vy_src, w}, V32<double>{vx_src, w + 1}, h, w, dt, x, ys, h + 1);
    } else {
        if (x > w || ys > h) return;
        pano_adv::advect_march3_body<true, kRows>(q_dst, vy_dst, vx_dst, V32<double>{q_src, w}, V32<double>{vy_src, w}, V32<double>{vx_src, w + 1}, h, w, dt, x, ys, h + 1);
    }
}

// The marching kernel on a slab: cell rows [ya, yb) (face rows up to yf), sources windowed (rows [wlo, whi), +1 for vy).
__global__ void __launch_bounds__(kThreads, 4)
k_advect_march3_slab(double *__restrict__ q_dst, double *__restrict__ vy_dst, double *__restrict__ vx_dst, const double *__restrict__ q_src,
                     const double *__restrict__ vy_src, const double *__restrict__ vx_src, int h, int w, double dt, int ya, int yb, int wlo,
                     int whi, unsigned int *err) {
    using pano_adv::V32W;
    const int yf = yb == h ? h + 1 : yb;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ys = ya + (blockIdx.y * 8 + (threadIdx.x >> 5)) * kAdvRows;
    const int bx0 = blockIdx.x * 32, by0 = ya + blockIdx.y * 8 * kAdvRows;
    const V32W q{q_src, w, wlo, whi > h ? h : whi, err}, vy{vy_src, w, wlo, whi + 1 > h + 1 ? h + 1 : whi + 1, err},
        vx{vx_src, w + 1, wlo, whi > h ? h : whi, err};
    if (bx0 >= 1 && bx0 + 31 <= w - 1 && by0 >= 1 && by0 + 8 * kAdvRows - 1 <= h - 1 && by0 + 8 * kAdvRows <= yb) {
        pano_adv::advect_march3_body<false, kAdvRows>(q_dst, vy_dst, vx_dst, q, vy, vx, h, w, dt, x, ys, yf);
    } else {
        if (x > w || ys >= yf) return;
        pano_adv::advect_march3_body<true, kAdvRows>(q_dst, vy_dst, vx_dst, q, vy, vx, h, w, dt, x, ys, yf);
    }
}

// ------------------------------------------------------------------ K3, second generation (no reductions: the
// solver computes max|b| and b.b itself): one column x kDivRows rows per thread, vy carried down in a register
// option "fused_wide": 0 / 8 / 16 (one GPU, f64).  Measured on B200 (4096^2 / 8192^2): -div 69.8 / 264 us -> 67.9 / 256 us with 8,
// projection 126.5 / 488 -> 124.9 / 481 us; 16 rows is no better.  Default: 8 from 1024 columns on.
constexpr int kFusedWideMinCols = 1024;
constexpr int kDivRows = 8;
// kWide = 0: a block is 32 columns x (8 warps x kDivRows rows).  kWide = R > 0: a block is 256 CONSECUTIVE columns x R rows (every warp
// 32 of them): 2 KB contiguous per row and array instead of 256 B, which the HBM pages like better.
template <class T, int kWide>
__global__ void __launch_bounds__(kThreads)
k_neg_divergence_march(T *__restrict__ b, const T *__restrict__ vy, const T *__restrict__ vx, int h, int w, RectI m, int ya, int yb) {
    // cell rows [ya, yb) (the whole grid on one GPU; a slab, whose pointers are virtual row-0 addresses, otherwise)
    constexpr int kRows = kWide > 0 ? kWide : kDivRows;
    const int x = kWide > 0 ? blockIdx.x * kThreads + threadIdx.x : blockIdx.x * 32 + (threadIdx.x & 31);
    const int ys = ya + (kWide > 0 ? blockIdx.y : blockIdx.y * 8 + (threadIdx.x >> 5)) * kRows;
    if (x >= w || ys >= yb) return;
    // block-uniform test: does this block touch the obstacle's edges?
    const int bw = kWide > 0 ? kThreads : 32, bh = kWide > 0 ? kRows : 8 * kRows;
    const int bx0 = blockIdx.x * bw, by0 = ya + blockIdx.y * bh;
    const bool masked = m.y1 > m.y0 && m.x1 > m.x0 && by0 < m.y1 && by0 + bh + 1 > m.y0 && bx0 < m.x1 && bx0 + bw + 1 > m.x0;
    T vy0 = vy[ys * w + x];
    if (masked && in_rect(m, ys, x)) vy0 = (T)0;
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const int y = ys + k;
        if (y >= yb) break;
        T vy1 = vy[(y + 1) * w + x], vx0 = vx[y * (w + 1) + x], vx1 = vx[y * (w + 1) + x + 1];
        if (masked) {
            if (in_rect(m, y + 1, x)) vy1 = (T)0;
            if (in_rect(m, y, x)) vx0 = (T)0;
            if (in_rect(m, y, x + 1)) vx1 = (T)0;
        }
        b[y * w + x] = pano::neg_divergence_cell<T>(vy0, vy1, vx0, vx1);
        vy0 = vy1;
    }
}

// ------------------------------------------------------------------ K3: -divergence (+ max|b|, b.b)
template <class T>
__global__ void __launch_bounds__(kThreads)
k_neg_divergence(T *__restrict__ b, const T *__restrict__ vy, const T *__restrict__ vx, int h, int w, RectI m, int ya, int yb,
                 double *__restrict__ partial_max, double *__restrict__ partial_dot) {
    __shared__ T scratch[32];
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    T amax = 0, adot = 0;
    if (x < w) {
        for (int y = ya + blockIdx.y * 8 + (threadIdx.x >> 5); y < yb; y += gridDim.y * 8) {
            T vy0 = in_rect(m, y, x) ? (T)0 : vy[(size_t)y * w + x];
            T vy1 = in_rect(m, y + 1, x) ? (T)0 : vy[(size_t)(y + 1) * w + x];
            T vx0 = in_rect(m, y, x) ? (T)0 : vx[(size_t)y * (w + 1) + x];
            T vx1 = in_rect(m, y, x + 1) ? (T)0 : vx[(size_t)y * (w + 1) + x + 1];
            T v = pano::neg_divergence_cell<T>(vy0, vy1, vx0, vx1);
            b[(size_t)y * w + x] = v;
            T a = v < 0 ? -v : v;
            amax = a > amax ? a : amax;
            adot = adot + v * v;
        }
    }
    T bm = block_max(amax, scratch);
    T bd = block_sum(adot, scratch);
    if (threadIdx.x == 0 && partial_max) {
        const int bid = blockIdx.y * gridDim.x + blockIdx.x;
        partial_max[bid] = (double)bm;
        partial_dot[bid] = (double)bd;
    }
}

__global__ void k_reduce_max_dot(const double *__restrict__ pmax, const double *__restrict__ pdot, int n,
                                 double *__restrict__ out) {
    __shared__ double scratch[32];
    double m = 0, d = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        m = pmax[i] > m ? pmax[i] : m;
        d += pdot[i];
    }
    m = block_max(m, scratch);
    d = block_sum(d, scratch);
    if (threadIdx.x == 0) {
        out[0] = m;
        out[1] = d;
    }
}

// ------------------------------------------------------------------ K5: stand-alone Laplacian
template <class T>
__global__ void __launch_bounds__(kThreads)
k_laplacian(T *__restrict__ z, const T *__restrict__ p, int h, int w, T dt, RectI m) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    if (x >= w) return;
    for (int y = blockIdx.y * 8 + (threadIdx.x >> 5); y < h; y += gridDim.y * 8) {
        const bool oN = y > 0 && !in_rect(m, y, x), oS = y < h - 1 && !in_rect(m, y + 1, x);
        const bool oW = x > 0 && !in_rect(m, y, x), oE = x < w - 1 && !in_rect(m, y, x + 1);
        const size_t i = (size_t)y * w + x;
        const T c = p[i];
        const T n = oN ? p[i - w] : (T)0, s = oS ? p[i + w] : (T)0;
        const T l = oW ? p[i - 1] : (T)0, r = oE ? p[i + 1] : (T)0;
        z[i] = pano::laplacian_cell<T>(c, n, s, l, r, oN, oS, oW, oE, dt);
    }
}

// ------------------------------------------------------------------ K8: projection + walls
template <class T>
__global__ void __launch_bounds__(kThreads)
k_project(T *__restrict__ vy, T *__restrict__ vx, const T *__restrict__ p, int h, int w, T dt, int ya, int yb) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    if (x > w) return;
    const int yf = yb == h ? h + 1 : yb;   // vy face rows [ya, yf), vx rows [ya, yb)
    for (int y = ya + blockIdx.y * 8 + (threadIdx.x >> 5); y < yf; y += gridDim.y * 8) {
        const T c = (y < h && x < w) ? p[(size_t)y * w + x] : (T)0;
        if (x < w) {   // vy[y, x]
            const size_t i = (size_t)y * w + x;
            if (y == 0 || y == h) vy[i] = (T)0;                                   // walls :137-140
            else vy[i] = vy[i] + dt * (-(c - p[(size_t)(y - 1) * w + x]));        // d0_dual :322-326, scaled_add :126
        }
        if (y < yb) {   // vx[y, x]
            const size_t i = (size_t)y * (w + 1) + x;
            if (x == 0 || x == w) vx[i] = (T)0;                                   // walls :132-135
            else vx[i] = vx[i] + dt * (p[(size_t)y * w + x - 1] - c);             // d0_dual :329-333
        }
    }
}

// K8, second generation: one column x kProjRows rows per thread, p carried down the column in a register (each p value is
// loaded once for vy and once per neighbour column for vx instead of three times), 32-bit indices
constexpr int kProjRows = 8;
template <class T, int kWide>   // kWide as k_neg_divergence_march
__global__ void __launch_bounds__(kThreads)
k_project_march(T *__restrict__ vy, T *__restrict__ vx, const T *__restrict__ p, int h, int w, T dt, int ya, int yb) {
    // vx rows [ya, yb) and vy face rows [ya, yf): the face row y belongs to the slab that owns cell row y, face row h to the last one
    constexpr int kRows = kWide > 0 ? kWide : kProjRows;
    const int yf = yb == h ? h + 1 : yb;
    const int x = kWide > 0 ? blockIdx.x * kThreads + threadIdx.x : blockIdx.x * 32 + (threadIdx.x & 31);
    const int ys = ya + (kWide > 0 ? blockIdx.y : blockIdx.y * 8 + (threadIdx.x >> 5)) * kRows;
    if (x > w || ys >= yf) return;
    const bool xin = x < w;
    T pn = (xin && ys > 0) ? p[(ys - 1) * w + x] : (T)0;     // p[y-1, x]
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const int y = ys + k;
        if (y >= yf) break;
        const T c = (y < h && xin) ? p[y * w + x] : (T)0;
        if (xin) {   // vy[y, x]
            const int i = y * w + x;
            if (y == 0 || y == h) vy[i] = (T)0;                                   // walls :137-140
            else vy[i] = vy[i] + dt * (-(c - pn));                               // d0_dual :322-326, scaled_add :126
        }
        if (y < h) {   // vx[y, x]
            const int i = y * (w + 1) + x;
            if (x == 0 || x == w) vx[i] = (T)0;                                   // walls :132-135
            else vx[i] = vx[i] + dt * (p[y * w + x - 1] - c);                     // d0_dual :329-333
        }
        pn = c;
    }
}

template <class T>
__global__ void k_to_u8(uint8_t *__restrict__ out, const T *__restrict__ d, int h, int w, T lower, T upper) {
    const size_t n = (size_t)h * w;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / w), x = (int)(i % w);
        T v = d[(size_t)(h - 1 - y) * w + x];                    // vertical flip, png.rs:12
        v = pano::tmax(pano::tmin(v, upper), lower);             // imgproc.rs:3
        T s = (v - lower) / (upper - lower) * (T)255;            // imgproc.rs:4
        out[i] = (uint8_t)(s < (T)0 ? (T)0 : (s > (T)255 ? (T)255 : s));   // Rust `as u8` saturates
    }
}

inline dim3 grid2d(int rows, int cols) {
    int gy = (rows + 7) / 8;
    if (gy > 8192) gy = 8192;
    if (gy < 1) gy = 1;
    int gx = (cols + 31) / 32;
    if (gx < 1) gx = 1;
    return dim3((unsigned)gx, (unsigned)gy);
}

}  // namespace

// internal: advection on raw pointers (used by pano_step.cu as well)
int pano_advect_launch(pano_ctx *ctx, int dtype, void *q_dst, void *vel_dst, const void *q_src, const void *mac_src,
                       const void *vel, size_t h, size_t w, double dt) {
    const bool sc = q_dst != nullptr, mac = vel_dst != nullptr;
    const size_t off1 = w * (h + 1);
    // "advect_kernel": 0 auto, 1 k_advect (any shape / dtype, separate advect / advect_mac), 2 k_advect_march, 3 k_advect_march3,
    // 4 k_advect_tma (pano_advect_tma.cu; f64, even sizes).  auto: the TMA kernel once the grid gives every SM a few tiles
    // (4 x 148 tiles of 32 x 64 cells, ~1.2 Mcell), the marching kernel below that.
    const int64_t ak = pano_option(ctx, "advect_kernel", 0);
    if (sc && mac && mac_src == vel && dtype == PANO_F64 && (ak == 0 || ak == 4) &&
        pano_advect_tma_supported(h, w, 0, 0, h, q_src, vel, (const double *)vel + off1) &&
        (ak == 4 || ((h + 31) / 32) * ((w + 63) / 64) >= (size_t)4 * ctx->num_sms)) {
        return pano_advect_tma_launch(ctx, (double *)q_dst, (double *)vel_dst, (double *)vel_dst + off1, (const double *)q_src,
                                      (const double *)vel, (const double *)vel + off1, h, w, dt, 0, (int)h, 0, h, h + 1, nullptr);
    }
    // the marching kernel: both outputs, self-advection, 32-bit indices
    if (sc && mac && mac_src == vel && (h + 1) * (w + 1) < ((size_t)1 << 31) && ak != 1) {
        dim3 gm((unsigned)((w + 1 + 31) / 32), (unsigned)((h + 1 + 8 * kAdvRows - 1) / (8 * kAdvRows)));
        if (dtype == PANO_F64 && ak != 2) {
            const int pf = (int)pano_option(ctx, "advect_prefetch", 1), mb = (int)pano_option(ctx, "advect_minblocks", 4);
            const int rows = (int)pano_option(ctx, "advect_rows", 4);
            if (rows != 2 && rows != 4 && rows != 8)   // the launch grid is sized from it; only these are instantiated
                PANO_FAIL(PANO_ERR_INVALID, "option advect_rows = %d: the marching kernel exists for 2, 4 and 8 rows per thread", rows);
            if (pf < 0 || pf > 3 || mb < 3 || mb > 5)
                PANO_FAIL(PANO_ERR_INVALID, "option advect_prefetch = %d (0..3) / advect_minblocks = %d (3..5) out of range", pf, mb);
            const int ahead = ctx->num_sms * (int)pano_option(ctx, "advect_ahead", mb);
            dim3 g3((unsigned)((w + 1 + 31) / 32), (unsigned)((h + 1 + 8 * rows - 1) / (8 * rows)));
#define PANO_ADV3(R, PF, MB)                                                                                                       \
    k_advect_march3<R, PF, MB><<<g3, kThreads, 0, ctx->stream>>>((double *)q_dst, (double *)vel_dst, (double *)vel_dst + off1, (const double *)q_src, \
                                                                 (const double *)vel, (const double *)vel + off1, (int)h, (int)w, dt, ahead)
#define PANO_ADV3_PF(R, MB)                                                                                     \
    do {                                                                                                        \
        if (pf == 0) PANO_ADV3(R, 0, MB); else if (pf == 1) PANO_ADV3(R, 1, MB); else if (pf == 2) PANO_ADV3(R, 2, MB); else PANO_ADV3(R, 3, MB); \
    } while (0)
            if (rows == 8) { if (mb >= 4) PANO_ADV3_PF(8, 4); else PANO_ADV3_PF(8, 3); }
            else if (rows == 2) { if (mb >= 5) PANO_ADV3_PF(2, 5); else PANO_ADV3_PF(2, 4); }
            else if (mb >= 5) PANO_ADV3_PF(4, 5);
            else if (mb == 4) PANO_ADV3_PF(4, 4);
            else PANO_ADV3_PF(4, 3);
#undef PANO_ADV3_PF
#undef PANO_ADV3
            return pano_after_launch(ctx, "advect_march3");
        }
        if (dtype == PANO_F64)
            k_advect_march<double><<<gm, kThreads, 0, ctx->stream>>>((double *)q_dst, (double *)vel_dst, (double *)vel_dst + off1,
                                                                     (const double *)q_src, (const double *)vel, (const double *)vel + off1,
                                                                     (int)h, (int)w, dt);
        else
            k_advect_march<float><<<gm, kThreads, 0, ctx->stream>>>((float *)q_dst, (float *)vel_dst, (float *)vel_dst + off1,
                                                                    (const float *)q_src, (const float *)vel, (const float *)vel + off1,
                                                                    (int)h, (int)w, (float)dt);
        return pano_after_launch(ctx, "advect_march");
    }
    dim3 g = grid2d((int)h + 1, (int)w + 1);
#define PANO_LAUNCH_ADV(T, S, M)                                                                                     \
    k_advect<T, S, M><<<g, kThreads, 0, ctx->stream>>>((T *)q_dst, (T *)vel_dst, (const T *)q_src, (const T *)mac_src, \
                                                       (const T *)vel, (int)h, (int)w, (T)dt)
    if (dtype == PANO_F64) {
        if (sc && mac) PANO_LAUNCH_ADV(double, true, true);
        else if (sc) PANO_LAUNCH_ADV(double, true, false);
        else PANO_LAUNCH_ADV(double, false, true);
    } else {
        if (sc && mac) PANO_LAUNCH_ADV(float, true, true);
        else if (sc) PANO_LAUNCH_ADV(float, true, false);
        else PANO_LAUNCH_ADV(float, false, true);
    }
#undef PANO_LAUNCH_ADV
    return pano_after_launch(ctx, "advect");
}

// internal: b = -div(vel).  want_scalars: also d_scalars[0] = max|b|, d_scalars[1] = b.b (no sync); the
// step does not need them because the CG kernel reduces max|b| and b.b itself in its first pass.
int pano_neg_divergence_launch(pano_ctx *ctx, int dtype, void *b, const void *vel, size_t h, size_t w, pano_rect obstacle,
                               bool want_scalars) {
    // the obstacle rectangle indexes vy (h+1, w) and vx (h, w+1) alike; clip to the union
    RectI m = pano_clip_rect(obstacle, h + 1, w + 1);
    const size_t off = w * (h + 1);
    if (!want_scalars && (h + 1) * (w + 1) < ((size_t)1 << 31)) {
        const int64_t wide = dtype == PANO_F64 ? pano_option(ctx, "fused_wide", w >= (size_t)kFusedWideMinCols ? 8 : 0) : 0;
        dim3 gm((unsigned)((w + 31) / 32), (unsigned)((h + 8 * kDivRows - 1) / (8 * kDivRows)));
        if (wide == 16) {
            dim3 gw((unsigned)((w + kThreads - 1) / kThreads), (unsigned)((h + 15) / 16));
            k_neg_divergence_march<double, 16><<<gw, kThreads, 0, ctx->stream>>>((double *)b, (const double *)vel, (const double *)vel + off, (int)h, (int)w, m, 0, (int)h);
        } else if (wide == 8) {
            dim3 gw((unsigned)((w + kThreads - 1) / kThreads), (unsigned)((h + 7) / 8));
            k_neg_divergence_march<double, 8><<<gw, kThreads, 0, ctx->stream>>>((double *)b, (const double *)vel, (const double *)vel + off, (int)h, (int)w, m, 0, (int)h);
        } else if (dtype == PANO_F64)
            k_neg_divergence_march<double, 0><<<gm, kThreads, 0, ctx->stream>>>((double *)b, (const double *)vel, (const double *)vel + off, (int)h, (int)w, m, 0, (int)h);
        else
            k_neg_divergence_march<float, 0><<<gm, kThreads, 0, ctx->stream>>>((float *)b, (const float *)vel, (const float *)vel + off, (int)h, (int)w, m, 0, (int)h);
        return pano_after_launch(ctx, "neg_divergence_march");
    }
    dim3 g = grid2d((int)h, (int)w);
    const int nb = (int)(g.x * g.y);
    PANO_TRY(pano_ensure_partials(ctx, 2 * (size_t)nb));
    if (dtype == PANO_F64)
        k_neg_divergence<double><<<g, kThreads, 0, ctx->stream>>>((double *)b, (const double *)vel, (const double *)vel + off, (int)h,
                                                                  (int)w, m, 0, (int)h, ctx->d_partials, ctx->d_partials + nb);
    else
        k_neg_divergence<float><<<g, kThreads, 0, ctx->stream>>>((float *)b, (const float *)vel, (const float *)vel + off, (int)h,
                                                                 (int)w, m, 0, (int)h, ctx->d_partials, ctx->d_partials + nb);
    PANO_TRY(pano_after_launch(ctx, "neg_divergence"));
    if (!want_scalars) return PANO_OK;
    k_reduce_max_dot<<<1, kThreads, 0, ctx->stream>>>(ctx->d_partials, ctx->d_partials + nb, nb, ctx->d_scalars);
    return pano_after_launch(ctx, "neg_divergence(reduce)");
}

int pano_project_launch(pano_ctx *ctx, int dtype, void *vel, const void *p, size_t h, size_t w, double dt) {
    const size_t off = w * (h + 1);
    if ((h + 1) * (w + 1) < ((size_t)1 << 31) && pano_option(ctx, "project_kernel", 0) != 1) {
        const int64_t wide = dtype == PANO_F64 ? pano_option(ctx, "fused_wide", w >= (size_t)kFusedWideMinCols ? 8 : 0) : 0;
        dim3 gm((unsigned)((w + 1 + 31) / 32), (unsigned)((h + 1 + 8 * kProjRows - 1) / (8 * kProjRows)));
        if (wide == 16) {
            dim3 gw((unsigned)((w + 1 + kThreads - 1) / kThreads), (unsigned)((h + 1 + 15) / 16));
            k_project_march<double, 16><<<gw, kThreads, 0, ctx->stream>>>((double *)vel, (double *)vel + off, (const double *)p, (int)h, (int)w, dt, 0, (int)h);
        } else if (wide == 8) {
            dim3 gw((unsigned)((w + 1 + kThreads - 1) / kThreads), (unsigned)((h + 1 + 7) / 8));
            k_project_march<double, 8><<<gw, kThreads, 0, ctx->stream>>>((double *)vel, (double *)vel + off, (const double *)p, (int)h, (int)w, dt, 0, (int)h);
        } else if (dtype == PANO_F64)
            k_project_march<double, 0><<<gm, kThreads, 0, ctx->stream>>>((double *)vel, (double *)vel + off, (const double *)p, (int)h, (int)w, dt, 0, (int)h);
        else
            k_project_march<float, 0><<<gm, kThreads, 0, ctx->stream>>>((float *)vel, (float *)vel + off, (const float *)p, (int)h, (int)w, (float)dt, 0, (int)h);
        return pano_after_launch(ctx, "project_march");
    }
    dim3 g = grid2d((int)h + 1, (int)w + 1);
    if (dtype == PANO_F64)
        k_project<double><<<g, kThreads, 0, ctx->stream>>>((double *)vel, (double *)vel + off, (const double *)p, (int)h, (int)w, dt, 0, (int)h);
    else
        k_project<float><<<g, kThreads, 0, ctx->stream>>>((float *)vel, (float *)vel + off, (const float *)p, (int)h, (int)w, (float)dt, 0, (int)h);
    return pano_after_launch(ctx, "project");
}

// CUDA loads kernels lazily, and loading may need the device to go idle: fatal when another stream is parked in a
// kernel that waits for this very launch.  The multi-GPU step therefore loads everything it will use up front.
int pano_preload_fused() {
    cudaFuncAttributes fa;
    PANO_CUDA(cudaFuncGetAttributes(&fa, k_advect_slab));
    PANO_CUDA(cudaFuncGetAttributes(&fa, k_advect_march3_slab));
    PANO_TRY(pano_preload_advect_tma());
    PANO_CUDA(cudaFuncGetAttributes(&fa, k_neg_divergence<double>));
    PANO_CUDA(cudaFuncGetAttributes(&fa, k_project<double>));
    PANO_CUDA(cudaFuncGetAttributes(&fa, (k_neg_divergence_march<double, 0>)));
    PANO_CUDA(cudaFuncGetAttributes(&fa, (k_project_march<double, 0>)));
    return PANO_OK;
}

// ---- slab forms used by the multi-GPU step (pano_dist.cu): f64, rows [ya, yb) of an h x w grid, every pointer is
// the virtual address of global row 0 of its array
// Stored rows of the sources: [ylo, ylo + rows_q) for q and vx, one more for vy (ylo < 0 on the first rank: storage, not data).
int pano_advect_slab_launch(pano_ctx *ctx, double *q_dst, double *vy_dst, double *vx_dst, const double *q_src, const double *vy_src,
                            const double *vx_src, size_t h, size_t w, double dt, int ya, int yb, int ylo, size_t rows_q, unsigned int *err) {
    const int64_t ak = pano_option(ctx, "advect_kernel", 0);
    if ((ak == 0 || ak == 4) && pano_advect_tma_supported(h, w, ya, ylo, rows_q, q_src + (ptrdiff_t)ylo * (ptrdiff_t)w,
                                                          vy_src + (ptrdiff_t)ylo * (ptrdiff_t)w, vx_src + (ptrdiff_t)ylo * (ptrdiff_t)(w + 1)) &&
        (ak == 4 || (size_t)((yb - ya + 31) / 32) * ((w + 63) / 64) >= (size_t)2 * ctx->num_sms))
        return pano_advect_tma_launch(ctx, q_dst, vy_dst, vx_dst, q_src, vy_src, vx_src, h, w, dt, ya, yb, ylo, rows_q, rows_q + 1, err);
    const int wlo = ylo > 0 ? ylo : 0, whi = ylo + (int)rows_q;
    if (ak != 1 && (h + 1) * (w + 1) < ((size_t)1 << 31)) {
        dim3 gm((unsigned)((w + 1 + 31) / 32), (unsigned)((yb - ya + 1 + 8 * kAdvRows - 1) / (8 * kAdvRows)));
        k_advect_march3_slab<<<gm, kThreads, 0, ctx->stream>>>(q_dst, vy_dst, vx_dst, q_src, vy_src, vx_src, (int)h, (int)w, dt, ya, yb, wlo, whi, err);
        return pano_after_launch(ctx, "advect_march3(slab)");
    }
    dim3 g = grid2d(yb - ya + 1, (int)w + 1);
    k_advect_slab<<<g, kThreads, 0, ctx->stream>>>(q_dst, vy_dst, vx_dst, q_src, vy_src, vx_src, (int)h, (int)w, dt, ya, yb, wlo, whi, err);
    return pano_after_launch(ctx, "advect_slab");
}

int pano_neg_divergence_slab_launch(pano_ctx *ctx, double *b, const double *vy, const double *vx, size_t h, size_t w, pano_rect obstacle,
                                    int ya, int yb) {
    RectI m = pano_clip_rect(obstacle, h + 1, w + 1);
    if ((h + 1) * (w + 1) < ((size_t)1 << 31) && pano_option(ctx, "slab_kernels", 0) == 0) {
        dim3 gm((unsigned)((w + 31) / 32), (unsigned)((yb - ya + 8 * kDivRows - 1) / (8 * kDivRows)));
        k_neg_divergence_march<double, 0><<<gm, kThreads, 0, ctx->stream>>>(b, vy, vx, (int)h, (int)w, m, ya, yb);
        return pano_after_launch(ctx, "neg_divergence_march(slab)");
    }
    dim3 g = grid2d(yb - ya, (int)w);
    k_neg_divergence<double><<<g, kThreads, 0, ctx->stream>>>(b, vy, vx, (int)h, (int)w, m, ya, yb, nullptr, nullptr);
    return pano_after_launch(ctx, "neg_divergence_slab");
}

int pano_project_slab_launch(pano_ctx *ctx, double *vy, double *vx, const double *p, size_t h, size_t w, double dt, int ya, int yb) {
    if ((h + 1) * (w + 1) < ((size_t)1 << 31) && pano_option(ctx, "slab_kernels", 0) == 0) {
        dim3 gm((unsigned)((w + 1 + 31) / 32), (unsigned)((yb - ya + 1 + 8 * kProjRows - 1) / (8 * kProjRows)));
        k_project_march<double, 0><<<gm, kThreads, 0, ctx->stream>>>(vy, vx, p, (int)h, (int)w, dt, ya, yb);
        return pano_after_launch(ctx, "project_march(slab)");
    }
    dim3 g = grid2d(yb - ya + 1, (int)w + 1);
    k_project<double><<<g, kThreads, 0, ctx->stream>>>(vy, vx, p, (int)h, (int)w, dt, ya, yb);
    return pano_after_launch(ctx, "project_slab");
}

extern "C" {

int pano_advect(pano_field *dst, const pano_field *src, double timestep, const pano_field *vel) {
    PANO_TRY(pano_check_kind(dst, PANO_SIMPLEX2, "pano_advect(dst)"));
    PANO_TRY(pano_check_kind(src, PANO_SIMPLEX2, "pano_advect(src)"));
    PANO_TRY(pano_check_kind(vel, PANO_SIMPLEX1, "pano_advect(vel)"));
    PANO_TRY(pano_check_same(dst, src, "pano_advect"));
    PANO_TRY(pano_check_grid(dst, vel, "pano_advect"));
    if (dst->d == src->d) PANO_FAIL(PANO_ERR_INVALID, "pano_advect: dst aliases src (the reference forbids it: &mut vs &)");
    if (dst->h < 2 || dst->w < 2)
        PANO_FAIL(PANO_ERR_SHAPE, "pano_advect: grid %zux%zu; the reference indexes q[iy+1, ix+1] and panics below 2x2", dst->h, dst->w);
    PANO_TRY(pano_activate(dst->ctx));
    return pano_advect_launch(dst->ctx, dst->dtype, dst->d, nullptr, src->d, nullptr, vel->d, dst->h, dst->w, timestep);
}

int pano_advect_mac(pano_field *dst, const pano_field *src, double timestep, const pano_field *vel) {
    PANO_TRY(pano_check_kind(dst, PANO_SIMPLEX1, "pano_advect_mac(dst)"));
    PANO_TRY(pano_check_kind(src, PANO_SIMPLEX1, "pano_advect_mac(src)"));
    PANO_TRY(pano_check_kind(vel, PANO_SIMPLEX1, "pano_advect_mac(vel)"));
    PANO_TRY(pano_check_same(dst, src, "pano_advect_mac"));
    PANO_TRY(pano_check_same(dst, vel, "pano_advect_mac"));
    if (dst->d == src->d || dst->d == vel->d) PANO_FAIL(PANO_ERR_INVALID, "pano_advect_mac: dst aliases an input");
    if (dst->h < 1 || dst->w < 1) PANO_FAIL(PANO_ERR_SHAPE, "pano_advect_mac: empty grid");
    PANO_TRY(pano_activate(dst->ctx));
    return pano_advect_launch(dst->ctx, dst->dtype, nullptr, dst->d, nullptr, src->d, vel->d, dst->h, dst->w, timestep);
}

int pano_advect_all(pano_field *q_dst, pano_field *vel_dst, const pano_field *q_src, const pano_field *vel, double timestep) {
    PANO_TRY(pano_check_kind(q_dst, PANO_SIMPLEX2, "pano_advect_all(q_dst)"));
    PANO_TRY(pano_check_kind(q_src, PANO_SIMPLEX2, "pano_advect_all(q_src)"));
    PANO_TRY(pano_check_kind(vel_dst, PANO_SIMPLEX1, "pano_advect_all(vel_dst)"));
    PANO_TRY(pano_check_kind(vel, PANO_SIMPLEX1, "pano_advect_all(vel)"));
    PANO_TRY(pano_check_same(q_dst, q_src, "pano_advect_all"));
    PANO_TRY(pano_check_same(vel_dst, vel, "pano_advect_all"));
    PANO_TRY(pano_check_grid(q_dst, vel, "pano_advect_all"));
    if (q_dst->d == q_src->d || vel_dst->d == vel->d) PANO_FAIL(PANO_ERR_INVALID, "pano_advect_all: an output aliases its input");
    if (q_dst->h < 2 || q_dst->w < 2) PANO_FAIL(PANO_ERR_SHAPE, "pano_advect_all: grid %zux%zu below 2x2", q_dst->h, q_dst->w);
    PANO_TRY(pano_activate(q_dst->ctx));
    return pano_advect_launch(q_dst->ctx, q_dst->dtype, q_dst->d, vel_dst->d, q_src->d, vel->d, vel->d, q_dst->h, q_dst->w, timestep);
}

int pano_neg_divergence(pano_field *b, const pano_field *vel, pano_rect obstacle, double *rhs_max) {
    PANO_TRY(pano_check_kind(b, PANO_SIMPLEX2, "pano_neg_divergence(b)"));
    PANO_TRY(pano_check_kind(vel, PANO_SIMPLEX1, "pano_neg_divergence(vel)"));
    PANO_TRY(pano_check_grid(b, vel, "pano_neg_divergence"));
    PANO_TRY(pano_check_rect_within(obstacle, b->h, b->w, "pano_neg_divergence(obstacle)"));
    pano_ctx *ctx = b->ctx;
    PANO_TRY(pano_activate(ctx));
    if (b->n == 0) {
        if (rhs_max) *rhs_max = 0.0;
        return PANO_OK;
    }
    PANO_TRY(pano_neg_divergence_launch(ctx, b->dtype, b->d, vel->d, b->h, b->w, obstacle, rhs_max != nullptr));
    if (rhs_max) {
        PANO_CUDA(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        PANO_CUDA(cudaStreamSynchronize(ctx->stream));
        *rhs_max = ctx->h_scalars[0];
    }
    return PANO_OK;
}

int pano_laplacian_apply(pano_field *z, const pano_field *s, double timestep, pano_rect obstacle) {
    PANO_TRY(pano_check_kind(z, PANO_SIMPLEX2, "pano_laplacian_apply(z)"));
    PANO_TRY(pano_check_kind(s, PANO_SIMPLEX2, "pano_laplacian_apply(s)"));
    PANO_TRY(pano_check_same(z, s, "pano_laplacian_apply"));
    if (z->d == s->d) PANO_FAIL(PANO_ERR_INVALID, "pano_laplacian_apply: z aliases s");
    PANO_TRY(pano_check_rect_within(obstacle, z->h, z->w, "pano_laplacian_apply(obstacle)"));
    pano_ctx *ctx = z->ctx;
    PANO_TRY(pano_activate(ctx));
    if (z->n == 0) return PANO_OK;
    RectI m = pano_clip_rect(obstacle, z->h + 1, z->w + 1);
    dim3 g = grid2d((int)z->h, (int)z->w);
    if (z->dtype == PANO_F64)
        k_laplacian<double><<<g, kThreads, 0, ctx->stream>>>((double *)z->d, (const double *)s->d, (int)z->h, (int)z->w, timestep, m);
    else
        k_laplacian<float><<<g, kThreads, 0, ctx->stream>>>((float *)z->d, (const float *)s->d, (int)z->h, (int)z->w, (float)timestep, m);
    return pano_after_launch(ctx, "pano_laplacian_apply");
}

int pano_project(pano_field *vel, const pano_field *pressure, double timestep) {
    PANO_TRY(pano_check_kind(vel, PANO_SIMPLEX1, "pano_project(vel)"));
    PANO_TRY(pano_check_kind(pressure, PANO_SIMPLEX2, "pano_project(pressure)"));
    PANO_TRY(pano_check_grid(vel, pressure, "pano_project"));
    PANO_TRY(pano_activate(vel->ctx));
    if (vel->n == 0) return PANO_OK;
    return pano_project_launch(vel->ctx, vel->dtype, vel->d, pressure->d, vel->h, vel->w, timestep);
}

int pano_density_to_u8(const pano_field *density, double lower, double upper, uint8_t *host_out) {
    PANO_TRY(pano_check_kind(density, PANO_SIMPLEX2, "pano_density_to_u8"));
    if (!host_out) PANO_FAIL(PANO_ERR_INVALID, "pano_density_to_u8: null output");
    pano_ctx *ctx = density->ctx;
    PANO_TRY(pano_activate(ctx));
    if (density->n == 0) return PANO_OK;
    uint8_t *d_out = nullptr;
    PANO_CUDA(cudaMallocAsync((void **)&d_out, density->n, ctx->stream));
    size_t blocks = (density->n + kThreads - 1) / kThreads;
    int g = (int)(blocks < (size_t)ctx->num_sms * 8 ? blocks : (size_t)ctx->num_sms * 8);
    if (density->dtype == PANO_F64)
        k_to_u8<double><<<g, kThreads, 0, ctx->stream>>>(d_out, (const double *)density->d, (int)density->h, (int)density->w, lower, upper);
    else
        k_to_u8<float><<<g, kThreads, 0, ctx->stream>>>(d_out, (const float *)density->d, (int)density->h, (int)density->w, (float)lower, (float)upper);
    PANO_TRY(pano_after_launch(ctx, "pano_density_to_u8"));
    PANO_CUDA(cudaMemcpyAsync(host_out, d_out, density->n, cudaMemcpyDeviceToHost, ctx->stream));
    PANO_CUDA(cudaFreeAsync(d_out, ctx->stream));
    PANO_CUDA(cudaStreamSynchronize(ctx->stream));
    return PANO_OK;
}

}  // extern "C"
