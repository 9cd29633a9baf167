// pano_cell_math.h -- per-cell arithmetic of the grid fluid step, shared by every kernel
// variant (and compilable as host code, so the expressions can be checked against the CPU
// oracle without a GPU: tests/test_cell_math_host.py).
//
// Expression and evaluation ORDER follow the reference exactly so that, built with
// --fmad=false, results are bit-identical to the Rust code (rustc never contracts a*b+c):
//   linear/bilinear          panopaea/src/math/interp.rs:7-20
//   advect                   examples/dec_fluid.rs:173-211
//   advect_mac               examples/dec_fluid.rs:213-291
//   Laplacian closure        examples/dec_fluid.rs:100-119 (dec/grid.rs:223-238, 295-305, 318-334)
//   -divergence              examples/dec_fluid.rs:69-83
//   projection + walls       examples/dec_fluid.rs:124-141
#pragma once

#if defined(__CUDACC__)
#define PANO_HD __host__ __device__ __forceinline__
#else
#define PANO_HD inline
#include <cmath>
#endif

namespace pano {

// no NaNs on this path, so fmin/fmax (one instruction on the device) and the ternary agree bit for bit
PANO_HD double tmin(double a, double b) { return fmin(a, b); }
PANO_HD double tmax(double a, double b) { return fmax(a, b); }
PANO_HD float tmin(float a, float b) { return fminf(a, b); }
PANO_HD float tmax(float a, float b) { return fmaxf(a, b); }
PANO_HD double tfloor(double v) { return floor(v); }
PANO_HD float tfloor(float v) { return floorf(v); }

template <class T>
PANO_HD T linear(T a0, T a1, T s) {
    return a0 * ((T)1 - s) + a1 * s;
}
template <class T>
PANO_HD T bilinear(T a00, T a01, T a10, T a11, T s, T t) {
    return linear(linear(a00, a01, s), linear(a10, a11, s), t);
}

// Rust `f as usize` for f >= 0 (saturating); bounded so that +1 cannot overflow an int.
template <class T>
PANO_HD int to_index(T f) {
    if (!(f > (T)0)) return 0;
    if (f > (T)1.0e9) return 1000000000;
    return (int)f;
}

// ---- advect (cell-centred scalar).  Q, VY, VX are callables (y, x) -> T ------------------
// *_uv forms take the already averaged velocity, so that kernels can share velocity loads between
// the three advected quantities; the wrappers below fix the reference's summation order.
template <class T, class Q>
PANO_HD T advect_cell_uv(int y, int x, int h, int w, T timestep, T ucx, T ucy, const Q &q) {
    const T ndt = -timestep;                                   // integrate_euler(pos, vel, -timestep)
    const T ppx = ((T)x + (T)0.5) + ndt * ucx;
    const T ppy = ((T)y + (T)0.5) + ndt * ucy;
    const T px = tmin(tmax(ppx - (T)0.5, (T)0), (T)w - (T)1.00001);
    const T py = tmin(tmax(ppy - (T)0.5, (T)0), (T)h - (T)1.00001);
    const int ix = (int)tfloor(px), iy = (int)tfloor(py);
    const T u = px - (T)ix, v = py - (T)iy;
    return bilinear(q(iy, ix), q(iy, ix + 1), q(iy + 1, ix), q(iy + 1, ix + 1), u, v);
}
template <class T, class Q, class VY, class VX>
PANO_HD T advect_cell(int y, int x, int h, int w, T timestep, const Q &q, const VY &vy, const VX &vx) {
    const T ucx = (vx(y, x) + vx(y, x + 1)) / (T)2;
    const T ucy = (vy(y, x) + vy(y + 1, x)) / (T)2;
    return advect_cell_uv<T>(y, x, h, w, timestep, ucx, ucy, q);
}

// common tail of both advect_mac loops: index-clamped bilinear gather on an (H, W) array
template <class T, class Q>
PANO_HD T mac_gather(T relx, T rely, int H, int W, const Q &q) {
    const int px = to_index(tmax(tfloor(relx), (T)0));
    const int py = to_index(tmax(tfloor(rely), (T)0));
    const int x0 = px < W - 1 ? px : W - 1, x1 = px + 1 < W - 1 ? px + 1 : W - 1;
    const int y0 = py < H - 1 ? py : H - 1, y1 = py + 1 < H - 1 ? py + 1 : H - 1;
    const T s = tmax(tmin(relx - (T)px, (T)1), (T)0);
    const T t = tmax(tmin(rely - (T)py, (T)1), (T)0);
    return bilinear(q(y0, x0), q(y0, x1), q(y1, x0), q(y1, x1), s, t);
}

// ---- advect_mac, x component: (y, x) in (h, w+1) ------------------------------------------
template <class T, class QX>
PANO_HD T advect_mac_x_uv(int y, int x, int h, int w, T timestep, T vvx, T vvy, const QX &qx) {
    const T ndt = -timestep;
    const T ppx = ((T)x + (T)0.0) + ndt * vvx;
    const T ppy = ((T)y + (T)0.5) + ndt * vvy;
    return mac_gather<T>(ppx - (T)0.0, ppy - (T)0.5, h, w + 1, qx);
}
template <class T, class QX, class VY, class VX>
PANO_HD T advect_mac_x(int y, int x, int h, int w, T timestep, const QX &qx, const VY &vy, const VX &vx) {
    const int xc = x < w - 1 ? x : w - 1, xm = x > 0 ? x - 1 : 0;
    const T vvx = vx(y, x);
    const T vvy = (vy(y, xc) + vy(y + 1, xc) + vy(y, xm) + vy(y + 1, xm)) / (T)4;
    return advect_mac_x_uv<T>(y, x, h, w, timestep, vvx, vvy, qx);
}

// ---- advect_mac, y component: (y, x) in (h+1, w) ------------------------------------------
template <class T, class QY>
PANO_HD T advect_mac_y_uv(int y, int x, int h, int w, T timestep, T vvx, T vvy, const QY &qy) {
    const T ndt = -timestep;
    const T ppx = ((T)x + (T)0.5) + ndt * vvx;
    const T ppy = ((T)y + (T)0.0) + ndt * vvy;
    return mac_gather<T>(ppx - (T)0.5, ppy - (T)0.0, h + 1, w, qy);
}
template <class T, class QY, class VY, class VX>
PANO_HD T advect_mac_y(int y, int x, int h, int w, T timestep, const QY &qy, const VY &vy, const VX &vx) {
    const int yc = y < h - 1 ? y : h - 1, ym = y > 0 ? y - 1 : 0;
    const T vvx = (vx(yc, x) + vx(yc, x + 1) + vx(ym, x) + vx(ym, x + 1)) / (T)4;
    const T vvy = vy(y, x);
    return advect_mac_y_uv<T>(y, x, h, w, timestep, vvx, vvy, qy);
}

// ---- Laplacian closure at one cell.  c = p[y,x]; n/s/w_/e = p at (y-1), (y+1), (x-1), (x+1);
// oN..oE say whether the edge towards that neighbour is open (interior and not masked);
// a closed edge contributes an exact zero, as the zeroed vel_temp entry does in the reference.
//   e.vy[y]   = p[y]   - p[y-1]      e.vy[y+1] = p[y+1] - p[y]
//   e.vx[x]   = p[x-1] - p[x]        e.vx[x+1] = p[x]   - p[x+1]
//   face = -bottom + top - left + right ; result = face * dt
template <class T>
PANO_HD T laplacian_cell(T c, T n, T s, T w_, T e, bool oN, bool oS, bool oW, bool oE, T dt) {
    const T top = oN ? (c - n) : (T)0;
    const T bottom = oS ? (s - c) : (T)0;
    const T left = oW ? (w_ - c) : (T)0;
    const T right = oE ? (c - e) : (T)0;
    return (-bottom + top - left + right) * dt;
}

// ---- -divergence at one cell; arguments are the four face velocities with masked ones
// already replaced by zero.  hodge_1_dual negates vy, so -bottom + top = vy1 + (-vy0).
template <class T>
PANO_HD T neg_divergence_cell(T vy0, T vy1, T vx0, T vx1) {
    const T top = -vy0, bottom = -vy1;
    return -(-bottom + top - vx0 + vx1);
}

}  // namespace pano
