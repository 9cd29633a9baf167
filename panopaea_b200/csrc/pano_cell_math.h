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

// Rust `f as usize` for f >= 0, then min(., lim): the index stays an int, the weight uses the floored REAL itself
// ((T)(f as usize) == f for every f below 2^64; beyond that both corners are the last sample and any weight in [0,1]
// returns it unchanged).
template <class T>
PANO_HD int to_index_min(T f, int lim) {
    return f < (T)lim ? (int)f : lim;
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
    const T fpx = tmax(tfloor(relx), (T)0), fpy = tmax(tfloor(rely), (T)0);     // px, py as REALs
    const int x0 = to_index_min(fpx, W - 1), x1 = x0 + 1 < W - 1 ? x0 + 1 : W - 1;   // min(px, W-1), min(px+1, W-1)
    const int y0 = to_index_min(fpy, H - 1), y1 = y0 + 1 < H - 1 ? y0 + 1 : H - 1;
    const T s = tmax(tmin(relx - fpx, (T)1), (T)0);
    const T t = tmax(tmin(rely - fpy, (T)1), (T)0);
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

// ---- exact fast forms (f64) ----------------------------------------------------------------
// The straightforward forms above cost ~120 issue slots per advected value on sm_100a: there is no
// f64 min/max instruction (every clamp is DSETP + 2 FSEL) and FRND / F2I / I2F on 64-bit types run
// at 1/8 of the FP64 rate.  The forms below give the SAME bits for every input with
// |coordinate| < 2^32 cells (anything else falls back to the general form), using
//   * floor of a non-negative value by one add in round-down mode: t = v (+, RD) 2^52 is
//     2^52 + floor(v) exactly, its low word IS floor(v) as an integer and t - 2^52 is floor(v)
//     as a double (exact); the high word of t equals that of 2^52 iff v < 2^32.
//   * advect_mac's three clamps per axis, max(floor(rel), 0), min(rel - p, 1), max(.., 0)
//     (dec_fluid.rs:232-246), collapse to ONE: with r = max(rel, 0), p = floor(r) and the weight is
//     r - p.  rel >= 0: floor(rel) >= 0 and rel - floor(rel) lies in [0, 1) and is exact, so both
//     weight clamps are no-ops; rel < 0: p = 0 and max(min(rel, 1), 0) = 0 = r - p.
// Only the sign of a zero weight can differ, which no later operation on this path can observe.
struct FloorNN {
    double f;          // floor(v)
    unsigned i;        // floor(v) mod 2^32
    unsigned hi;       // == kFloorHi iff v < 2^32
};
constexpr unsigned kFloorHi = 0x43300000u;
PANO_HD FloorNN floor_nonneg(double v) {   // requires v >= 0
    FloorNN r;
#if defined(__CUDA_ARCH__)
    const double t = __dadd_rd(v, 4503599627370496.0);
    r.i = (unsigned)__double2loint(t);
    r.hi = (unsigned)__double2hiint(t);
#else
    const double t = floor(v) + 4503599627370496.0;   // the same value: this sum is exact below 2^52
    unsigned long long bits;
    __builtin_memcpy(&bits, &t, 8);
    r.i = (unsigned)bits;
    r.hi = (unsigned)(bits >> 32);
#endif
    r.f = t - 4503599627370496.0;
    return r;
}

// max(v, 0) and min(v, lim) without the compiler's NaN-propagating max/min expansion (6 instructions on sm_100a):
// the first is three integer instructions on the sign bit, the second one compare and a 64-bit select.
PANO_HD double clamp_lo0(double v) {          // v <= 0 (or -0.0) -> +0.0
#if defined(__CUDA_ARCH__)
    const int hi = __double2hiint(v), lo = __double2loint(v);
    const int m = ~(hi >> 31);
    return __hiloint2double(hi & m, lo & m);
#else
    return v > 0.0 ? v : 0.0;
#endif
}
PANO_HD double clamp_hi(double v, double lim) {   // v < lim ? v : lim
#if defined(__CUDA_ARCH__)
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(v), "d"(lim));
    return r;
#else
    return v < lim ? v : lim;
#endif
}

// The gather split in two so that a kernel can compute the coordinates of several values first and then issue all of
// their loads together.  `bad` != 0 when a coordinate is beyond 2^32 cells: the caller then uses mac_gather.
struct MacCoord {
    unsigned x0, x1, y0, y1;
    double s, t;
    unsigned bad;
};
PANO_HD MacCoord mac_coord_fast(double relx, double rely, int H, int W) {
    MacCoord c;
    const double rx = clamp_lo0(relx), ry = clamp_lo0(rely);
    const FloorNN fx = floor_nonneg(rx), fy = floor_nonneg(ry);
    c.bad = (fx.hi ^ kFloorHi) | (fy.hi ^ kFloorHi);
    const unsigned wm = (unsigned)(W - 1), hm = (unsigned)(H - 1);
    c.x0 = fx.i < wm ? fx.i : wm; c.x1 = c.x0 + 1u < wm ? c.x0 + 1u : wm;   // = min(p, W-1), min(p+1, W-1)
    c.y0 = fy.i < hm ? fy.i : hm; c.y1 = c.y0 + 1u < hm ? c.y0 + 1u : hm;
    c.s = rx - fx.f; c.t = ry - fy.f;
    return c;
}
template <class Q>
PANO_HD double mac_gather_at(const MacCoord &c, const Q &q) {
    return bilinear(q((int)c.y0, (int)c.x0), q((int)c.y0, (int)c.x1), q((int)c.y1, (int)c.x0), q((int)c.y1, (int)c.x1), c.s, c.t);
}
template <class Q>
PANO_HD bool mac_gather_fast(double relx, double rely, int H, int W, const Q &q, double &out) {
    const MacCoord c = mac_coord_fast(relx, rely, H, W);
    if (c.bad != 0u) return false;
    out = mac_gather_at(c, q);
    return true;
}

// advect's coordinates: always in range (clamped), so no `bad`
struct CellCoord {
    int ix, iy;
    double u, v;
};
PANO_HD CellCoord advect_coord_fast(double xh, double yh, double wlim, double hlim, double ndt, double ucx, double ucy) {
    const double ppx = xh + ndt * ucx, ppy = yh + ndt * ucy;
    const double px = clamp_hi(clamp_lo0(ppx - 0.5), wlim), py = clamp_hi(clamp_lo0(ppy - 0.5), hlim);
    const FloorNN fx = floor_nonneg(px), fy = floor_nonneg(py);      // px, py in [0, 2^31): no guard needed
    CellCoord c;
    c.ix = (int)fx.i; c.iy = (int)fy.i;
    c.u = px - fx.f; c.v = py - fy.f;
    return c;
}
template <class Q>
PANO_HD double advect_gather_at(const CellCoord &c, const Q &q) {
    return bilinear(q(c.iy, c.ix), q(c.iy, c.ix + 1), q(c.iy + 1, c.ix), q(c.iy + 1, c.ix + 1), c.u, c.v);
}

// xh = x + 0.5, yh = y + 0.5 as doubles (the caller carries them instead of converting per cell);
// wlim = w - 1.00001, hlim = h - 1.00001
template <class Q>
PANO_HD double advect_cell_fast(double xh, double yh, double wlim, double hlim, double ndt, double ucx, double ucy, const Q &q) {
    return advect_gather_at(advect_coord_fast(xh, yh, wlim, hlim, ndt, ucx, ucy), q);
}
// advect_mac: the backtraced position relative to the component's own sample grid.
// x component at (x, y + 0.5): xd = (double)x, yh = y + 0.5; `ppx - 0.0` of the reference is the identity.
PANO_HD void mac_x_rel(double xd, double yh, double ndt, double vvx, double vvy, double &relx, double &rely) {
    relx = xd + ndt * vvx;
    rely = (yh + ndt * vvy) - 0.5;
}
// y component at (x + 0.5, y)
PANO_HD void mac_y_rel(double xh, double yd, double ndt, double vvx, double vvy, double &relx, double &rely) {
    relx = (xh + ndt * vvx) - 0.5;
    rely = yd + ndt * vvy;
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

// ================================================================================= Grid3d
// The reference has `trilinear` (panopaea/src/math/interp.rs:23-36) and the bare struct Grid3d (domain/grid.rs:17-20);
// everything else below is dec_fluid.rs carried to (z, y, x) rule by rule -- the 2-D line each expression extends is cited.
template <class T>
PANO_HD T trilinear(T a000, T a001, T a010, T a011, T a100, T a101, T a110, T a111, T s, T t, T u) {
    return linear(bilinear(a000, a001, a010, a011, s, t), bilinear(a100, a101, a110, a111, s, t), u);
}

// ---- advect in 3-D (dec_fluid.rs:173-211 + a z axis): Q(z, y, x) -> T; the averaged velocity is passed in
template <class T, class Q>
PANO_HD T advect3_cell_uv(int z, int y, int x, int d, int h, int w, T timestep, T ucx, T ucy, T ucz, const Q &q) {
    const T ndt = -timestep;
    const T ppx = ((T)x + (T)0.5) + ndt * ucx;
    const T ppy = ((T)y + (T)0.5) + ndt * ucy;
    const T ppz = ((T)z + (T)0.5) + ndt * ucz;
    const T px = tmin(tmax(ppx - (T)0.5, (T)0), (T)w - (T)1.00001);
    const T py = tmin(tmax(ppy - (T)0.5, (T)0), (T)h - (T)1.00001);
    const T pz = tmin(tmax(ppz - (T)0.5, (T)0), (T)d - (T)1.00001);
    const int ix = (int)tfloor(px), iy = (int)tfloor(py), iz = (int)tfloor(pz);
    const T s = px - (T)ix, t = py - (T)iy, u = pz - (T)iz;
    return trilinear(q(iz, iy, ix), q(iz, iy, ix + 1), q(iz, iy + 1, ix), q(iz, iy + 1, ix + 1), q(iz + 1, iy, ix),
                     q(iz + 1, iy, ix + 1), q(iz + 1, iy + 1, ix), q(iz + 1, iy + 1, ix + 1), s, t, u);
}

// one axis of advect_mac's index-clamped gather (dec_fluid.rs:232-246), general form
struct MacAxis {
    int i0, i1;
    double s;
};
PANO_HD MacAxis mac_axis(double rel, int N) {
    const double fp = tmax(tfloor(rel), 0.0);
    MacAxis a;
    a.i0 = to_index_min(fp, N - 1);
    a.i1 = a.i0 + 1 < N - 1 ? a.i0 + 1 : N - 1;
    a.s = tmax(tmin(rel - fp, 1.0), 0.0);
    return a;
}
// the same bits through the exact fast forms above; anything beyond 2^32 cells takes the general form
PANO_HD MacAxis mac_axis_fast(double rel, int N) {
    const double r = clamp_lo0(rel);
    const FloorNN f = floor_nonneg(r);
    if (f.hi != kFloorHi) return mac_axis(rel, N);
    const unsigned nm = (unsigned)(N - 1);
    MacAxis a;
    const unsigned i0 = f.i < nm ? f.i : nm;
    a.i0 = (int)i0;
    a.i1 = (int)(i0 + 1u < nm ? i0 + 1u : nm);
    a.s = r - f.f;
    return a;
}
// index-clamped trilinear gather on a (D, H, W) array at the position (relx, rely, relz) relative to its samples
template <bool kFast, class Q>
PANO_HD double mac3_gather(double relx, double rely, double relz, int D, int H, int W, const Q &q) {
    const MacAxis ax = kFast ? mac_axis_fast(relx, W) : mac_axis(relx, W);
    const MacAxis ay = kFast ? mac_axis_fast(rely, H) : mac_axis(rely, H);
    const MacAxis az = kFast ? mac_axis_fast(relz, D) : mac_axis(relz, D);
    return trilinear(q(az.i0, ay.i0, ax.i0), q(az.i0, ay.i0, ax.i1), q(az.i0, ay.i1, ax.i0), q(az.i0, ay.i1, ax.i1),
                     q(az.i1, ay.i0, ax.i0), q(az.i1, ay.i0, ax.i1), q(az.i1, ay.i1, ax.i0), q(az.i1, ay.i1, ax.i1), ax.s, ay.s, az.s);
}
// one axis of advect's coordinate-clamped gather, fast form: ch = coordinate + 0.5, lim = N - 1.00001
struct CellAxis {
    int i;
    double f;
};
PANO_HD CellAxis advect_axis_fast(double ch, double lim, double ndt, double u) {
    const double pp = ch + ndt * u;
    const double p = clamp_hi(clamp_lo0(pp - 0.5), lim);
    const FloorNN fl = floor_nonneg(p);
    CellAxis a;
    a.i = (int)fl.i;
    a.f = p - fl.f;
    return a;
}
template <class Q>
PANO_HD double advect3_cell_fast(int z, int y, int x, int d, int h, int w, double timestep, double ucx, double ucy, double ucz,
                                 const Q &q) {
    const double ndt = -timestep;
    const CellAxis ax = advect_axis_fast((double)x + 0.5, (double)w - 1.00001, ndt, ucx);
    const CellAxis ay = advect_axis_fast((double)y + 0.5, (double)h - 1.00001, ndt, ucy);
    const CellAxis az = advect_axis_fast((double)z + 0.5, (double)d - 1.00001, ndt, ucz);
    return trilinear(q(az.i, ay.i, ax.i), q(az.i, ay.i, ax.i + 1), q(az.i, ay.i + 1, ax.i), q(az.i, ay.i + 1, ax.i + 1),
                     q(az.i + 1, ay.i, ax.i), q(az.i + 1, ay.i, ax.i + 1), q(az.i + 1, ay.i + 1, ax.i), q(az.i + 1, ay.i + 1, ax.i + 1),
                     ax.f, ay.f, az.f);
}

// ---- the 3-D advection of every quantity stored at one (z, y, x): VZ/VY/VX/Q are callables (z, y, x) -> double.
// Velocity sampling extends dec_fluid.rs:222-228 / :259-265: the component's own value, and the four-sample mean of
// each other component over the two faces either side along its axis and the two cells either side of this face.
template <bool kFast, class Q, class VZ, class VY, class VX>
PANO_HD double advect3_cell(int z, int y, int x, int d, int h, int w, double dt, const Q &q, const VZ &vz, const VY &vy, const VX &vx) {
    const double ucx = (vx(z, y, x) + vx(z, y, x + 1)) / 2.0;       // :182-185
    const double ucy = (vy(z, y, x) + vy(z, y + 1, x)) / 2.0;
    const double ucz = (vz(z, y, x) + vz(z + 1, y, x)) / 2.0;
    if (kFast) return advect3_cell_fast(z, y, x, d, h, w, dt, ucx, ucy, ucz, q);
    return advect3_cell_uv<double>(z, y, x, d, h, w, dt, ucx, ucy, ucz, q);
}
// *_uv forms: the sampled velocity is passed in, so that a kernel can share velocity loads between the four quantities
template <bool kFast, class Q>
PANO_HD double advect3_mac_x_uv(int z, int y, int x, int d, int h, int w, double dt, double vvx, double vvy, double vvz, const Q &qx) {
    const double ndt = -dt;
    const double ppx = ((double)x + 0.0) + ndt * vvx, ppy = ((double)y + 0.5) + ndt * vvy, ppz = ((double)z + 0.5) + ndt * vvz;
    return mac3_gather<kFast>(ppx - 0.0, ppy - 0.5, ppz - 0.5, d, h, w + 1, qx);
}
template <bool kFast, class Q>
PANO_HD double advect3_mac_y_uv(int z, int y, int x, int d, int h, int w, double dt, double vvx, double vvy, double vvz, const Q &qy) {
    const double ndt = -dt;
    const double ppx = ((double)x + 0.5) + ndt * vvx, ppy = ((double)y + 0.0) + ndt * vvy, ppz = ((double)z + 0.5) + ndt * vvz;
    return mac3_gather<kFast>(ppx - 0.5, ppy - 0.0, ppz - 0.5, d, h + 1, w, qy);
}
template <bool kFast, class Q>
PANO_HD double advect3_mac_z_uv(int z, int y, int x, int d, int h, int w, double dt, double vvx, double vvy, double vvz, const Q &qz) {
    const double ndt = -dt;
    const double ppx = ((double)x + 0.5) + ndt * vvx, ppy = ((double)y + 0.5) + ndt * vvy, ppz = ((double)z + 0.0) + ndt * vvz;
    return mac3_gather<kFast>(ppx - 0.5, ppy - 0.5, ppz - 0.0, d + 1, h, w, qz);
}
template <bool kFast, class Q, class VZ, class VY, class VX>
PANO_HD double advect3_mac_x(int z, int y, int x, int d, int h, int w, double dt, const Q &qx, const VZ &vz, const VY &vy, const VX &vx) {
    const int xc = x < w - 1 ? x : w - 1, xm = x > 0 ? x - 1 : 0;   // (z, y, x) in (d, h, w+1), :218-253
    const double vvx = vx(z, y, x);
    const double vvy = (vy(z, y, xc) + vy(z, y + 1, xc) + vy(z, y, xm) + vy(z, y + 1, xm)) / 4.0;
    const double vvz = (vz(z, y, xc) + vz(z + 1, y, xc) + vz(z, y, xm) + vz(z + 1, y, xm)) / 4.0;
    return advect3_mac_x_uv<kFast>(z, y, x, d, h, w, dt, vvx, vvy, vvz, qx);
}
template <bool kFast, class Q, class VZ, class VY, class VX>
PANO_HD double advect3_mac_y(int z, int y, int x, int d, int h, int w, double dt, const Q &qy, const VZ &vz, const VY &vy, const VX &vx) {
    const int yc = y < h - 1 ? y : h - 1, ym = y > 0 ? y - 1 : 0;   // (z, y, x) in (d, h+1, w), :255-290
    const double vvx = (vx(z, yc, x) + vx(z, yc, x + 1) + vx(z, ym, x) + vx(z, ym, x + 1)) / 4.0;
    const double vvy = vy(z, y, x);
    const double vvz = (vz(z, yc, x) + vz(z + 1, yc, x) + vz(z, ym, x) + vz(z + 1, ym, x)) / 4.0;
    return advect3_mac_y_uv<kFast>(z, y, x, d, h, w, dt, vvx, vvy, vvz, qy);
}
template <bool kFast, class Q, class VZ, class VY, class VX>
PANO_HD double advect3_mac_z(int z, int y, int x, int d, int h, int w, double dt, const Q &qz, const VZ &vz, const VY &vy, const VX &vx) {
    const int zc = z < d - 1 ? z : d - 1, zm = z > 0 ? z - 1 : 0;   // (z, y, x) in (d+1, h, w)
    const double vvx = (vx(zc, y, x) + vx(zc, y, x + 1) + vx(zm, y, x) + vx(zm, y, x + 1)) / 4.0;
    const double vvy = (vy(zc, y, x) + vy(zc, y + 1, x) + vy(zm, y, x) + vy(zm, y + 1, x)) / 4.0;
    const double vvz = vz(z, y, x);
    return advect3_mac_z_uv<kFast>(z, y, x, d, h, w, dt, vvx, vvy, vvz, qz);
}

// ---- the 7-point closure at one cell (dec_fluid.rs:100-119 with a z pair in front): f / k = p at (z-1) / (z+1)
//   e.vz[z] = p[z] - p[z-1], e.vz[z+1] = p[z+1] - p[z];  cell = -back + front - bottom + top - left + right;  * dt
template <class T>
PANO_HD T laplacian3_cell(T c, T f, T k, T n, T s, T w_, T e, bool oF, bool oK, bool oN, bool oS, bool oW, bool oE, T dt) {
    const T front = oF ? (c - f) : (T)0;
    const T back = oK ? (k - c) : (T)0;
    const T top = oN ? (c - n) : (T)0;
    const T bottom = oS ? (s - c) : (T)0;
    const T left = oW ? (w_ - c) : (T)0;
    const T right = oE ? (c - e) : (T)0;
    return (-back + front - bottom + top - left + right) * dt;
}
// ---- -divergence (dec_fluid.rs:69-83): the hodge negates vz and vy, masked faces arrive as zero
template <class T>
PANO_HD T neg_divergence3_cell(T vz0, T vz1, T vy0, T vy1, T vx0, T vx1) {
    const T front = -vz0, back = -vz1, top = -vy0, bottom = -vy1;
    return -(-back + front - bottom + top - vx0 + vx1);
}

}  // namespace pano
