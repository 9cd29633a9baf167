// Host build of panopaea_b200/csrc/pano_cell_math.h: lets the CPU test-suite compare the
// exact per-cell expressions the CUDA kernels use against the oracle, without a GPU.
// TEST INFRASTRUCTURE ONLY (built by tests/test_cell_math_host.py with g++ -ffp-contract=off).
#include <cstddef>

#include "../../panopaea_b200/csrc/pano_cell_math.h"

template <class T>
struct HView {
    const T *p;
    int pitch;
    T operator()(int y, int x) const { return p[(size_t)y * pitch + x]; }
};

struct RectI { int y0, y1, x0, x1; };
static inline bool in_rect(const RectI &r, int y, int x) { return y >= r.y0 && y < r.y1 && x >= r.x0 && x < r.x1; }

template <class T>
static void advect_all(int h, int w, T *q_dst, T *vel_dst, const T *q_src, const T *vel, T dt) {
    const size_t off = (size_t)w * (h + 1);
    HView<T> vy{vel, w}, vx{vel + off, w + 1}, q{q_src, w};
    for (int y = 0; y <= h; ++y)
        for (int x = 0; x <= w; ++x) {
            if (y < h && x < w) q_dst[(size_t)y * w + x] = pano::advect_cell<T>(y, x, h, w, dt, q, vy, vx);
            if (y < h) vel_dst[off + (size_t)y * (w + 1) + x] = pano::advect_mac_x<T>(y, x, h, w, dt, vx, vy, vx);
            if (x < w) vel_dst[(size_t)y * w + x] = pano::advect_mac_y<T>(y, x, h, w, dt, vy, vy, vx);
        }
}

// the exact fast forms used by k_advect_march3 (f64), with the kernel's own operand carrying
static void advect_all_fast(int h, int w, double *q_dst, double *vel_dst, const double *q_src, const double *vel, double dt) {
    const size_t off = (size_t)w * (h + 1);
    HView<double> vy{vel, w}, vx{vel + off, w + 1}, q{q_src, w};
    const double ndt = -dt, wlim = (double)w - 1.00001, hlim = (double)h - 1.00001;
    for (int y = 0; y <= h; ++y)
        for (int x = 0; x <= w; ++x) {
            const double xd = (double)x, xh = xd + 0.5, yd = (double)y, yh = yd + 0.5;
            if (y < h && x < w) {
                const double ucx = (vx(y, x) + vx(y, x + 1)) / 2.0, ucy = (vy(y, x) + vy(y + 1, x)) / 2.0;
                q_dst[(size_t)y * w + x] = pano::advect_cell_fast(xh, yh, wlim, hlim, ndt, ucx, ucy, q);
            }
            if (y < h) {
                const int xc = x < w - 1 ? x : w - 1, xm = x > 0 ? x - 1 : 0;
                const double vvy = (vy(y, xc) + vy(y + 1, xc) + vy(y, xm) + vy(y + 1, xm)) / 4.0;
                double rx, ry, v;
                pano::mac_x_rel(xd, yh, ndt, vx(y, x), vvy, rx, ry);
                if (!pano::mac_gather_fast(rx, ry, h, w + 1, vx, v)) v = pano::mac_gather<double>(rx, ry, h, w + 1, vx);
                vel_dst[off + (size_t)y * (w + 1) + x] = v;
            }
            if (x < w) {
                const int yc = y < h - 1 ? y : h - 1, ym = y > 0 ? y - 1 : 0;
                const double vvx = (vx(yc, x) + vx(yc, x + 1) + vx(ym, x) + vx(ym, x + 1)) / 4.0;
                double rx, ry, v;
                pano::mac_y_rel(xh, yd, ndt, vvx, vy(y, x), rx, ry);
                if (!pano::mac_gather_fast(rx, ry, h + 1, w, vy, v)) v = pano::mac_gather<double>(rx, ry, h + 1, w, vy);
                vel_dst[(size_t)y * w + x] = v;
            }
        }
}

template <class T>
static void laplacian(int h, int w, T *z, const T *p, T dt, RectI m) {
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const bool oN = y > 0 && !in_rect(m, y, x), oS = y < h - 1 && !in_rect(m, y + 1, x);
            const bool oW = x > 0 && !in_rect(m, y, x), oE = x < w - 1 && !in_rect(m, y, x + 1);
            const size_t i = (size_t)y * w + x;
            const T c = p[i];
            const T n = oN ? p[i - w] : (T)0, s = oS ? p[i + w] : (T)0;
            const T l = oW ? p[i - 1] : (T)0, r = oE ? p[i + 1] : (T)0;
            z[i] = pano::laplacian_cell<T>(c, n, s, l, r, oN, oS, oW, oE, dt);
        }
}

template <class T>
static void neg_divergence(int h, int w, T *b, const T *vel, RectI m) {
    const T *vy = vel, *vx = vel + (size_t)w * (h + 1);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            T vy0 = in_rect(m, y, x) ? (T)0 : vy[(size_t)y * w + x];
            T vy1 = in_rect(m, y + 1, x) ? (T)0 : vy[(size_t)(y + 1) * w + x];
            T vx0 = in_rect(m, y, x) ? (T)0 : vx[(size_t)y * (w + 1) + x];
            T vx1 = in_rect(m, y, x + 1) ? (T)0 : vx[(size_t)y * (w + 1) + x + 1];
            b[(size_t)y * w + x] = pano::neg_divergence_cell<T>(vy0, vy1, vx0, vx1);
        }
}

extern "C" {
void hc_advect_all_f64(int h, int w, double *qd, double *vd, const double *q, const double *v, double dt) { advect_all<double>(h, w, qd, vd, q, v, dt); }
void hc_advect_all_fast_f64(int h, int w, double *qd, double *vd, const double *q, const double *v, double dt) { advect_all_fast(h, w, qd, vd, q, v, dt); }
void hc_advect_all_f32(int h, int w, float *qd, float *vd, const float *q, const float *v, float dt) { advect_all<float>(h, w, qd, vd, q, v, dt); }
void hc_laplacian_f64(int h, int w, double *z, const double *p, double dt, int y0, int y1, int x0, int x1) { laplacian<double>(h, w, z, p, dt, RectI{y0, y1, x0, x1}); }
void hc_laplacian_f32(int h, int w, float *z, const float *p, float dt, int y0, int y1, int x0, int x1) { laplacian<float>(h, w, z, p, dt, RectI{y0, y1, x0, x1}); }
void hc_neg_divergence_f64(int h, int w, double *b, const double *v, int y0, int y1, int x0, int x1) { neg_divergence<double>(h, w, b, v, RectI{y0, y1, x0, x1}); }
void hc_neg_divergence_f32(int h, int w, float *b, const float *v, int y0, int y1, int x0, int x1) { neg_divergence<float>(h, w, b, v, RectI{y0, y1, x0, x1}); }
}

// ------------------------------------------------------------------ Grid3d forms
struct H3 {
    const double *p;
    int H, W;
    double operator()(int z, int y, int x) const { return p[((size_t)z * H + y) * W + x]; }
};
template <bool kFast>
static void advect3_all(int d, int h, int w, double *qd, double *vd, const double *q, const double *src, const double *vel, double dt) {
    const size_t nz = (size_t)(d + 1) * h * w, ny = (size_t)d * (h + 1) * w;
    const H3 vz{vel, h, w}, vy{vel + nz, h + 1, w}, vx{vel + nz + ny, h, w + 1};
    const H3 qz{src, h, w}, qy{src + nz, h + 1, w}, qx{src + nz + ny, h, w + 1};
    for (int z = 0; z <= d; ++z)
        for (int y = 0; y <= h; ++y)
            for (int x = 0; x <= w; ++x) {
                const bool xin = x < w, yin = y < h, zin = z < d;
                if (xin && yin && zin) qd[((size_t)z * h + y) * w + x] = pano::advect3_cell<kFast>(z, y, x, d, h, w, dt, H3{q, h, w}, vz, vy, vx);
                if (yin && zin) vd[nz + ny + ((size_t)z * h + y) * (w + 1) + x] = pano::advect3_mac_x<kFast>(z, y, x, d, h, w, dt, qx, vz, vy, vx);
                if (xin && zin) vd[nz + ((size_t)z * (h + 1) + y) * w + x] = pano::advect3_mac_y<kFast>(z, y, x, d, h, w, dt, qy, vz, vy, vx);
                if (xin && yin) vd[((size_t)z * h + y) * w + x] = pano::advect3_mac_z<kFast>(z, y, x, d, h, w, dt, qz, vz, vy, vx);
            }
}
struct Box3 {
    int z0, z1, y0, y1, x0, x1;
    bool has(int z, int y, int x) const { return z >= z0 && z < z1 && y >= y0 && y < y1 && x >= x0 && x < x1; }
};
static void laplacian3(int d, int h, int w, double *out, const double *p, double dt, Box3 m) {
    const size_t plane = (size_t)h * w;
    for (int z = 0; z < d; ++z)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const size_t i = z * plane + (size_t)y * w + x;
                const bool here = m.has(z, y, x);
                const bool oF = z > 0 && !here, oK = z < d - 1 && !m.has(z + 1, y, x);
                const bool oN = y > 0 && !here, oS = y < h - 1 && !m.has(z, y + 1, x);
                const bool oW = x > 0 && !here, oE = x < w - 1 && !m.has(z, y, x + 1);
                out[i] = pano::laplacian3_cell<double>(p[i], oF ? p[i - plane] : 0.0, oK ? p[i + plane] : 0.0, oN ? p[i - w] : 0.0,
                                                       oS ? p[i + w] : 0.0, oW ? p[i - 1] : 0.0, oE ? p[i + 1] : 0.0, oF, oK, oN, oS, oW, oE, dt);
            }
}
static void neg_divergence3(int d, int h, int w, double *b, const double *vel, Box3 m) {
    const size_t nz = (size_t)(d + 1) * h * w, ny = (size_t)d * (h + 1) * w;
    const H3 vz{vel, h, w}, vy{vel + nz, h + 1, w}, vx{vel + nz + ny, h, w + 1};
    for (int z = 0; z < d; ++z)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const bool here = m.has(z, y, x);
                b[((size_t)z * h + y) * w + x] = pano::neg_divergence3_cell<double>(
                    here ? 0.0 : vz(z, y, x), m.has(z + 1, y, x) ? 0.0 : vz(z + 1, y, x), here ? 0.0 : vy(z, y, x),
                    m.has(z, y + 1, x) ? 0.0 : vy(z, y + 1, x), here ? 0.0 : vx(z, y, x), m.has(z, y, x + 1) ? 0.0 : vx(z, y, x + 1));
            }
}

extern "C" {
void hc_advect3_all(int fast, int d, int h, int w, double *qd, double *vd, const double *q, const double *src, const double *vel, double dt) {
    if (fast) advect3_all<true>(d, h, w, qd, vd, q, src, vel, dt);
    else advect3_all<false>(d, h, w, qd, vd, q, src, vel, dt);
}
void hc_laplacian3(int d, int h, int w, double *out, const double *p, double dt, int z0, int z1, int y0, int y1, int x0, int x1) {
    laplacian3(d, h, w, out, p, dt, Box3{z0, z1, y0, y1, x0, x1});
}
void hc_neg_divergence3(int d, int h, int w, double *b, const double *v, int z0, int z1, int y0, int y1, int x0, int x1) {
    neg_divergence3(d, h, w, b, v, Box3{z0, z1, y0, y1, x0, x1});
}
}
