// pano_grid3.cu -- Grid3d: the dec_fluid loop body on a (z, y, x) staggered grid.
//
// The reference has the bare struct Grid3d (panopaea/src/domain/grid.rs:17-20) and the unused `trilinear`
// (panopaea/src/math/interp.rs:23-36); DESIGN.md 5c defines the rest as examples/dec_fluid.rs:46-141 with a z axis
// added rule by rule (the test suite holds a plain-C and a numpy statement of it).  Per-cell arithmetic: pano_cell_math.h.
//
// Layout (as the 2-D containers, dec/grid.rs:37-62, 76): CELL3 = (d, h, w) row-major; FACE3 = one flat buffer,
// vz (d+1, h, w) | vy (d, h+1, w) | vx (d, h, w+1).
//
// Kernels (all HBM-bound; algorithmic bytes per cell in f64):
//   k3_advect        q, vz, vy, vx advected in ONE pass                      64 B   (4 reads + 4 writes)
//   k3_neg_div       hodge -> box zero -> d2 -> negate fused, z-marching      32 B
//   k3_cg_tile       the WHOLE pcg.rs:14-82 loop, one persistent cooperative kernel, two phases per iteration as
//                    pano_cg.cu; 64 x 16 x zc tiles whose planes arrive as 3-D TMA boxes through a four-stage mbarrier
//                    ring, s'[z-1], s'[z], s'[z+1] in registers (even widths)                 64 B per iteration
//   k3_cg            the same loop, one column per thread, neighbours through L1 (any shape)
//   k3_project       gradient + axpy + walls fused, z-marching                56 B   (p, 3 faces read; 3 faces written)
#include "pano_cell_math.h"
#include "pano_gridsync.cuh"
#include "pano_internal.cuh"
#include "pano_sm100.cuh"

namespace {

constexpr int kThreads = 256;   // 32 (x) x 8 (y)

struct BoxI {
    int z0, z1, y0, y1, x0, x1;
};
__device__ __forceinline__ bool in_box(const BoxI &b, int z, int y, int x) {
    return z >= b.z0 && z < b.z1 && y >= b.y0 && y < b.y1 && x >= b.x0 && x < b.x1;
}
BoxI clip_box(pano_box r, size_t dmax, size_t hmax, size_t wmax) {
    auto clip = [](int64_t v, size_t hi) -> int { return v < 0 ? 0 : (v > (int64_t)hi ? (int)hi : (int)v); };
    BoxI o{clip(r.z0, dmax), clip(r.z1, dmax), clip(r.y0, hmax), clip(r.y1, hmax), clip(r.x0, wmax), clip(r.x1, wmax)};
    if (o.z1 <= o.z0 || o.y1 <= o.y0 || o.x1 <= o.x0) o = BoxI{0, 0, 0, 0, 0, 0};
    return o;
}

// (z, y, x) -> value on a (., H, W) array (fewer than 2^31 samples: 32-bit index arithmetic)
struct V3 {
    const double *p;
    int H, W;
    __device__ __forceinline__ double operator()(int z, int y, int x) const { return __ldg(p + (unsigned)((z * H + y) * W + x)); }
};

// ------------------------------------------------------------------ fills
__global__ void k3_fill_box(double *a, int H, int W, BoxI b, double v) {
    const int nx = b.x1 - b.x0, ny = b.y1 - b.y0;
    const size_t n = (size_t)(b.z1 - b.z0) * ny * nx;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = b.x0 + (int)(i % nx), y = b.y0 + (int)((i / nx) % ny), z = b.z0 + (int)(i / ((size_t)nx * ny));
        a[((size_t)z * H + y) * W + x] = v;
    }
}

// ------------------------------------------------------------------ advection: every quantity stored at (z, y, x), one pass
// grid: x tiles of 32 over [0, w], y groups of 8 over [0, h], z over [0, d]
template <bool kScalar, bool kMac>
__global__ void __launch_bounds__(kThreads)
k3_advect(double *__restrict__ q_dst, double *__restrict__ vz_dst, double *__restrict__ vy_dst, double *__restrict__ vx_dst,
          const double *__restrict__ q_src, const double *__restrict__ mz_src, const double *__restrict__ my_src,
          const double *__restrict__ mx_src, const double *__restrict__ vz_p, const double *__restrict__ vy_p,
          const double *__restrict__ vx_p, int d, int h, int w, double dt) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), z = blockIdx.z;
    const V3 vz{vz_p, h, w}, vy{vy_p, h + 1, w}, vx{vx_p, h, w + 1};
    // Blocks away from every wall (block-uniform test): all four quantities exist at every point, the clamped neighbour indices of
    // dec_fluid.rs:222-228 are plain x-1 / y-1 / z-1, and the 18 distinct velocity samples the four backtraces use are loaded once
    // (33 loads in the general form).  Same operands in the same order: the same bits.
    const int bx0 = blockIdx.x * 32, by0 = blockIdx.y * 8;
    if (kScalar && kMac && bx0 >= 1 && bx0 + 31 <= w - 2 && by0 >= 1 && by0 + 7 <= h - 2 && z >= 1 && z <= d - 2) {
        const unsigned W = (unsigned)w, H = (unsigned)h, HW = H * W;
        const unsigned ic = ((unsigned)z * H + (unsigned)y) * W + (unsigned)x;                 // cell and vz arrays
        const unsigned iy = ((unsigned)z * (H + 1u) + (unsigned)y) * W + (unsigned)x;          // vy array
        const unsigned ix = ((unsigned)z * H + (unsigned)y) * (W + 1u) + (unsigned)x;          // vx array
        const double a0 = __ldg(vx_p + ix), a1 = __ldg(vx_p + ix + 1), b0 = __ldg(vx_p + ix - (W + 1u)), b1 = __ldg(vx_p + ix - W),
                     c0 = __ldg(vx_p + ix - H * (W + 1u)), c1 = __ldg(vx_p + ix - H * (W + 1u) + 1);
        const double p0 = __ldg(vy_p + iy), p1 = __ldg(vy_p + iy + W), q0 = __ldg(vy_p + iy - 1), q1 = __ldg(vy_p + iy + W - 1),
                     r0 = __ldg(vy_p + iy - (H + 1u) * W), r1 = __ldg(vy_p + iy - (H + 1u) * W + W);
        const double u0 = __ldg(vz_p + ic), u1 = __ldg(vz_p + ic + HW), s0 = __ldg(vz_p + ic - 1), s1 = __ldg(vz_p + ic + HW - 1),
                     t0 = __ldg(vz_p + ic - W), t1 = __ldg(vz_p + ic + HW - W);
        const double sa = a0 + a1, sp = p0 + p1, su = u0 + u1;     // the leading partial sums the reference's left-to-right sums share
        q_dst[ic] = pano::advect3_cell_fast(z, y, x, d, h, w, dt, sa / 2.0, sp / 2.0, su / 2.0, V3{q_src, h, w});
        vx_dst[ix] = pano::advect3_mac_x_uv<true>(z, y, x, d, h, w, dt, a0, (sp + q0 + q1) / 4.0, (su + s0 + s1) / 4.0, V3{mx_src, h, w + 1});
        vy_dst[iy] = pano::advect3_mac_y_uv<true>(z, y, x, d, h, w, dt, (sa + b0 + b1) / 4.0, p0, (su + t0 + t1) / 4.0, V3{my_src, h + 1, w});
        vz_dst[ic] = pano::advect3_mac_z_uv<true>(z, y, x, d, h, w, dt, (sa + c0 + c1) / 4.0, (sp + r0 + r1) / 4.0, u0, V3{mz_src, h, w});
        return;
    }
    if (x > w || y > h) return;
    const bool xin = x < w, yin = y < h, zin = z < d;
    if (kScalar && xin && yin && zin)
        q_dst[((size_t)z * h + y) * w + x] = pano::advect3_cell<true>(z, y, x, d, h, w, dt, V3{q_src, h, w}, vz, vy, vx);
    if (kMac) {
        if (yin && zin)
            vx_dst[((size_t)z * h + y) * (w + 1) + x] = pano::advect3_mac_x<true>(z, y, x, d, h, w, dt, V3{mx_src, h, w + 1}, vz, vy, vx);
        if (xin && zin)
            vy_dst[((size_t)z * (h + 1) + y) * w + x] = pano::advect3_mac_y<true>(z, y, x, d, h, w, dt, V3{my_src, h + 1, w}, vz, vy, vx);
        if (xin && yin)
            vz_dst[((size_t)z * h + y) * w + x] = pano::advect3_mac_z<true>(z, y, x, d, h, w, dt, V3{mz_src, h, w}, vz, vy, vx);
    }
}

// ------------------------------------------------------------------ -divergence: a column of kZ cells per thread, vz carried
constexpr int kDivZ = 8;
__global__ void __launch_bounds__(kThreads)
k3_neg_div(double *__restrict__ b, const double *__restrict__ vz, const double *__restrict__ vy, const double *__restrict__ vx, int d,
           int h, int w, BoxI m) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), zs = blockIdx.z * kDivZ;
    if (x >= w || y >= h) return;
    const size_t plane = (size_t)h * w;
    double vz0 = in_box(m, zs, y, x) ? 0.0 : vz[zs * plane + (size_t)y * w + x];
#pragma unroll
    for (int k = 0; k < kDivZ; ++k) {
        const int z = zs + k;
        if (z >= d) break;
        const bool here = in_box(m, z, y, x);
        const double vz1 = in_box(m, z + 1, y, x) ? 0.0 : vz[(z + 1) * plane + (size_t)y * w + x];
        const double vy0 = here ? 0.0 : vy[((size_t)z * (h + 1) + y) * w + x];
        const double vy1 = in_box(m, z, y + 1, x) ? 0.0 : vy[((size_t)z * (h + 1) + y + 1) * w + x];
        const double vx0 = here ? 0.0 : vx[((size_t)z * h + y) * (w + 1) + x];
        const double vx1 = in_box(m, z, y, x + 1) ? 0.0 : vx[((size_t)z * h + y) * (w + 1) + x + 1];
        b[z * plane + (size_t)y * w + x] = pano::neg_divergence3_cell<double>(vz0, vz1, vy0, vy1, vx0, vx1);
        vz0 = vz1;
    }
}

// ------------------------------------------------------------------ stand-alone 7-point closure
__global__ void __launch_bounds__(kThreads)
k3_laplacian(double *__restrict__ out, const double *__restrict__ p, int d, int h, int w, double dt, BoxI m) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), z = blockIdx.z;
    if (x >= w || y >= h) return;
    const size_t plane = (size_t)h * w, i = z * plane + (size_t)y * w + x;
    const bool here = in_box(m, z, y, x);
    const bool oF = z > 0 && !here, oK = z < d - 1 && !in_box(m, z + 1, y, x);
    const bool oN = y > 0 && !here, oS = y < h - 1 && !in_box(m, z, y + 1, x);
    const bool oW = x > 0 && !here, oE = x < w - 1 && !in_box(m, z, y, x + 1);
    const double c = p[i];
    const double f = oF ? p[i - plane] : 0.0, k = oK ? p[i + plane] : 0.0;
    const double n = oN ? p[i - w] : 0.0, s = oS ? p[i + w] : 0.0;
    const double l = oW ? p[i - 1] : 0.0, e = oE ? p[i + 1] : 0.0;
    out[i] = pano::laplacian3_cell<double>(c, f, k, n, s, l, e, oF, oK, oN, oS, oW, oE, dt);
}

// ------------------------------------------------------------------ projection + walls: z-marching, p[z-1] carried
constexpr int kProjZ = 8;
__global__ void __launch_bounds__(kThreads)
k3_project(double *__restrict__ vz, double *__restrict__ vy, double *__restrict__ vx, const double *__restrict__ p, int d, int h, int w,
           double dt) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), zs = blockIdx.z * kProjZ;
    if (x > w || y > h) return;
    const bool xin = x < w, yin = y < h, cell = xin && yin;
    const size_t plane = (size_t)h * w;
    double pf = (cell && zs > 0) ? p[(zs - 1) * plane + (size_t)y * w + x] : 0.0;   // p[z-1, y, x]
#pragma unroll
    for (int k = 0; k < kProjZ; ++k) {
        const int z = zs + k;
        if (z > d) break;
        const bool zin = z < d;
        const double c = (cell && zin) ? p[z * plane + (size_t)y * w + x] : 0.0;
        if (cell) {   // vz[z, y, x]
            const size_t i = z * plane + (size_t)y * w + x;
            if (z == 0 || z == d) vz[i] = 0.0;
            else vz[i] = vz[i] + dt * (pf - c);
        }
        if (xin && zin) {   // vy[z, y, x], y in [0, h]
            const size_t i = ((size_t)z * (h + 1) + y) * w + x;
            if (y == 0 || y == h) vy[i] = 0.0;
            else vy[i] = vy[i] + dt * (p[z * plane + (size_t)(y - 1) * w + x] - c);
        }
        if (yin && zin) {   // vx[z, y, x], x in [0, w]
            const size_t i = ((size_t)z * h + y) * (w + 1) + x;
            if (x == 0 || x == w) vx[i] = 0.0;
            else vx[i] = vx[i] + dt * (p[z * plane + (size_t)y * w + x - 1] - c);
        }
        pf = c;
    }
}

// ------------------------------------------------------------------ the solve (pcg.rs:14-82, 7-point closure)
// Phases exactly as pano_cg.cu: P1 s' = r + beta s (folded search update; recomputed on the six neighbours from their OLD values,
// hence s double-buffered), z = A s' evaluated and reduced against s' but never stored; P2 z recomputed from s', x += alpha s',
// r -= alpha z, r.r and max|r| reduced.  64 B per cell and iteration.  A tile is 32 x 8 columns of `zc` cells; a thread walks its
// column with s'[z-1], s'[z], s'[z+1] in registers, so along z every value is formed once per tile.
struct Cg3Args {
    double *x;
    const double *b;
    double *r, *s0, *s1;
    int d, h, w;
    double dt, threshold;
    int max_iter;
    BoxI m;
    double *partials;   // 5 * G
    PanoCgControl *ctl;
    int tiles_x, tiles_y, tiles_z, zc;
};

template <bool kCg>
__device__ __forceinline__ double ld3(const double *p) {
    if (kCg) return __ldcg(p);
    return *p;
}

// kCg: every field load through L2 only (ld.global.cg); otherwise (default) plain loads, whose L1 lines the grid barrier's
// gpu-scope fence invalidates (the contract of cooperative_groups::grid_group::sync()) -- the four lateral neighbours then hit L1.
template <bool kCg>
__global__ void __launch_bounds__(kThreads, 4) k3_cg(Cg3Args a) {
    __shared__ double scratch[32];
    __shared__ int s_flag;
    const int G = gridDim.x;
    const int d = a.d, h = a.h, w = a.w;
    const size_t plane = (size_t)h * w;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int txy = a.tiles_x * a.tiles_y, ntiles = txy * a.tiles_z;
    double *pA = a.partials, *pB = pA + G, *pC = pA + 2 * G, *pD = pA + 3 * G, *pE = pA + 4 * G;
    unsigned long long nbar = 0;
    double sigma = 0, alpha = 0, beta = 0, rmax = 0, bmax = 0;
    int it = 0, applies = 0;
    bool converged = false;
    double *s_cur = a.s0, *s_old = a.s1;

    for (it = 0; it < a.max_iter; ++it) {
        const bool first = it == 0;
        const double *r_src = first ? a.b : a.r;
        // ---------------------------------------------------------------- P1
        double acc_zs = 0, acc_bb = 0, acc_bmax = 0;
        for (int t = blockIdx.x; t < ntiles; t += G) {
            const int x = (t % a.tiles_x) * 32 + lx, y = ((t / a.tiles_x) % a.tiles_y) * 8 + ly, z0 = (t / txy) * a.zc;
            if (x >= w || y >= h) continue;
            const int z1 = z0 + a.zc < d ? z0 + a.zc : d;
            size_t i = z0 * plane + (size_t)y * w + x;
            auto sval = [&](size_t j) -> double { return first ? ld3<kCg>(r_src + j) : ld3<kCg>(r_src + j) + beta * ld3<kCg>(s_old + j); };
            double sm1 = z0 > 0 ? sval(i - plane) : 0.0, sc = sval(i);
#pragma unroll 2
            for (int z = z0; z < z1; ++z, i += plane) {
                const double sp1 = z + 1 < d ? sval(i + plane) : 0.0;
                const bool here = in_box(a.m, z, y, x);
                const bool oF = z > 0 && !here, oK = z < d - 1 && !in_box(a.m, z + 1, y, x);
                const bool oN = y > 0 && !here, oS = y < h - 1 && !in_box(a.m, z, y + 1, x);
                const bool oW = x > 0 && !here, oE = x < w - 1 && !in_box(a.m, z, y, x + 1);
                const double n = oN ? sval(i - w) : 0.0, s = oS ? sval(i + w) : 0.0;
                const double l = oW ? sval(i - 1) : 0.0, e = oE ? sval(i + 1) : 0.0;
                if (first) {
                    const double ab = sc < 0 ? -sc : sc;
                    acc_bmax = ab > acc_bmax ? ab : acc_bmax;
                    acc_bb = acc_bb + sc * sc;
                } else {
                    s_cur[i] = sc;
                }
                const double zv = pano::laplacian3_cell<double>(sc, sm1, sp1, n, s, l, e, oF, oK, oN, oS, oW, oE, a.dt);
                acc_zs = acc_zs + zv * sc;
                sm1 = sc;
                sc = sp1;
            }
        }
        {
            const double v = block_sum(acc_zs, scratch);
            if (threadIdx.x == 0) pA[blockIdx.x] = v;
            if (first) {
                const double v2 = block_sum(acc_bb, scratch), v3 = block_max(acc_bmax, scratch);
                if (threadIdx.x == 0) {
                    pB[blockIdx.x] = v2;
                    pC[blockIdx.x] = v3;
                }
            }
        }
        if (!grid_barrier(a.ctl, (++nbar) * (unsigned long long)G, &s_flag)) return;
        const double zs = sum_partials<double>(pA, G, scratch);
        if (first) {
            sigma = sum_partials<double>(pB, G, scratch);      // pcg.rs:46
            bmax = max_partials<double>(pC, G, scratch);       // pcg.rs:35
            rmax = bmax;
            if (bmax < a.threshold) {                          // early out: x = 0, scratch untouched
                const size_t n = plane * d;
                for (size_t j = blockIdx.x * (size_t)kThreads + threadIdx.x; j < n; j += (size_t)G * kThreads) a.x[j] = 0.0;
                if (blockIdx.x == 0 && threadIdx.x == 0) {
                    a.ctl->iterations = -1;
                    a.ctl->applies = 0;
                    a.ctl->final_residual = bmax;
                    a.ctl->rhs_max = bmax;
                }
                return;
            }
        }
        ++applies;
        alpha = sigma / zs;                                    // pcg.rs:53
        const double nalpha = -alpha;
        // ---------------------------------------------------------------- P2
        double acc_rr = 0, acc_rmax = 0;
        const double *s_rd = first ? a.b : s_cur;
        for (int t = blockIdx.x; t < ntiles; t += G) {
            const int x = (t % a.tiles_x) * 32 + lx, y = ((t / a.tiles_x) % a.tiles_y) * 8 + ly, z0 = (t / txy) * a.zc;
            if (x >= w || y >= h) continue;
            const int z1 = z0 + a.zc < d ? z0 + a.zc : d;
            size_t i = z0 * plane + (size_t)y * w + x;
            double sm1 = z0 > 0 ? ld3<kCg>(s_rd + i - plane) : 0.0, sc = ld3<kCg>(s_rd + i);
#pragma unroll 2
            for (int z = z0; z < z1; ++z, i += plane) {
                const double sp1 = z + 1 < d ? ld3<kCg>(s_rd + i + plane) : 0.0;
                const bool here = in_box(a.m, z, y, x);
                const bool oF = z > 0 && !here, oK = z < d - 1 && !in_box(a.m, z + 1, y, x);
                const bool oN = y > 0 && !here, oS = y < h - 1 && !in_box(a.m, z, y + 1, x);
                const bool oW = x > 0 && !here, oE = x < w - 1 && !in_box(a.m, z, y, x + 1);
                const double n = oN ? ld3<kCg>(s_rd + i - w) : 0.0, s = oS ? ld3<kCg>(s_rd + i + w) : 0.0;
                const double l = oW ? ld3<kCg>(s_rd + i - 1) : 0.0, e = oE ? ld3<kCg>(s_rd + i + 1) : 0.0;
                const double zv = pano::laplacian3_cell<double>(sc, sm1, sp1, n, s, l, e, oF, oK, oN, oS, oW, oE, a.dt);
                if (first) a.s0[i] = sc;                       // pcg.rs:40-42: s = aux = r = b
                const double xo = first ? 0.0 : ld3<kCg>(a.x + i);
                a.x[i] = xo + alpha * sc;                      // pcg.rs:55
                const double rn = ld3<kCg>(r_src + i) + nalpha * zv;   // pcg.rs:56
                a.r[i] = rn;
                const double ar = rn < 0 ? -rn : rn;
                acc_rmax = ar > acc_rmax ? ar : acc_rmax;
                acc_rr = acc_rr + rn * rn;
                sm1 = sc;
                sc = sp1;
            }
        }
        {
            const double v = block_sum(acc_rr, scratch), v2 = block_max(acc_rmax, scratch);
            if (threadIdx.x == 0) {
                pD[blockIdx.x] = v;
                pE[blockIdx.x] = v2;
            }
        }
        if (!grid_barrier(a.ctl, (++nbar) * (unsigned long long)G, &s_flag)) return;
        const double rr = sum_partials<double>(pD, G, scratch);
        rmax = max_partials<double>(pE, G, scratch);           // pcg.rs:58
        if (rmax < a.threshold) {                              // pcg.rs:60-63
            converged = true;
            break;
        }
        beta = rr / sigma;                                     // pcg.rs:67-68
        sigma = rr;                                            // pcg.rs:79
        double *tmp = s_cur;
        s_cur = s_old;
        s_old = tmp;
    }
    // as pano_cg.cu: `search` must end up holding what the reference leaves in it (pcg.rs:72-77 runs once more when the loop runs out)
    const double *s_fin = converged ? s_cur : s_old;
    if (a.max_iter > 0 && (!converged || s_fin != a.s0)) {
        const size_t n = plane * d;
        for (size_t j = blockIdx.x * (size_t)kThreads + threadIdx.x; j < n; j += (size_t)G * kThreads) {
            const double sv = ld3<kCg>(s_fin + j);
            a.s0[j] = converged ? sv : ld3<kCg>(a.r + j) + beta * sv;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.ctl->iterations = converged ? it : a.max_iter;
        a.ctl->applies = applies;
        a.ctl->final_residual = rmax;
        a.ctl->rhs_max = bmax;
    }
}

// ------------------------------------------------------------------ the solve, second generation (even widths): plane tiles
// streamed through shared memory by TMA.  A CTA (512 threads) owns a 64 (x) x 16 (y) column bundle of `zc` planes at a time; every
// thread owns 2 adjacent cells of one row and walks along z with s'[z-1], s'[z], s'[z+1] of its two cells in registers.
//   * Planes arrive through a FOUR-STAGE ring of TMA boxes (cp.async.bulk.tensor.3d, one elected thread, mbarrier completion):
//     an 18 x 68 box brings a plane of the tile WITH its one-cell ring (and two pad columns that keep the box start 16-byte
//     aligned; outside the grid TMA fills zeros), 16 x 64 boxes bring x and r.  Plane z+3 is requested while plane z is
//     evaluated, no register holds data in flight, one __syncthreads per plane.
//   * P1 turns the raw r, s of plane z+1 into s' = r + beta s IN PLACE (own pair + one ring cell for the first 160 threads, the
//     owner's expression: the same bits) one step before the plane is evaluated, so the four lateral neighbours are plain
//     shared-memory reads (or the thread's own registers); P2 stages s' (with ring), x and r.
//   * Tiles are claimed from a counter (first tile static) in x-fastest order, so neighbouring columns of a slab run at the same
//     time and their rings meet in L2; every tile leaves its own partial sums, and after the grid barrier every CTA adds the
//     per-TILE partials in tile order: the scalars do not depend on which CTA ran which tile -- deterministic.
//   * Tiles that no wall and no obstacle face touches skip the open/closed flags, tiles at a wall only compare coordinates.
//   Measured and rejected (B200, 256^3 / 512^3, the cp.async form of this kernel 12.9 / 129 ms per solve): static equal shares of a
//   plane stream whose ring never drains across tile changes -- 15.3 / 154 ms (the per-slot role dispatch costs more than the
//   pipeline fills it removes); the same with column-major shares 17.1 / 174 ms (ring cells and z-neighbour planes then miss L2);
//   a tile's last two request slots fetching the first two planes of the CTA's next tile (12.7 / 126 ms: the second CTA of the SM
//   already covers a tile's pipeline fill).
using pano_sm100::mbar_arrive;
using pano_sm100::mbar_arrive_expect_tx;
using pano_sm100::mbar_init;
using pano_sm100::mbar_wait;
using pano_sm100::smem_u32;

constexpr int kTileThreads = 512;
constexpr int kTX = 64, kTY = 16;
constexpr int kSW = kTX + 4;                  // row stride = box width: interior from column 2, ring in columns 1 and kTX + 2
constexpr int kRingBoxBytes = (kTY + 2) * kSW * 8;        // 9792: what one ringed box transfers
constexpr int kSPlane = 1232;                 // doubles reserved for a ringed plane (9856 B, a multiple of 128)
constexpr int kIntBoxBytes = kTX * kTY * 8;   // 8192
constexpr int kStages = 4;
constexpr int kStageDoubles = kSPlane + 2 * kTX * kTY;   // P2: s' (ring) | x | r;  P1: r (ring) | s (ring) = 2 kSPlane, smaller
constexpr size_t kTileSmemBytes = (size_t)kStages * kStageDoubles * sizeof(double) + 128;   // + slack to align the ring to 128 B
static_assert(2 * kSPlane <= kStageDoubles && (kSPlane * 8) % 128 == 0 && (kStageDoubles * 8) % 128 == 0, "stage layout");
static_assert((kTY + 2) * kSW <= kSPlane, "ringed plane must fit");

struct Cg3TileArgs {
    Cg3Args c;
    unsigned long long *claim;   // zero at launch
    // (w, h, d) tensors: ringed 68 x 18 x 1 boxes of b, r, s0, s1; interior 64 x 16 x 1 boxes of x, r, b
    CUtensorMap mr_b, mr_r, mr_s0, mr_s1, mi_x, mi_r, mi_b;
};

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct TileGeo {
    int x0, y0, z0, z1;
    int mode;   // 0: every face of every cell is open; 1: walls only; 2: obstacle faces too
};
__device__ __forceinline__ TileGeo tile_geo(const Cg3Args &a, int t) {
    TileGeo g;
    const int txy = a.tiles_x * a.tiles_y;
    g.x0 = (t % a.tiles_x) * kTX;
    g.y0 = ((t / a.tiles_x) % a.tiles_y) * kTY;
    g.z0 = (t / txy) * a.zc;
    g.z1 = g.z0 + a.zc < a.d ? g.z0 + a.zc : a.d;
    const bool walls = g.x0 == 0 || g.x0 + kTX >= a.w || g.y0 == 0 || g.y0 + kTY >= a.h || g.z0 == 0 || g.z1 >= a.d;
    // a cell is touched by the obstacle if one of its six faces has its index in the box: cells [b0 - 1, b1) per axis
    const bool box = a.m.z1 > a.m.z0 && g.z0 < a.m.z1 && g.z1 > a.m.z0 - 1 && g.y0 < a.m.y1 && g.y0 + kTY > a.m.y0 - 1 && g.x0 < a.m.x1 &&
                     g.x0 + kTX > a.m.x0 - 1;
    g.mode = box ? 2 : walls ? 1 : 0;
    return g;
}

struct Own3 {        // this thread's share of a plane tile (offsets in elements; every array holds fewer than 2^31)
    int xa, ya;      // cells (ya, xa) and (ya, xa + 1)
    bool v;          // inside the grid
    bool ring;       // this thread also transforms one ring cell
    int so;          // shared-memory offset of (ya, xa) in a ringed plane
    int xo;          // ... in an interior-only 16 x 64 plane
    int rso;         // shared-memory offset of the ring cell
    unsigned j;      // offset of (ya, xa) inside a plane of the grid
};
__device__ __forceinline__ Own3 own_of(const Cg3Args &a, const TileGeo &g) {
    Own3 o;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    o.xa = g.x0 + 2 * tx;
    o.ya = g.y0 + ty;
    o.v = o.xa < a.w && o.ya < a.h;
    o.so = (ty + 1) * kSW + 2 + 2 * tx;
    o.xo = ty * kTX + 2 * tx;
    o.j = o.v ? (unsigned)o.ya * (unsigned)a.w + (unsigned)o.xa : 0u;
    o.ring = tid < 2 * kTX + 2 * kTY;
    o.rso = 0;
    if (tid < kTX) o.rso = 2 + tid;                                                     // row y0 - 1
    else if (tid < 2 * kTX) o.rso = (kTY + 1) * kSW + 2 + tid - kTX;                    // row y0 + 16
    else if (tid < 2 * kTX + kTY) o.rso = (tid - 2 * kTX + 1) * kSW + 1;                // column x0 - 1
    else if (o.ring) o.rso = (tid - 2 * kTX - kTY + 1) * kSW + kTX + 2;                 // column x0 + 64
    return o;
}

__device__ __forceinline__ double2 ldv2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void stv2(double *p, double x, double y) { *reinterpret_cast<double2 *>(p) = make_double2(x, y); }

// z = A s' on the two cells of plane z; c = s'[z], f = s'[z-1], k = s'[z+1] in registers, the plane with its ring in `buf`
template <int kMode>
__device__ __forceinline__ double2 lap_pair(const Cg3Args &a, const Own3 &o, const double *buf, int z, double2 c, double2 f, double2 k) {
    const double2 n = ldv2(buf + o.so - kSW), s = ldv2(buf + o.so + kSW);
    const double l = buf[o.so - 1], e = buf[o.so + 2];
    double2 out;
    if (kMode == 0) {
        out.x = pano::laplacian3_cell<double>(c.x, f.x, k.x, n.x, s.x, l, c.y, true, true, true, true, true, true, a.dt);
        out.y = pano::laplacian3_cell<double>(c.y, f.y, k.y, n.y, s.y, c.x, e, true, true, true, true, true, true, a.dt);
    } else {
        const int y = o.ya, x = o.xa;
        const bool zlo = z > 0, zhi = z < a.d - 1, ylo = y > 0, yhi = y < a.h - 1;
        bool oF0 = zlo, oK0 = zhi, oN0 = ylo, oS0 = yhi, oW0 = x > 0, oE0 = true;            // x + 1 <= w - 1 (even width)
        bool oF1 = zlo, oK1 = zhi, oN1 = ylo, oS1 = yhi, oW1 = true, oE1 = x + 1 < a.w - 1;
        if (kMode == 2) {
            const bool h0 = in_box(a.m, z, y, x), h1 = in_box(a.m, z, y, x + 1);
            oF0 = oF0 && !h0; oN0 = oN0 && !h0; oW0 = oW0 && !h0;
            oF1 = oF1 && !h1; oN1 = oN1 && !h1; oW1 = oW1 && !h1;
            oK0 = oK0 && !in_box(a.m, z + 1, y, x); oS0 = oS0 && !in_box(a.m, z, y + 1, x); oE0 = oE0 && !h1;
            oK1 = oK1 && !in_box(a.m, z + 1, y, x + 1); oS1 = oS1 && !in_box(a.m, z, y + 1, x + 1); oE1 = oE1 && !in_box(a.m, z, y, x + 2);
        }
        out.x = pano::laplacian3_cell<double>(c.x, f.x, k.x, n.x, s.x, l, c.y, oF0, oK0, oN0, oS0, oW0, oE0, a.dt);
        out.y = pano::laplacian3_cell<double>(c.y, f.y, k.y, n.y, s.y, c.x, e, oF1, oK1, oN1, oS1, oW1, oE1, a.dt);
    }
    return out;
}

// The stage ring of one CTA: request number L (counted over the CTA's whole life) uses stage L & 3, and its mbarrier completes
// phase (L >> 2) & 1.  Only planes inside the grid are requested, and every requested plane is waited for by every thread before
// its stage comes round again, so no phase of a barrier ever completes unobserved.
struct Ring3 {
    double *smem;
    uint64_t *full;
    volatile unsigned int *err;
    unsigned L;          // the next slot to be requested
};
__device__ __forceinline__ double *stage_of(const Ring3 &rg, unsigned slot) { return rg.smem + (slot & 3u) * kStageDoubles; }
__device__ __forceinline__ bool wait_slot(const Ring3 &rg, unsigned slot) { return mbar_wait(&rg.full[slot & 3u], (slot >> 2) & 1u, rg.err); }

// thread 0: request plane p (inside the grid) of the tile for P1 into the next slot
__device__ __forceinline__ void request_p1(Ring3 &rg, const TileGeo &g, int p, const CUtensorMap *m_r, const CUtensorMap *m_s) {
    if (threadIdx.x == 0) {
        uint64_t *bar = &rg.full[rg.L & 3u];
        double *st = stage_of(rg, rg.L);
        mbar_arrive_expect_tx(bar, m_s ? 2 * kRingBoxBytes : kRingBoxBytes);
        tma_load_3d(st, m_r, bar, g.x0 - 2, g.y0 - 1, p);
        if (m_s) tma_load_3d(st + kSPlane, m_s, bar, g.x0 - 2, g.y0 - 1, p);
    }
    ++rg.L;
}
// thread 0: request plane p for P2 (s' with ring; x and r of the planes the tile evaluates)
__device__ __forceinline__ void request_p2(Ring3 &rg, const TileGeo &g, int p, const CUtensorMap *m_s, const CUtensorMap *m_x, const CUtensorMap *m_r) {
    if (threadIdx.x == 0) {
        uint64_t *bar = &rg.full[rg.L & 3u];
        double *st = stage_of(rg, rg.L);
        const bool xr = p >= g.z0 && p < g.z1;
        mbar_arrive_expect_tx(bar, kRingBoxBytes + (xr ? (m_x ? 2 : 1) * kIntBoxBytes : 0));
        tma_load_3d(st, m_s, bar, g.x0 - 2, g.y0 - 1, p);
        if (xr) {
            if (m_x) tma_load_3d(st + kSPlane, m_x, bar, g.x0, g.y0, p);
            tma_load_3d(st + kSPlane + kTX * kTY, m_r, bar, g.x0, g.y0, p);
        }
    }
    ++rg.L;
}
// s' = r + beta s (pcg.rs:75-77) of the own pair and the ring cell of a landed plane, in place; returns the own pair
__device__ __forceinline__ double2 transform_p1(const Own3 &o, double *st, bool ws, double beta) {
    double2 rv = ldv2(st + o.so);
    if (ws) {
        const double2 sv = ldv2(st + kSPlane + o.so);
        rv = make_double2(rv.x + beta * sv.x, rv.y + beta * sv.y);
        stv2(st + o.so, rv.x, rv.y);
        if (o.ring) st[o.rso] = st[o.rso] + beta * st[kSPlane + o.rso];
        fence_async_smem();                            // generic writes now, TMA writes to the same stage four slots later
    }
    return rv;
}

// P1 over one tile (m_s: s' = r + beta s_old, stored to s_cur; null: the opening pass on b, max|b| and b.b too).  false: a wait expired
template <int kMode>
__device__ __forceinline__ bool p1_tile(const Cg3Args &a, const TileGeo &g, Ring3 &rg, const CUtensorMap *m_r, const CUtensorMap *m_s, double *s_cur,
                                        double beta, double &zs, double &bb, double &bmax) {
    const Own3 o = own_of(a, g);
    const bool ws = m_s != nullptr;
    const unsigned plane = (unsigned)(a.h * a.w);
    const int plast = g.z1 < a.d ? g.z1 : a.d - 1, pfirst = g.z0 > 0 ? g.z0 - 1 : 0;
    const unsigned L0 = rg.L - (unsigned)pfirst;       // plane p of the tile lives in slot L0 + p
    for (int pq = pfirst; pq <= g.z0 + 2 && pq <= plast; ++pq) request_p1(rg, g, pq, m_r, m_s);
    double2 f = make_double2(0.0, 0.0);
    if (g.z0 > 0) {
        if (!wait_slot(rg, L0 + g.z0 - 1)) return false;
        const double *st = stage_of(rg, L0 + g.z0 - 1);   // plane z0 - 1: only this thread's pair is ever needed, nothing is written back
        f = ldv2(st + o.so);
        if (ws) {
            const double2 sv = ldv2(st + kSPlane + o.so);
            f = make_double2(f.x + beta * sv.x, f.y + beta * sv.y);
        }
    }
    if (!wait_slot(rg, L0 + g.z0)) return false;
    double2 c = transform_p1(o, stage_of(rg, L0 + g.z0), ws, beta);
    unsigned pz = (unsigned)g.z0 * plane;
    for (int z = g.z0; z < g.z1; ++z, pz += plane) {
        const unsigned q = L0 + (unsigned)z;           // slot of plane z
        const bool more = z + 1 < a.d;
        if (more && !wait_slot(rg, q + 1)) return false;    // plane z + 1 has landed
        __syncthreads();                               // plane z is transformed by everybody; plane z - 1 is no longer read
        if (z + 3 <= plast) request_p1(rg, g, z + 3, m_r, m_s);   // slot q + 3 = the stage of plane z - 1
        double2 k = make_double2(0.0, 0.0);
        if (more) k = transform_p1(o, stage_of(rg, q + 1), ws, beta);
        const double2 zv = lap_pair<kMode>(a, o, stage_of(rg, q), z, c, f, k);
        if (o.v) {
            if (ws) stv2(s_cur + pz + o.j, c.x, c.y);
            zs = zs + zv.x * c.x;
            zs = zs + zv.y * c.y;
            if (!ws) {   // max|b| and b.b (pcg.rs:35, 46)
                const double a0 = c.x < 0 ? -c.x : c.x, a1 = c.y < 0 ? -c.y : c.y;
                bmax = a0 > bmax ? a0 : bmax;
                bmax = a1 > bmax ? a1 : bmax;
                bb = bb + c.x * c.x;
                bb = bb + c.y * c.y;
            }
        }
        f = c;
        c = k;
    }
    return true;
}

// P2 over one tile: z recomputed from s', x += alpha s', r -= alpha z, r.r and max|r|.  m_x null: the first pass (x = 0)
template <int kMode>
__device__ __forceinline__ bool p2_tile(const Cg3Args &a, const TileGeo &g, Ring3 &rg, const CUtensorMap *m_s, const CUtensorMap *m_x,
                                        const CUtensorMap *m_r, double alpha, double &rr, double &rmax) {
    const Own3 o = own_of(a, g);
    const bool first = m_x == nullptr;
    const unsigned plane = (unsigned)(a.h * a.w);
    const double nalpha = -alpha;
    const int plast = g.z1 < a.d ? g.z1 : a.d - 1, pfirst = g.z0 > 0 ? g.z0 - 1 : 0;
    const unsigned L0 = rg.L - (unsigned)pfirst;
    for (int pq = pfirst; pq <= g.z0 + 2 && pq <= plast; ++pq) request_p2(rg, g, pq, m_s, m_x, m_r);
    double2 f = make_double2(0.0, 0.0);
    if (g.z0 > 0) {
        if (!wait_slot(rg, L0 + g.z0 - 1)) return false;
        f = ldv2(stage_of(rg, L0 + g.z0 - 1) + o.so);
    }
    if (!wait_slot(rg, L0 + g.z0)) return false;
    double2 c = ldv2(stage_of(rg, L0 + g.z0) + o.so);
    unsigned pz = (unsigned)g.z0 * plane;
    for (int z = g.z0; z < g.z1; ++z, pz += plane) {
        const unsigned q = L0 + (unsigned)z;
        const bool more = z + 1 < a.d;
        if (more && !wait_slot(rg, q + 1)) return false;
        __syncthreads();
        if (z + 3 <= plast) request_p2(rg, g, z + 3, m_s, m_x, m_r);
        const double *st = stage_of(rg, q);
        double2 k = make_double2(0.0, 0.0);
        if (more) k = ldv2(stage_of(rg, q + 1) + o.so);
        const double2 zv = lap_pair<kMode>(a, o, st, z, c, f, k);
        if (o.v) {
            double2 xv = make_double2(0.0, 0.0);
            if (!first) xv = ldv2(st + kSPlane + o.xo);
            const double2 rv = ldv2(st + kSPlane + kTX * kTY + o.xo);
            if (first) stv2(a.s0 + pz + o.j, c.x, c.y);                       // pcg.rs:40-42: s = aux = r = b
            stv2(a.x + pz + o.j, xv.x + alpha * c.x, xv.y + alpha * c.y);     // pcg.rs:55
            const double r0 = rv.x + nalpha * zv.x, r1 = rv.y + nalpha * zv.y;   // pcg.rs:56
            stv2(a.r + pz + o.j, r0, r1);
            const double a0 = r0 < 0 ? -r0 : r0, a1 = r1 < 0 ? -r1 : r1;
            rmax = a0 > rmax ? a0 : rmax;
            rmax = a1 > rmax ? a1 : rmax;
            rr = rr + r0 * r0;
            rr = rr + r1 * r1;
        }
        f = c;
        c = k;
    }
    return true;
}

__global__ void __launch_bounds__(kTileThreads, 2) k3_cg_tile(const __grid_constant__ Cg3TileArgs ta) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ double scratch[32];
    __shared__ int s_flag;
    __shared__ int s_next[2];
    __shared__ __align__(8) uint64_t full[kStages];
    const Cg3Args &a = ta.c;
    const int G = gridDim.x;
    const int ntiles = a.tiles_x * a.tiles_y * a.tiles_z;   // >= G
    double *pA = a.partials, *pB = pA + ntiles, *pC = pA + 2 * ntiles, *pD = pA + 3 * ntiles, *pE = pA + 4 * ntiles;   // per TILE
    unsigned long long nbar = 0, phase = 0;
    double sigma = 0, alpha = 0, beta = 0, rmax = 0, bmax = 0;
    int it = 0, applies = 0;
    bool converged = false;
    double *s_cur = a.s0, *s_old = a.s1;
    const CUtensorMap *ms_cur = &ta.mr_s0, *ms_old = &ta.mr_s1;
    const size_t n = (size_t)a.d * a.h * a.w;
    Ring3 rg;
    rg.smem = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    rg.full = full;
    rg.err = &a.ctl->error;
    rg.L = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        pano_sm100::fence_mbar_init();
    }
    __syncthreads();

    for (it = 0; it < a.max_iter; ++it) {
        const bool first = it == 0;
        // ---------------------------------------------------------------- P1
        {
            const unsigned long long base = phase * (unsigned long long)ntiles;
            const CUtensorMap *m_r = first ? &ta.mr_b : &ta.mr_r, *m_s = first ? nullptr : ms_old;
            int t = blockIdx.x;
            for (int i = 0; t < ntiles; ++i) {
                if (threadIdx.x == 0) s_next[i & 1] = G + (int)(atomicAdd(ta.claim, 1ULL) - base);   // claimed a tile ahead
                const TileGeo g = tile_geo(a, t);
                double zs = 0, bb = 0, bm = 0;
                bool ok;
                if (g.mode == 0) ok = p1_tile<0>(a, g, rg, m_r, m_s, s_cur, beta, zs, bb, bm);
                else if (g.mode == 1) ok = p1_tile<1>(a, g, rg, m_r, m_s, s_cur, beta, zs, bb, bm);
                else ok = p1_tile<2>(a, g, rg, m_r, m_s, s_cur, beta, zs, bb, bm);
                if (!ok) return;
                const double v = block_sum(zs, scratch);     // (its barriers also fence the stages against the next tile's requests)
                if (threadIdx.x == 0) pA[t] = v;
                if (first) {
                    const double v2 = block_sum(bb, scratch), v3 = block_max(bm, scratch);
                    if (threadIdx.x == 0) {
                        pB[t] = v2;
                        pC[t] = v3;
                    }
                }
                __syncthreads();
                t = s_next[i & 1];
            }
            ++phase;
        }
        if (!grid_barrier(a.ctl, (++nbar) * (unsigned long long)G, &s_flag)) return;
        const double zs_all = sum_partials<double>(pA, ntiles, scratch);
        if (first) {
            sigma = sum_partials<double>(pB, ntiles, scratch);   // pcg.rs:46
            bmax = max_partials<double>(pC, ntiles, scratch);    // pcg.rs:35
            rmax = bmax;
            if (bmax < a.threshold) {                            // early out: x = 0, scratch untouched
                for (size_t j = blockIdx.x * (size_t)kTileThreads + threadIdx.x; j < n; j += (size_t)G * kTileThreads) a.x[j] = 0.0;
                if (blockIdx.x == 0 && threadIdx.x == 0) {
                    a.ctl->iterations = -1;
                    a.ctl->applies = 0;
                    a.ctl->final_residual = bmax;
                    a.ctl->rhs_max = bmax;
                }
                return;
            }
        }
        ++applies;
        alpha = sigma / zs_all;                                  // pcg.rs:53
        // ---------------------------------------------------------------- P2
        {
            const CUtensorMap *m_s = first ? &ta.mr_b : ms_cur, *m_x = first ? nullptr : &ta.mi_x, *m_r = first ? &ta.mi_b : &ta.mi_r;
            const unsigned long long base = phase * (unsigned long long)ntiles;
            int t = blockIdx.x;
            for (int i = 0; t < ntiles; ++i) {
                if (threadIdx.x == 0) s_next[i & 1] = G + (int)(atomicAdd(ta.claim, 1ULL) - base);
                const TileGeo g = tile_geo(a, t);
                double rr = 0, rm = 0;
                bool ok;
                if (g.mode == 0) ok = p2_tile<0>(a, g, rg, m_s, m_x, m_r, alpha, rr, rm);
                else if (g.mode == 1) ok = p2_tile<1>(a, g, rg, m_s, m_x, m_r, alpha, rr, rm);
                else ok = p2_tile<2>(a, g, rg, m_s, m_x, m_r, alpha, rr, rm);
                if (!ok) return;
                const double v = block_sum(rr, scratch), v2 = block_max(rm, scratch);
                if (threadIdx.x == 0) {
                    pD[t] = v;
                    pE[t] = v2;
                }
                __syncthreads();
                t = s_next[i & 1];
            }
            ++phase;
        }
        if (!grid_barrier(a.ctl, (++nbar) * (unsigned long long)G, &s_flag)) return;
        const double rr_all = sum_partials<double>(pD, ntiles, scratch);
        rmax = max_partials<double>(pE, ntiles, scratch);        // pcg.rs:58
        if (rmax < a.threshold) {                                // pcg.rs:60-63
            converged = true;
            break;
        }
        beta = rr_all / sigma;                                   // pcg.rs:67-68
        sigma = rr_all;                                          // pcg.rs:79
        double *tmp = s_cur;
        s_cur = s_old;
        s_old = tmp;
        const CUtensorMap *tm = ms_cur;
        ms_cur = ms_old;
        ms_old = tm;
    }
    const double *s_fin = converged ? s_cur : s_old;
    if (a.max_iter > 0 && (!converged || s_fin != a.s0)) {       // what the reference leaves in `search` (pcg.rs:72-77), as k3_cg
        for (size_t j = blockIdx.x * (size_t)kTileThreads + threadIdx.x; j < n; j += (size_t)G * kTileThreads) {
            const double sv = s_fin[j];
            a.s0[j] = converged ? sv : a.r[j] + beta * sv;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.ctl->iterations = converged ? it : a.max_iter;
        a.ctl->applies = applies;
        a.ctl->final_residual = rmax;
        a.ctl->rhs_max = bmax;
    }
}

// (w, h, d) f64 tensor, boxes of box_w x box_h x 1 elements
int make_tensor_map_3d(CUtensorMap *map, const void *base, uint64_t w, uint64_t h, uint64_t d, uint32_t box_w, uint32_t box_h) {
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
    const cuuint64_t dims[3] = {w, h, d};
    const cuuint64_t strides[2] = {w * 8, w * h * 8};
    const cuuint32_t box[3] = {box_w, box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS)
        PANO_FAIL(PANO_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d (%llu x %llu x %llu, box %u x %u)", (int)rc,
                  (unsigned long long)d, (unsigned long long)h, (unsigned long long)w, box_w, box_h);
    return PANO_OK;
}

// ---------------------------------------------------------------------------------- host side
int check3(const pano_field *f, int kind, const char *name) {
    PANO_TRY(pano_check_field(f, name));
    if (f->kind != kind || f->dep == 0)
        PANO_FAIL(PANO_ERR_SHAPE, "%s: expected a %s of a Grid3d, got kind %d", name, kind == PANO_CELL3 ? "cell field" : "face field", f->kind);
    if (f->dtype != PANO_F64) PANO_FAIL(PANO_ERR_SHAPE, "%s: the Grid3d path is f64 only", name);
    return PANO_OK;
}
int same_grid3(const pano_field *a, const pano_field *b, const char *what) {
    if (a->ctx != b->ctx) PANO_FAIL(PANO_ERR_INVALID, "%s: fields belong to different contexts", what);
    if (a->dep != b->dep || a->h != b->h || a->w != b->w)
        PANO_FAIL(PANO_ERR_SHAPE, "%s: grid mismatch: %zux%zux%zu vs %zux%zux%zu", what, a->dep, a->h, a->w, b->dep, b->h, b->w);
    return PANO_OK;
}
int check_box_within(const pano_box &r, size_t d, size_t h, size_t w, const char *what) {
    if (r.z0 < 0 || r.y0 < 0 || r.x0 < 0 || r.z1 < r.z0 || r.y1 < r.y0 || r.x1 < r.x0) PANO_FAIL(PANO_ERR_INVALID, "%s: malformed box", what);
    if (r.z1 > r.z0 && r.y1 > r.y0 && r.x1 > r.x0 && ((size_t)r.z1 > d || (size_t)r.y1 > h || (size_t)r.x1 > w))
        PANO_FAIL(PANO_ERR_SHAPE, "%s: box exceeds the %zux%zux%zu grid (an index panic in Rust)", what, d, h, w);
    return PANO_OK;
}
struct Face3 {
    double *vz, *vy, *vx;
};
Face3 split3(const pano_field *f) {   // Simplex1::split (dec/grid.rs:48-61) with a third part in front
    double *p = (double *)f->d;
    const size_t nz = (f->dep + 1) * f->h * f->w, ny = f->dep * (f->h + 1) * f->w;
    return Face3{p, p + nz, p + nz + ny};
}
dim3 grid3(size_t xs, size_t ys, size_t zs) { return dim3((unsigned)((xs + 31) / 32), (unsigned)((ys + 7) / 8), (unsigned)zs); }

int advect3_launch(pano_ctx *ctx, bool scalar, bool mac, double *q_dst, pano_field *mac_dst, const double *q_src, const pano_field *mac_src,
                   const pano_field *vel, double dt) {
    const size_t d = vel->dep, h = vel->h, w = vel->w;
    const Face3 v = split3(vel);
    Face3 md{nullptr, nullptr, nullptr}, ms{nullptr, nullptr, nullptr};
    if (mac) {
        md = split3(mac_dst);
        ms = split3(mac_src);
    }
    const dim3 g = grid3(w + 1, h + 1, d + 1);
    if (g.y > 65535u || g.z > 65535u) PANO_FAIL(PANO_ERR_SHAPE, "Grid3d: %zux%zux%zu exceeds the launch grid", d, h, w);
    if (scalar && mac)
        k3_advect<true, true><<<g, kThreads, 0, ctx->stream>>>(q_dst, md.vz, md.vy, md.vx, q_src, ms.vz, ms.vy, ms.vx, v.vz, v.vy, v.vx, (int)d, (int)h, (int)w, dt);
    else if (scalar)
        k3_advect<true, false><<<g, kThreads, 0, ctx->stream>>>(q_dst, nullptr, nullptr, nullptr, q_src, nullptr, nullptr, nullptr, v.vz, v.vy, v.vx, (int)d, (int)h, (int)w, dt);
    else
        k3_advect<false, true><<<g, kThreads, 0, ctx->stream>>>(nullptr, md.vz, md.vy, md.vx, nullptr, ms.vz, ms.vy, ms.vx, v.vz, v.vy, v.vx, (int)d, (int)h, (int)w, dt);
    return pano_after_launch(ctx, "k3_advect");
}

int neg_div3_launch(pano_ctx *ctx, double *b, const pano_field *vel, pano_box ob) {
    const size_t d = vel->dep, h = vel->h, w = vel->w;
    const Face3 v = split3(vel);
    k3_neg_div<<<grid3(w, h, (d + kDivZ - 1) / kDivZ), kThreads, 0, ctx->stream>>>(b, v.vz, v.vy, v.vx, (int)d, (int)h, (int)w,
                                                                                    clip_box(ob, d + 1, h + 1, w + 1));
    return pano_after_launch(ctx, "k3_neg_div");
}

int project3_launch(pano_ctx *ctx, pano_field *vel, const double *p, double dt) {
    const size_t d = vel->dep, h = vel->h, w = vel->w;
    const Face3 v = split3(vel);
    k3_project<<<grid3(w + 1, h + 1, (d + 1 + kProjZ - 1) / kProjZ), kThreads, 0, ctx->stream>>>(v.vz, v.vy, v.vx, p, (int)d, (int)h, (int)w, dt);
    return pano_after_launch(ctx, "k3_project");
}

int cg3_solve_raw(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, double *s1, size_t d, size_t h, size_t w,
                  int max_iterations, double threshold, double dt, pano_box ob, pano_pcg_info *info) {
    const size_t n = d * h * w;
    if (max_iterations <= 0) {   // pcg.rs:32-46 with an empty loop, as pano_cg_solve_raw
        double bmax = 0.0;
        PANO_TRY(pano_norm_max_raw(ctx, PANO_F64, b, n, &bmax));
        PANO_CUDA(cudaMemsetAsync(x, 0, n * 8, ctx->stream));
        const bool early = bmax < threshold;
        if (!early) {
            PANO_CUDA(cudaMemcpyAsync(r, b, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            PANO_CUDA(cudaMemcpyAsync(s0, b, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        if (info) {
            PANO_CUDA(cudaStreamSynchronize(ctx->stream));
            *info = pano_pcg_info{early ? -1 : 0, 0, bmax, bmax};
        }
        return PANO_OK;
    }
    // "cg3_kernel": 0 auto (the shared-memory plane-tile kernel for even widths, else the column kernel) / 1 column kernel / 2 plane tiles
    const int64_t want = pano_option(ctx, "cg3_kernel", 0);
    const bool tile_ok = w % 2 == 0;
    if (want == 2 && !tile_ok) PANO_FAIL(PANO_ERR_INVALID, "cg3_kernel=2 (plane tiles) needs an even width (%zu given)", w);
    const bool tiled = tile_ok && want != 1;
    int per_sm = 0;
    const bool cg_loads = pano_option(ctx, "cg_ldcg", 0) != 0;
    const void *fn = tiled ? (const void *)k3_cg_tile : cg_loads ? (const void *)k3_cg<true> : (const void *)k3_cg<false>;
    const int nthreads = tiled ? kTileThreads : kThreads;
    const size_t dyn_smem = tiled ? kTileSmemBytes : 0;
    if (tiled) PANO_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    PANO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, nthreads, dyn_smem));
    if (per_sm < 1) PANO_FAIL(PANO_ERR_CUDA, "cg3: kernel does not fit on an SM");
    const int64_t cap = pano_option(ctx, "cg_blocks_per_sm", 0);
    if (cap > 0 && cap < per_sm) per_sm = (int)cap;
    const int tw = tiled ? kTX : 32, th = tiled ? kTY : 8;
    Cg3Args a{x, b, r, s0, s1, (int)d, (int)h, (int)w, dt, threshold, max_iterations, clip_box(ob, d + 1, h + 1, w + 1), nullptr, ctx->d_cg,
              (int)((w + tw - 1) / tw), (int)((h + th - 1) / th), 0, 0};
    const int64_t want_zc = pano_option(ctx, "cg3_zc", 0);
    const int Gfull = ctx->num_sms * per_sm;
    if (tiled) {
        // tiles are claimed dynamically: long columns (two extra planes and one pipeline fill per tile) as long as every CTA still
        // gets a few tiles to even out the tail.  256^3 on B200: zc 8 / 16 / 32 / 64 = 13.7 / 12.9 / 13.5 / 13.1 ms per 62-iteration solve
        int zc = 32;
        while (zc > 16 && ((long long)((d + zc - 1) / zc) * a.tiles_x * a.tiles_y) < 3LL * Gfull) zc /= 2;
        while (zc > 4 && ((long long)((d + zc - 1) / zc) * a.tiles_x * a.tiles_y) < (long long)Gfull) zc /= 2;   // small grids: fill the GPU first
        a.zc = want_zc > 0 ? (int)want_zc : zc;
    } else {
        // static round-robin: pick the column length with the best (useful planes / planes read) x (tiles / rounded-up tiles per CTA)
        int best_zc = 8;
        double best = -1.0;
        for (int zc : {4, 6, 8, 10, 12, 16, 20, 24, 32, 48, 64}) {
            if (want_zc > 0 && zc != want_zc) continue;
            const long long tz = ((long long)d + zc - 1) / zc, nt = tz * a.tiles_x * a.tiles_y;
            const long long rounds = (nt + Gfull - 1) / Gfull;   // every SM slot should be busy in every round: the SMs share the HBM
            const double eff = ((double)zc / (zc + 1.0)) * ((double)nt / (double)(rounds * Gfull));
            if (eff > best) {
                best = eff;
                best_zc = zc;
            }
        }
        a.zc = want_zc > 0 && best < 0 ? (int)want_zc : best_zc;
    }
    a.tiles_z = (int)((d + a.zc - 1) / a.zc);
    const long long ntiles = (long long)a.tiles_x * a.tiles_y * a.tiles_z;
    int G = Gfull;
    if (G > ntiles) G = (int)ntiles;
    if (G < 1) G = 1;
    PANO_TRY(pano_ensure_partials(ctx, 5 * (size_t)(tiled ? ntiles : G)));
    a.partials = ctx->d_partials;
    PANO_TRY(pano_cg_control_reset(ctx));
    if (tiled) {
        if (!ctx->d_claim) PANO_CUDA(cudaMalloc((void **)&ctx->d_claim, 4 * sizeof(unsigned long long)));
        PANO_CUDA(cudaMemsetAsync(ctx->d_claim, 0, 4 * sizeof(unsigned long long), ctx->stream));
        Cg3TileArgs ta;
        memset(&ta, 0, sizeof(ta));
        ta.c = a;
        ta.claim = ctx->d_claim;
        PANO_TRY(make_tensor_map_3d(&ta.mr_b, b, w, h, d, kSW, kTY + 2));
        PANO_TRY(make_tensor_map_3d(&ta.mr_r, r, w, h, d, kSW, kTY + 2));
        PANO_TRY(make_tensor_map_3d(&ta.mr_s0, s0, w, h, d, kSW, kTY + 2));
        PANO_TRY(make_tensor_map_3d(&ta.mr_s1, s1, w, h, d, kSW, kTY + 2));
        PANO_TRY(make_tensor_map_3d(&ta.mi_x, x, w, h, d, kTX, kTY));
        PANO_TRY(make_tensor_map_3d(&ta.mi_r, r, w, h, d, kTX, kTY));
        PANO_TRY(make_tensor_map_3d(&ta.mi_b, b, w, h, d, kTX, kTY));
        void *kargs[] = {(void *)&ta};
        PANO_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)G), dim3(nthreads), kargs, dyn_smem, ctx->stream));
    } else {
        void *kargs[] = {(void *)&a};
        PANO_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)G), dim3(nthreads), kargs, 0, ctx->stream));
    }
    PANO_TRY(pano_after_launch(ctx, "k3_cg"));
    if (info) {
        PANO_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(PanoCgControl), cudaMemcpyDeviceToHost, ctx->stream));
        PANO_CUDA(cudaStreamSynchronize(ctx->stream));
        PANO_TRY(pano_check_device_error(ctx, "pano_pcg3_solve"));
        *info = pano_pcg_info{ctx->h_cg->iterations, ctx->h_cg->applies, ctx->h_cg->final_residual, ctx->h_cg->rhs_max};
    }
    return PANO_OK;
}

}  // namespace

extern "C" {

int pano_field3_num_elem(int kind, size_t d, size_t h, size_t w, size_t *n) {
    if (!n) PANO_FAIL(PANO_ERR_INVALID, "pano_field3_num_elem: null out pointer");
    if (kind == PANO_CELL3) *n = d * h * w;
    else if (kind == PANO_FACE3) *n = (d + 1) * h * w + d * (h + 1) * w + d * h * (w + 1);
    else PANO_FAIL(PANO_ERR_INVALID, "pano_field3_num_elem: bad kind %d", kind);
    return PANO_OK;
}

int pano_field3_new(pano_ctx *ctx, int kind, size_t d, size_t h, size_t w, pano_field **out) {
    if (!ctx || !out) PANO_FAIL(PANO_ERR_INVALID, "pano_field3_new: null argument");
    *out = nullptr;
    if (kind != PANO_CELL3 && kind != PANO_FACE3) PANO_FAIL(PANO_ERR_INVALID, "pano_field3_new: bad kind %d", kind);
    if (d == 0 || h == 0 || w == 0) PANO_FAIL(PANO_ERR_SHAPE, "pano_field3_new: empty grid %zux%zux%zu", d, h, w);
    if (d > 4096 || h > 65536 || w > 65536 || (d + 1) * (h + 1) * (w + 1) >= ((size_t)1 << 31))
        PANO_FAIL(PANO_ERR_INVALID, "pano_field3_new: grid %zux%zux%zu exceeds 2^31 samples per array", d, h, w);
    PANO_TRY(pano_activate(ctx));
    pano_field *f = new pano_field();
    f->ctx = ctx;
    f->kind = kind;
    f->dtype = PANO_F64;
    f->dep = d;
    f->h = h;
    f->w = w;
    pano_field3_num_elem(kind, d, h, w, &f->n);
    const size_t bytes = f->n * 8;
    cudaError_t e = cudaMalloc(&f->d, bytes + 256);
    if (e == cudaSuccess) e = cudaMemsetAsync(f->d, 0, bytes + 256, ctx->stream);
    if (e != cudaSuccess) {
        if (f->d) cudaFree(f->d);
        delete f;
        PANO_FAIL(PANO_ERR_CUDA, "pano_field3_new: allocating %zu bytes -> %s", bytes, cudaGetErrorString(e));
    }
    pano_ctx_field_born(ctx);
    *out = f;
    return PANO_OK;
}

int pano_field3_dim(const pano_field *f, size_t *d, size_t *h, size_t *w) {
    PANO_TRY(pano_check_field(f, "pano_field3_dim"));
    if (f->dep == 0) PANO_FAIL(PANO_ERR_SHAPE, "pano_field3_dim: a 2-D field");
    if (d) *d = f->dep;
    if (h) *h = f->h;
    if (w) *w = f->w;
    return PANO_OK;
}

int pano_field3_fill_box(pano_field *f, int comp, pano_box box, double value) {
    PANO_TRY(pano_check_field(f, "pano_field3_fill_box"));
    if (f->dep == 0 || f->dtype != PANO_F64) PANO_FAIL(PANO_ERR_SHAPE, "pano_field3_fill_box: not a Grid3d field");
    pano_ctx *ctx = f->ctx;
    PANO_TRY(pano_activate(ctx));
    struct Part { size_t off, D, H, W; };
    Part parts[3];
    int np = 0;
    const size_t d = f->dep, h = f->h, w = f->w;
    if (f->kind == PANO_CELL3) {
        if (comp != PANO_COMP_ALL) PANO_FAIL(PANO_ERR_INVALID, "pano_field3_fill_box: component %d on a cell field", comp);
        parts[np++] = Part{0, d, h, w};
    } else {
        const size_t nz = (d + 1) * h * w, ny = d * (h + 1) * w;
        if (comp == PANO_COMP_ALL || comp == PANO_COMP_VZ) parts[np++] = Part{0, d + 1, h, w};
        if (comp == PANO_COMP_ALL || comp == PANO_COMP_VY) parts[np++] = Part{nz, d, h + 1, w};
        if (comp == PANO_COMP_ALL || comp == PANO_COMP_VX) parts[np++] = Part{nz + ny, d, h, w + 1};
        if (np == 0) PANO_FAIL(PANO_ERR_INVALID, "pano_field3_fill_box: bad component %d", comp);
    }
    for (int i = 0; i < np; ++i) PANO_TRY(check_box_within(box, parts[i].D, parts[i].H, parts[i].W, "pano_field3_fill_box"));
    if (box.z1 == box.z0 || box.y1 == box.y0 || box.x1 == box.x0) return PANO_OK;
    const size_t cells = (size_t)(box.z1 - box.z0) * (size_t)(box.y1 - box.y0) * (size_t)(box.x1 - box.x0);
    size_t g = (cells + kThreads - 1) / kThreads;
    if (g > (size_t)ctx->num_sms * 8) g = (size_t)ctx->num_sms * 8;
    const BoxI bi{(int)box.z0, (int)box.z1, (int)box.y0, (int)box.y1, (int)box.x0, (int)box.x1};
    for (int i = 0; i < np; ++i) {
        k3_fill_box<<<(unsigned)g, kThreads, 0, ctx->stream>>>((double *)f->d + parts[i].off, (int)parts[i].H, (int)parts[i].W, bi, value);
        PANO_TRY(pano_after_launch(ctx, "pano_field3_fill_box"));
    }
    return PANO_OK;
}

double pano_trilinear(double a000, double a001, double a010, double a011, double a100, double a101, double a110, double a111, double s,
                      double t, double u) {
    return pano::trilinear<double>(a000, a001, a010, a011, a100, a101, a110, a111, s, t, u);
}

int pano_advect3(pano_field *dst, const pano_field *src, double timestep, const pano_field *vel) {
    PANO_TRY(check3(dst, PANO_CELL3, "pano_advect3(dst)"));
    PANO_TRY(check3(src, PANO_CELL3, "pano_advect3(src)"));
    PANO_TRY(check3(vel, PANO_FACE3, "pano_advect3(vel)"));
    PANO_TRY(same_grid3(dst, src, "pano_advect3"));
    PANO_TRY(same_grid3(dst, vel, "pano_advect3"));
    if (dst->d == src->d) PANO_FAIL(PANO_ERR_INVALID, "pano_advect3: dst aliases src");
    PANO_TRY(pano_activate(dst->ctx));
    return advect3_launch(dst->ctx, true, false, (double *)dst->d, nullptr, (const double *)src->d, nullptr, vel, timestep);
}

int pano_advect3_mac(pano_field *dst, const pano_field *src, double timestep, const pano_field *vel) {
    PANO_TRY(check3(dst, PANO_FACE3, "pano_advect3_mac(dst)"));
    PANO_TRY(check3(src, PANO_FACE3, "pano_advect3_mac(src)"));
    PANO_TRY(check3(vel, PANO_FACE3, "pano_advect3_mac(vel)"));
    PANO_TRY(same_grid3(dst, src, "pano_advect3_mac"));
    PANO_TRY(same_grid3(dst, vel, "pano_advect3_mac"));
    if (dst->d == src->d || dst->d == vel->d) PANO_FAIL(PANO_ERR_INVALID, "pano_advect3_mac: dst aliases src or vel");
    PANO_TRY(pano_activate(dst->ctx));
    return advect3_launch(dst->ctx, false, true, nullptr, dst, nullptr, src, vel, timestep);
}

int pano_advect3_all(pano_field *q_dst, pano_field *vel_dst, const pano_field *q_src, const pano_field *vel, double timestep) {
    PANO_TRY(check3(q_dst, PANO_CELL3, "pano_advect3_all(q_dst)"));
    PANO_TRY(check3(q_src, PANO_CELL3, "pano_advect3_all(q_src)"));
    PANO_TRY(check3(vel_dst, PANO_FACE3, "pano_advect3_all(vel_dst)"));
    PANO_TRY(check3(vel, PANO_FACE3, "pano_advect3_all(vel)"));
    PANO_TRY(same_grid3(q_dst, q_src, "pano_advect3_all"));
    PANO_TRY(same_grid3(q_dst, vel, "pano_advect3_all"));
    PANO_TRY(same_grid3(q_dst, vel_dst, "pano_advect3_all"));
    if (q_dst->d == q_src->d || vel_dst->d == vel->d) PANO_FAIL(PANO_ERR_INVALID, "pano_advect3_all: destination aliases source");
    PANO_TRY(pano_activate(q_dst->ctx));
    return advect3_launch(q_dst->ctx, true, true, (double *)q_dst->d, vel_dst, (const double *)q_src->d, vel, vel, timestep);
}

int pano_neg_divergence3(pano_field *b, const pano_field *vel, pano_box obstacle, double *rhs_max) {
    PANO_TRY(check3(b, PANO_CELL3, "pano_neg_divergence3(b)"));
    PANO_TRY(check3(vel, PANO_FACE3, "pano_neg_divergence3(vel)"));
    PANO_TRY(same_grid3(b, vel, "pano_neg_divergence3"));
    PANO_TRY(check_box_within(obstacle, b->dep, b->h, b->w, "pano_neg_divergence3(obstacle)"));
    PANO_TRY(pano_activate(b->ctx));
    PANO_TRY(neg_div3_launch(b->ctx, (double *)b->d, vel, obstacle));
    if (rhs_max) return pano_norm_max_raw(b->ctx, PANO_F64, b->d, b->n, rhs_max);
    return PANO_OK;
}

int pano_laplacian3_apply(pano_field *z, const pano_field *s, double timestep, pano_box obstacle) {
    PANO_TRY(check3(z, PANO_CELL3, "pano_laplacian3_apply(z)"));
    PANO_TRY(check3(s, PANO_CELL3, "pano_laplacian3_apply(s)"));
    PANO_TRY(same_grid3(z, s, "pano_laplacian3_apply"));
    if (z->d == s->d) PANO_FAIL(PANO_ERR_INVALID, "pano_laplacian3_apply: z aliases s");
    PANO_TRY(check_box_within(obstacle, z->dep, z->h, z->w, "pano_laplacian3_apply(obstacle)"));
    pano_ctx *ctx = z->ctx;
    PANO_TRY(pano_activate(ctx));
    const size_t d = z->dep, h = z->h, w = z->w;
    k3_laplacian<<<grid3(w, h, d), kThreads, 0, ctx->stream>>>((double *)z->d, (const double *)s->d, (int)d, (int)h, (int)w, timestep,
                                                                clip_box(obstacle, d + 1, h + 1, w + 1));
    return pano_after_launch(ctx, "k3_laplacian");
}

int pano_project3(pano_field *vel, const pano_field *pressure, double timestep) {
    PANO_TRY(check3(vel, PANO_FACE3, "pano_project3(vel)"));
    PANO_TRY(check3(pressure, PANO_CELL3, "pano_project3(pressure)"));
    PANO_TRY(same_grid3(vel, pressure, "pano_project3"));
    PANO_TRY(pano_activate(vel->ctx));
    return project3_launch(vel->ctx, vel, (const double *)pressure->d, timestep);
}

int pano_pcg3_solve(int precond, pano_field *x, const pano_field *b, int32_t max_iterations, double threshold, pano_field *residual,
                    pano_field *auxiliary, pano_field *search, double timestep, pano_box obstacle, pano_pcg_info *info) {
    if (precond != PANO_PRECOND_IDENTITY)
        PANO_FAIL(PANO_ERR_UNIMPLEMENTED, "pano_pcg3_solve: only the identity preconditioner `()` (pcg.rs:8-12) exists on a Grid3d");
    const pano_field *all[] = {x, b, residual, auxiliary, search};
    const char *names[] = {"x", "b", "residual", "auxiliary", "search"};
    for (int i = 0; i < 5; ++i) {
        char nm[64];
        snprintf(nm, sizeof(nm), "pano_pcg3_solve(%s)", names[i]);
        PANO_TRY(check3(all[i], PANO_CELL3, nm));
        PANO_TRY(same_grid3(all[0], all[i], "pano_pcg3_solve"));
        for (int j = 0; j < i; ++j)
            if (all[i]->d == all[j]->d) PANO_FAIL(PANO_ERR_INVALID, "pano_pcg3_solve: %s aliases %s", names[i], names[j]);
    }
    PANO_TRY(check_box_within(obstacle, x->dep, x->h, x->w, "pano_pcg3_solve(obstacle)"));
    PANO_TRY(pano_activate(x->ctx));
    return cg3_solve_raw(x->ctx, (double *)x->d, (const double *)b->d, (double *)residual->d, (double *)search->d, (double *)auxiliary->d,
                         x->dep, x->h, x->w, max_iterations, threshold, timestep, obstacle, info);
}

// after_advect (nullable) runs once the new density is final: the host-buffer entry point starts its download there
typedef int (*Step3Hook)(pano_ctx *ctx, void *user);
static int fluid3_step_impl(const pano_step3_params *params, pano_field *density, pano_field *vel, pano_field *pressure, pano_field *temp,
                            pano_field *vel_temp, pano_field *residual, pano_field *auxiliary, pano_field *search, pano_pcg_info *info,
                            Step3Hook after_advect, void *user) {
    if (!params) PANO_FAIL(PANO_ERR_INVALID, "pano_fluid3_step: null params");
    const pano_field *c3[] = {density, pressure, temp, residual, auxiliary, search};
    const char *n3[] = {"density", "pressure", "temp", "residual", "auxiliary", "search"};
    for (int i = 0; i < 6; ++i) {
        char nm[64];
        snprintf(nm, sizeof(nm), "pano_fluid3_step(%s)", n3[i]);
        PANO_TRY(check3(c3[i], PANO_CELL3, nm));
        PANO_TRY(same_grid3(density, c3[i], "pano_fluid3_step"));
        for (int j = 0; j < i; ++j)
            if (c3[i]->d == c3[j]->d) PANO_FAIL(PANO_ERR_INVALID, "pano_fluid3_step: %s aliases %s", n3[i], n3[j]);
    }
    PANO_TRY(check3(vel, PANO_FACE3, "pano_fluid3_step(vel)"));
    PANO_TRY(check3(vel_temp, PANO_FACE3, "pano_fluid3_step(vel_temp)"));
    PANO_TRY(same_grid3(density, vel, "pano_fluid3_step"));
    PANO_TRY(same_grid3(density, vel_temp, "pano_fluid3_step"));
    if (vel->d == vel_temp->d) PANO_FAIL(PANO_ERR_INVALID, "pano_fluid3_step: vel aliases vel_temp");
    if (params->precond != PANO_PRECOND_IDENTITY)
        PANO_FAIL(PANO_ERR_UNIMPLEMENTED, "pano_fluid3_step: only the identity preconditioner exists on a Grid3d");
    const size_t d = density->dep, h = density->h, w = density->w;
    if (d < 2 || h < 2 || w < 2) PANO_FAIL(PANO_ERR_SHAPE, "pano_fluid3_step: grid %zux%zux%zu below 2x2x2", d, h, w);
    PANO_TRY(check_box_within(params->inflow, d, h, w, "pano_fluid3_step(inflow)"));
    PANO_TRY(check_box_within(params->obstacle, d, h, w, "pano_fluid3_step(obstacle)"));
    pano_ctx *ctx = density->ctx;
    PANO_TRY(pano_activate(ctx));
    const double dt = params->timestep;
    PANO_TRY(pano_phase_mark(ctx, 0));
    PANO_TRY(pano_field3_fill_box(density, PANO_COMP_ALL, params->inflow, params->inflow_density));   // :48-57
    PANO_TRY(pano_field3_fill_box(vel, PANO_COMP_VY, params->inflow, params->inflow_vy));
    PANO_TRY(pano_phase_mark(ctx, 1));
    PANO_TRY(advect3_launch(ctx, true, true, (double *)temp->d, vel_temp, (const double *)density->d, vel, vel, dt));   // :59-60
    PANO_TRY(pano_field_swap(density, temp));                                                          // :62-63
    PANO_TRY(pano_field_swap(vel, vel_temp));
    if (after_advect) PANO_TRY(after_advect(ctx, user));
    PANO_TRY(pano_phase_mark(ctx, 2));
    PANO_TRY(neg_div3_launch(ctx, (double *)temp->d, vel, params->obstacle));                           // :69-83
    PANO_TRY(pano_phase_mark(ctx, 3));
    PANO_TRY(cg3_solve_raw(ctx, (double *)pressure->d, (const double *)temp->d, (double *)residual->d, (double *)search->d,
                           (double *)auxiliary->d, d, h, w, params->max_iterations, params->threshold, dt, params->obstacle, nullptr));
    PANO_TRY(pano_phase_mark(ctx, 4));
    PANO_TRY(project3_launch(ctx, vel, (const double *)pressure->d, dt));                               // :124-141
    PANO_TRY(pano_phase_mark(ctx, 5));
    if (info) {
        if (params->max_iterations <= 0) {
            PANO_CUDA(cudaStreamSynchronize(ctx->stream));
            double bmax = 0.0;
            PANO_TRY(pano_norm_max_raw(ctx, PANO_F64, temp->d, temp->n, &bmax));
            *info = pano_pcg_info{bmax < params->threshold ? -1 : 0, 0, bmax, bmax};
            return PANO_OK;
        }
        PANO_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(PanoCgControl), cudaMemcpyDeviceToHost, ctx->stream));
        PANO_CUDA(cudaStreamSynchronize(ctx->stream));
        PANO_TRY(pano_check_device_error(ctx, "pano_fluid3_step"));
        *info = pano_pcg_info{ctx->h_cg->iterations, ctx->h_cg->applies, ctx->h_cg->final_residual, ctx->h_cg->rhs_max};
    }
    return PANO_OK;
}

int pano_fluid3_step(const pano_step3_params *params, pano_field *density, pano_field *vel, pano_field *pressure, pano_field *temp,
                     pano_field *vel_temp, pano_field *residual, pano_field *auxiliary, pano_field *search, pano_pcg_info *info) {
    return fluid3_step_impl(params, density, vel, pressure, temp, vel_temp, residual, auxiliary, search, info, nullptr, nullptr);
}

struct Host3Copy {
    PanoWorkspace *ws;
    double *density_host;
    size_t bytes;
};
static int start_density3_download(pano_ctx *ctx, void *user) {
    Host3Copy *c = static_cast<Host3Copy *>(user);
    PANO_CUDA(cudaEventRecord(ctx->ev_advect, ctx->stream));
    PANO_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_advect, 0));
    PANO_CUDA(cudaMemcpyAsync(c->density_host, c->ws->density->d, c->bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    PANO_CUDA(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
    return PANO_OK;
}
static int get_workspace3(pano_ctx *ctx, size_t d, size_t h, size_t w, PanoWorkspace **out) {
    const auto key = std::make_pair(d, std::make_pair(h, w));
    auto it = ctx->workspaces3.find(key);
    if (it != ctx->workspaces3.end()) {
        *out = it->second;
        return PANO_OK;
    }
    PanoWorkspace *ws = new PanoWorkspace();
    ws->dep = d; ws->h = h; ws->w = w;
    struct { pano_field **f; int kind; } want[] = {
        {&ws->density, PANO_CELL3}, {&ws->vel, PANO_FACE3}, {&ws->pressure, PANO_CELL3}, {&ws->temp, PANO_CELL3},
        {&ws->vel_temp, PANO_FACE3}, {&ws->residual, PANO_CELL3}, {&ws->auxiliary, PANO_CELL3}, {&ws->search, PANO_CELL3}};
    for (auto &wf : want) {          // cached only when complete, as the 2-D workspace
        const int rc = pano_field3_new(ctx, wf.kind, d, h, w, wf.f);
        if (rc != PANO_OK) {
            for (auto &g : want) pano_field_free(*g.f);
            delete ws;
            return rc;
        }
    }
    ctx->workspaces3[key] = ws;
    *out = ws;
    return PANO_OK;
}

int pano_fluid3_step_host(pano_ctx *ctx, const pano_step3_params *params, size_t d, size_t h, size_t w, double *density, double *vel,
                          double *pressure, pano_pcg_info *info) {
    if (!ctx || !params || !density || !vel) PANO_FAIL(PANO_ERR_INVALID, "pano_fluid3_step_host: null argument");
    PANO_TRY(pano_activate(ctx));
    PanoWorkspace *ws = nullptr;
    PANO_TRY(get_workspace3(ctx, d, h, w, &ws));
    const size_t nc = ws->density->n * sizeof(double), nf = ws->vel->n * sizeof(double);
    PANO_CUDA(cudaMemcpyAsync(ws->density->d, density, nc, cudaMemcpyHostToDevice, ctx->stream));
    PANO_CUDA(cudaMemcpyAsync(ws->vel->d, vel, nf, cudaMemcpyHostToDevice, ctx->stream));
    Host3Copy hook{ws, density, nc};
    PANO_TRY(fluid3_step_impl(params, ws->density, ws->vel, ws->pressure, ws->temp, ws->vel_temp, ws->residual, ws->auxiliary, ws->search,
                              nullptr, start_density3_download, &hook));      // the density goes home under the solve
    PANO_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
    PANO_CUDA(cudaMemcpyAsync(vel, ws->vel->d, nf, cudaMemcpyDeviceToHost, ctx->stream));
    if (pressure) PANO_CUDA(cudaMemcpyAsync(pressure, ws->pressure->d, nc, cudaMemcpyDeviceToHost, ctx->stream));
    PANO_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(PanoCgControl), cudaMemcpyDeviceToHost, ctx->stream));
    PANO_CUDA(cudaStreamSynchronize(ctx->stream));
    PANO_TRY(pano_check_device_error(ctx, "pano_fluid3_step_host"));
    if (info) {
        if (params->max_iterations <= 0) {
            double bmax = 0.0;
            PANO_TRY(pano_norm_max_raw(ctx, PANO_F64, ws->temp->d, ws->temp->n, &bmax));
            *info = pano_pcg_info{bmax < params->threshold ? -1 : 0, 0, bmax, bmax};
        } else {
            *info = pano_pcg_info{ctx->h_cg->iterations, ctx->h_cg->applies, ctx->h_cg->final_residual, ctx->h_cg->rhs_max};
        }
    }
    return PANO_OK;
}

}  // extern "C"
