// pano_mg.cu -- non-identity preconditioners behind the reference's trait seam
//   pub trait Preconditioner<L> { fn apply(&self, dst: &mut L, src: &L); }      panopaea/src/pcg.rs:4-6
// The reference implements it for `()` only (pcg.rs:8-12, the identity); SURVEY.md 8(f) rank 3 asks for more, because
// at N >= 512 its CG runs into the 100-iteration cap without reaching the threshold.  Two are provided, both for the
// operator of examples/dec_fluid.rs:100-119 (A = dt * graph Laplacian over the open faces):
//   * Jacobi:     dst = src / diag(A)
//   * Multigrid:  one V-cycle from a zero guess -- 2x2 aggregation, piecewise-constant transfers, coarse operator
//                 1/2 P^T A P (face weights halve-and-add), damped Jacobi smoothing (omega 0.8, 2 + 2 sweeps, 16 on the
//                 coarsest level).  Symmetric positive definite, so pcg.rs:14-82 applies unchanged.
// The arithmetic is SPECIFIED in DESIGN.md section 5b (plain unfused loops, restated by the test suite's CPU checker);
// every expression below keeps that evaluation order and the library is built with --fmad=false, so apply() is bit-identical to it.
//
// B200 mapping: the fine level never stores face weights (recomputed from the wall/obstacle geometry), sweeps are fused
// (2 pre-sweeps from zero = ONE pass over f; prolongation fused into the first post-sweep), so level 0 costs
// 16 + 26 + 26 + 24 B/cell; every level that fits 32 x 32 cells runs inside ONE single-CTA kernel with block barriers
// instead of ~25 latency-bound launches.
#include <cmath>

#include "pano_internal.cuh"

namespace {

constexpr double kOmega = 0.8;
constexpr int kCoarseSweeps = 16;   // even; the first two are the fused pre-sweep pair
constexpr int kCoarsestMax = 8;     // stop coarsening when max(h, w) <= 8
constexpr int kTailMax = 32;        // levels with max(h, w) <= 32 run in the single-CTA tail kernel
constexpr int kMaxLevels = 24;
constexpr int kTailLevels = 4;      // 32, 16, 8 (+1 spare for odd sizes)
constexpr int kThreads = 256;

// ---- level geometry: face weights n, s, w, e of a cell and od = omega / (dt * (n + s + w + e))
struct GeomStored {
    int h, w;
    const double *wy, *wx, *od;
    __device__ __forceinline__ void weights(int y, int x, double &n, double &s, double &w_, double &e) const {
        n = wy[y * w + x];
        s = wy[(y + 1) * w + x];
        w_ = wx[y * (w + 1) + x];
        e = wx[y * (w + 1) + x + 1];
    }
    __device__ __forceinline__ double odv(int y, int x) const { return od[y * w + x]; }
    __device__ __forceinline__ double face_y(int y, int x) const { return wy[y * w + x]; }   // y in 0..=h
    __device__ __forceinline__ double face_x(int y, int x) const { return wx[y * (w + 1) + x]; }   // x in 0..=w
};

struct Geom0 {   // level 0: weight 1 on open faces (walls and the masked rectangle closed), nothing stored
    int h, w;
    RectI m;
    double od[5];   // omega / (dt * k), od[0] = 0
    __device__ __forceinline__ void weights(int y, int x, double &n, double &s, double &w_, double &e) const {
        const bool in = in_rect(m, y, x);
        n = (y > 0 && !in) ? 1.0 : 0.0;
        s = (y < h - 1 && !in_rect(m, y + 1, x)) ? 1.0 : 0.0;
        w_ = (x > 0 && !in) ? 1.0 : 0.0;
        e = (x < w - 1 && !in_rect(m, y, x + 1)) ? 1.0 : 0.0;
    }
    __device__ __forceinline__ double odv(int y, int x) const {
        const bool in = in_rect(m, y, x);
        const int k = (int)(y > 0 && !in) + (int)(y < h - 1 && !in_rect(m, y + 1, x)) + (int)(x > 0 && !in) +
                      (int)(x < w - 1 && !in_rect(m, y, x + 1));
        return k == 4 ? od[4] : (k == 3 ? od[3] : (k == 2 ? od[2] : (k == 1 ? od[1] : 0.0)));   // selects: a dynamic index would put od[] on the stack
    }
    __device__ __forceinline__ double face_y(int y, int x) const { return (y > 0 && y < h && !in_rect(m, y, x)) ? 1.0 : 0.0; }
    __device__ __forceinline__ double face_x(int y, int x) const { return (x > 0 && x < w && !in_rect(m, y, x)) ? 1.0 : 0.0; }
};

// (A u)[y, x] with u given as a callable; a closed face never evaluates its neighbour (`A u`)
template <class G, class U>
__device__ __forceinline__ double mg_au(const G &g, const U &u, int y, int x, double c, double dt) {
    double n, s, w_, e;
    g.weights(y, x, n, s, w_, e);
    const double tn = n != 0.0 ? n * (c - u(y - 1, x)) : 0.0;
    const double ts = s != 0.0 ? s * (c - u(y + 1, x)) : 0.0;
    const double tw = w_ != 0.0 ? w_ * (c - u(y, x - 1)) : 0.0;
    const double te = e != 0.0 ? e * (c - u(y, x + 1)) : 0.0;
    return (((tn + ts) + tw) + te) * dt;
}

// ---- the four per-cell operations of a V-cycle
// two sweeps from a zero guess, fused: u1 = od * f is formed on the fly for the cell and its neighbours
template <class G>
__device__ __forceinline__ double mg_pre_cell(const G &g, const double *__restrict__ f, int y, int x, double dt) {
    const int w = g.w;
    auto u1 = [&](int yy, int xx) { return g.odv(yy, xx) * f[yy * w + xx]; };
    const double odc = g.odv(y, x), fc = f[y * w + x];
    const double c = odc * fc;
    return c + odc * (fc - mg_au(g, u1, y, x, c, dt));
}
// one sweep: u_in + od * (f - A u_in)
template <class G>
__device__ __forceinline__ double mg_sweep_cell(const G &g, const double *__restrict__ uin, const double *__restrict__ f, int y, int x,
                                                double dt) {
    const int w = g.w;
    auto u = [&](int yy, int xx) { return uin[yy * w + xx]; };
    const double c = uin[y * w + x];
    return c + g.odv(y, x) * (f[y * w + x] - mg_au(g, u, y, x, c, dt));
}
// prolongation fused into the first post-sweep: v = u + e_coarse[parent] formed on the fly
template <class G>
__device__ __forceinline__ double mg_post1_cell(const G &g, const double *__restrict__ uin, const double *__restrict__ ec, int wc,
                                                const double *__restrict__ f, int y, int x, double dt) {
    const int w = g.w;
    auto v = [&](int yy, int xx) { return uin[yy * w + xx] + ec[(yy >> 1) * wc + (xx >> 1)]; };
    const double c = v(y, x);
    return c + g.odv(y, x) * (f[y * w + x] - mg_au(g, v, y, x, c, dt));
}
// residual of the four children of coarse cell (Y, X), summed in the specified order
template <class G>
__device__ __forceinline__ double mg_restrict_cell(const G &g, const double *__restrict__ uin, const double *__restrict__ f, int Y, int X,
                                                   double dt) {
    const int h = g.h, w = g.w;
    auto u = [&](int yy, int xx) { return uin[yy * w + xx]; };
    auto r = [&](int yy, int xx) { return f[yy * w + xx] - mg_au(g, u, yy, xx, uin[yy * w + xx], dt); };
    const int y = 2 * Y, x = 2 * X;
    const double r00 = r(y, x);
    const double r01 = x + 1 < w ? r(y, x + 1) : 0.0;
    const double r10 = y + 1 < h ? r(y + 1, x) : 0.0;
    const double r11 = (y + 1 < h && x + 1 < w) ? r(y + 1, x + 1) : 0.0;
    return ((r00 + r01) + r10) + r11;
}

// ---- one kernel per operation for the levels that live in HBM (32 x 8 cells per block, rows grid-strided)
enum { OP_PRE = 0, OP_SWEEP = 1, OP_POST1 = 2, OP_RESTRICT = 3 };

template <class G, int kOp>
__global__ void __launch_bounds__(kThreads)
k_mg_level(const G g, double *__restrict__ out, const double *__restrict__ uin, const double *__restrict__ f, const double *__restrict__ ec,
           int hc, int wc, double dt) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rows = kOp == OP_RESTRICT ? hc : g.h, cols = kOp == OP_RESTRICT ? wc : g.w;
    if (x >= cols) return;
    for (int y = blockIdx.y * 8 + (threadIdx.x >> 5); y < rows; y += gridDim.y * 8) {
        double v;
        if (kOp == OP_PRE) v = mg_pre_cell(g, f, y, x, dt);
        else if (kOp == OP_SWEEP) v = mg_sweep_cell(g, uin, f, y, x, dt);
        else if (kOp == OP_POST1) v = mg_post1_cell(g, uin, ec, wc, f, y, x, dt);
        else v = mg_restrict_cell(g, uin, f, y, x, dt);
        out[y * cols + x] = v;
    }
}

// ---- fused forms for the levels that live in HBM: a 16 x 64 tile, staged in shared memory with a two-cell halo,
// so that a whole half of the V-cycle is ONE pass over the level:
//   k_mg_down: f -> u (two sweeps from zero), and the restricted residual f' = R (f - A u)          8 B read, 10 B written per cell
//   k_mg_up:   u, u' (coarse correction), f -> two more sweeps of u + P u'                           18 B read, 8 B written
// The halo cells are recomputed by the neighbouring tiles (same expressions, so the same bits); out-of-place (down writes
// `t`, up reads `t` and writes `u`) because a tile reads its neighbours' cells.
// Persistent blocks walk the tiles with a two-stage cp.async pipeline: the loads of tile k+1 are in flight while tile k
// is computed, so the DRAM latency is not paid once per tile (first version: load -> barrier -> compute, 45 % of what
// the same kernels reach now).  Tiles whose faces are all open (`regular`: level 0 away from walls / obstacle; coarser
// levels by a per-tile flag computed at set-up) use constant weights and never read the geometry arrays.
constexpr int FT_H = 16, FT_W = 64;                       // tile (even, so that 2x2 aggregates never straddle tiles)
constexpr int F2_H = FT_H + 4, F2_W = FT_W + 4;           // with the two-cell halo
constexpr int F1_H = FT_H + 2, F1_W = FT_W + 2;           // with the one-cell halo
constexpr int EC_H = FT_H / 2 + 2, EC_W = FT_W / 2 + 2;   // coarse cells under the tile + 2
constexpr int kDownStage = F2_H * F2_W;                               // doubles per pipeline stage of k_mg_down: f
constexpr int kUpStage = 2 * F2_H * F2_W + EC_H * EC_W;               // k_mg_up: u, f, coarse correction
constexpr int kDownSmem = (2 * kDownStage + 2 * F2_H * F2_W) * 8;   // stages, U1, U2 (U2 in the tile + 2 layout)
constexpr int kUpSmem = (2 * kUpStage + F2_H * F2_W) * 8;

struct GeomRegular {   // every face open: weights 1, od = omega / (dt * 4)
    int h, w;
    double od4;
    __device__ __forceinline__ void weights(int, int, double &n, double &s, double &w_, double &e) const { n = s = w_ = e = 1.0; }
    __device__ __forceinline__ double odv(int, int) const { return od4; }
};
// true when every cell of [ya, yb] x [xa, xb] has its four faces open
__device__ __forceinline__ bool region_open(const Geom0 &g, int ya, int yb, int xa, int xb) {
    if (ya < 1 || yb > g.h - 2 || xa < 1 || xb > g.w - 2) return false;
    const RectI &m = g.m;
    return !(m.y1 > m.y0 && m.x1 > m.x0 && ya < m.y1 && yb + 1 >= m.y0 && xa < m.x1 && xb + 1 >= m.x0);
}

// 8-byte asynchronous copy global -> shared; `ok` false: the destination is zero-filled and nothing is read
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// fn(ly, lx) for every cell of an H x W region (W = 64 + 0 / 2 / 4) without a division: 256 threads sweep 4 rows x 64
// columns per pass, the columns beyond 64 are one extra pass of H x (W - 64) threads
template <int H, int W, class Fn>
__device__ __forceinline__ void for_region(const Fn &fn) {
    static_assert(kThreads == 256 && W >= 64 && W <= 68 && (W - 64) * H <= kThreads, "region shape");
    const int tid = threadIdx.x, lx = tid & 63, r = tid >> 6;
#pragma unroll
    for (int k = 0; k < (H + 3) / 4; ++k) {
        const int ly = r + 4 * k;
        if (ly < H) fn(ly, lx);
    }
    constexpr int E = W - 64;
    if (E > 0 && tid < H * E) fn(tid / (E > 0 ? E : 1), 64 + tid % (E > 0 ? E : 1));
}

// stage the (tile + 2) box of a level array
__device__ __forceinline__ void stage_box2(double *dst, const double *__restrict__ src, int h, int w, int ty0, int tx0) {
    for_region<F2_H, F2_W>([&](int ly, int lx) {
        const int y = ty0 - 2 + ly, x = tx0 - 2 + lx;
        const bool ok = y >= 0 && y < h && x >= 0 && x < w;
        cp_async8(dst + ly * F2_W + lx, ok ? src + y * w + x : src, ok);
    });
}

// the same for a box that lies entirely inside the grid and starts on a 16-byte boundary (regular tile, even width):
// 16-byte copies, no bounds, no per-element index arithmetic
__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void stage_box2_vec(double *dst, const double *__restrict__ src, int w, int ty0, int tx0) {
    const double *base = src + (ty0 - 2) * w + tx0 - 2;
    const int tid = threadIdx.x, j = tid & 31, r = tid >> 5;
#pragma unroll
    for (int k = 0; k < (F2_H + 7) / 8; ++k) {
        const int ly = r + 8 * k;
        if (ly < F2_H) cp_async16(dst + ly * F2_W + 2 * j, base + ly * w + 2 * j);
    }
    if (tid < 2 * F2_H) cp_async16(dst + (tid >> 1) * F2_W + 64 + 2 * (tid & 1), base + (tid >> 1) * w + 64 + 2 * (tid & 1));
}

// compute part of k_mg_down for one staged tile: F = f on the tile + 2
template <class G>
__device__ __forceinline__ void mg_down_tile(const G &g, int h, int w, double *__restrict__ uout, double *__restrict__ fc, int wc, double dt,
                                             int ty0, int tx0, const double *F, double *U1, double *U2) {
    // A: u1 = od * f on the tile + 2 (first sweep from zero)
    for_region<F2_H, F2_W>([&](int ly, int lx) {
        const int y = ty0 - 2 + ly, x = tx0 - 2 + lx;
        U1[ly * F2_W + lx] = (y >= 0 && y < h && x >= 0 && x < w) ? g.odv(y, x) * F[ly * F2_W + lx] : 0.0;
    });
    __syncthreads();
    // B: second sweep on the tile + 1 (c + od * (f - A u1)); the tile itself goes to global memory
    auto u1 = [&](int yy, int xx) { return U1[(yy - ty0 + 2) * F2_W + (xx - tx0 + 2)]; };
    for_region<F1_H, F1_W>([&](int ly, int lx) {
        const int y = ty0 - 1 + ly, x = tx0 - 1 + lx;
        double v = 0.0;
        if (y >= 0 && y < h && x >= 0 && x < w) {
            const double c = u1(y, x);
            v = c + g.odv(y, x) * (F[(ly + 1) * F2_W + lx + 1] - mg_au(g, u1, y, x, c, dt));
            if (ly >= 1 && ly <= FT_H && lx >= 1 && lx <= FT_W) uout[y * w + x] = v;
        }
        U2[ly * F1_W + lx] = v;
    });
    __syncthreads();
    // C: residual of the four children of every coarse cell of the tile, summed in the specified order
    auto u2 = [&](int yy, int xx) { return U2[(yy - ty0 + 1) * F1_W + (xx - tx0 + 1)]; };
    auto r = [&](int yy, int xx) { return F[(yy - ty0 + 2) * F2_W + (xx - tx0 + 2)] - mg_au(g, u2, yy, xx, u2(yy, xx), dt); };
    {
        const int i = threadIdx.x;                        // (FT_H / 2) * (FT_W / 2) == kThreads coarse cells
        const int y = ty0 + 2 * (i / (FT_W / 2)), x = tx0 + 2 * (i % (FT_W / 2));
        if (y < h && x < w) {
            const double r00 = r(y, x);
            const double r01 = x + 1 < w ? r(y, x + 1) : 0.0;
            const double r10 = y + 1 < h ? r(y + 1, x) : 0.0;
            const double r11 = (y + 1 < h && x + 1 < w) ? r(y + 1, x + 1) : 0.0;
            fc[(y >> 1) * wc + (x >> 1)] = ((r00 + r01) + r10) + r11;
        }
    }
}
// ---- regular tiles (every face of the tile + 2 open, all of it inside the grid): no bounds, no geometry, two columns per
// thread with 128-bit shared-memory accesses.  Same expressions as mg_au / the general forms, so the same bits.
// fn(ly, j) for the H rows x 34 column pairs of a (tile + 2) box: 8 rows x 32 pairs per pass, pairs 32 and 33 in one extra pass
template <int H0, int H1, class Fn>
__device__ __forceinline__ void for_pairs(const Fn &fn) {   // rows [H0, H1)
    constexpr int H = H1 - H0;
    static_assert(kThreads == 256 && 2 * H <= kThreads, "pair region shape");
    const int tid = threadIdx.x, j = tid & 31, r = tid >> 5;
#pragma unroll
    for (int k = 0; k < (H + 7) / 8; ++k) {
        const int ly = H0 + r + 8 * k;
        if (ly < H1) fn(ly, j);
    }
    if (tid < 2 * H) fn(H0 + (tid >> 1), 32 + (tid & 1));
}
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }
// one damped-Jacobi sweep at the two cells of a pair, all weights 1: c + od4 * (f - (((c-n) + (c-s)) + (c-w)) + (c-e)) * dt)
__device__ __forceinline__ double2 sweep_pair(double2 up, double2 ce, double2 dn, double l, double r, double2 f, double od4, double dt) {
    double2 v;
    v.x = ce.x + od4 * (f.x - ((((ce.x - up.x) + (ce.x - dn.x)) + (ce.x - l)) + (ce.x - ce.y)) * dt);
    v.y = ce.y + od4 * (f.y - ((((ce.y - up.y) + (ce.y - dn.y)) + (ce.y - ce.x)) + (ce.y - r)) * dt);
    return v;
}
// the sweep of box row ly, pair j, reading A (tile + 2 layout) and the right-hand side Fb; pair 0 / 33 touch a column outside
// the box (their outer element is never used; the neighbouring index stays inside the array for rows >= 1)
__device__ __forceinline__ double2 sweep_at(const double *A, const double *Fb, int ly, int j, double od4, double dt) {
    const double *p = A + ly * F2_W + 2 * j;
    return sweep_pair(ld2(p - F2_W), ld2(p), ld2(p + F2_W), p[-1], p[2], ld2(Fb + ly * F2_W + 2 * j), od4, dt);
}

__device__ __forceinline__ void mg_down_tile_regular(double od4, int w, double *__restrict__ uout, double *__restrict__ fc, int wc, double dt,
                                                     int ty0, int tx0, const double *F, double *U1, double *U2) {
    // A: u1 = od4 * f on the whole box
    for_pairs<0, F2_H>([&](int ly, int j) {
        const double2 f = ld2(F + ly * F2_W + 2 * j);
        st2(U1 + ly * F2_W + 2 * j, make_double2(od4 * f.x, od4 * f.y));
    });
    __syncthreads();
    // B: second sweep on box rows 1..18 (columns 1..66 are meaningful); the tile (rows 2..17, pairs 1..32) goes to global memory
    const bool vec = (w & 1) == 0;
    for_pairs<1, F2_H - 1>([&](int ly, int j) {
        const double2 v = sweep_at(U1, F, ly, j, od4, dt);
        st2(U2 + ly * F2_W + 2 * j, v);
        if (ly >= 2 && ly <= FT_H + 1 && j >= 1 && j <= FT_W / 2) {
            double *g = uout + (ty0 + ly - 2) * w + tx0 + 2 * j - 2;
            if (vec) st2(g, v);
            else { g[0] = v.x; g[1] = v.y; }
        }
    });
    __syncthreads();
    // C: restricted residual; coarse cell (cy, cx) = box rows 2cy+2, 2cy+3, pair cx+1
    {
        const int i = threadIdx.x, cy = i / (FT_W / 2), cx = i % (FT_W / 2), ly = 2 * cy + 2, j = cx + 1;
        const double *p = U2 + ly * F2_W + 2 * j;
        const double2 n = ld2(p - F2_W), a = ld2(p), b = ld2(p + F2_W), s = ld2(p + 2 * F2_W);
        const double al = p[-1], ar = p[2], bl = p[F2_W - 1], br = p[F2_W + 2];
        const double2 fa = ld2(F + ly * F2_W + 2 * j), fb = ld2(F + (ly + 1) * F2_W + 2 * j);
        const double r00 = fa.x - ((((a.x - n.x) + (a.x - b.x)) + (a.x - al)) + (a.x - a.y)) * dt;
        const double r01 = fa.y - ((((a.y - n.y) + (a.y - b.y)) + (a.y - a.x)) + (a.y - ar)) * dt;
        const double r10 = fb.x - ((((b.x - a.x) + (b.x - s.x)) + (b.x - bl)) + (b.x - b.y)) * dt;
        const double r11 = fb.y - ((((b.y - a.y) + (b.y - s.y)) + (b.y - b.x)) + (b.y - br)) * dt;
        fc[((ty0 >> 1) + cy) * wc + (tx0 >> 1) + cx] = ((r00 + r01) + r10) + r11;
    }
}

// compute part of k_mg_up for one staged tile: V = u, Fs = f on the tile + 2, EC = coarse correction under it
template <class G>
__device__ __forceinline__ void mg_up_tile(const G &g, int h, int w, double *__restrict__ uout, double dt, int ty0, int tx0, double *V,
                                           const double *Fs, const double *EC, double *T1) {
    // A: v = u + P e on the tile + 2, in place (the parent of box cell (ly, lx) is EC cell (ly / 2, lx / 2): the box starts at even - 2)
    for_region<F2_H, F2_W>([&](int ly, int lx) {
        const int y = ty0 - 2 + ly, x = tx0 - 2 + lx;
        if (y >= 0 && y < h && x >= 0 && x < w) V[ly * F2_W + lx] = V[ly * F2_W + lx] + EC[(ly >> 1) * EC_W + (lx >> 1)];
    });
    __syncthreads();
    auto vv = [&](int yy, int xx) { return V[(yy - ty0 + 2) * F2_W + (xx - tx0 + 2)]; };
    for_region<F1_H, F1_W>([&](int ly, int lx) {
        const int y = ty0 - 1 + ly, x = tx0 - 1 + lx;
        double t = 0.0;
        if (y >= 0 && y < h && x >= 0 && x < w) {
            const double c = vv(y, x);
            t = c + g.odv(y, x) * (Fs[(ly + 1) * F2_W + lx + 1] - mg_au(g, vv, y, x, c, dt));
        }
        T1[ly * F1_W + lx] = t;
    });
    __syncthreads();
    auto t1 = [&](int yy, int xx) { return T1[(yy - ty0 + 1) * F1_W + (xx - tx0 + 1)]; };
    for_region<FT_H, FT_W>([&](int ly, int lx) {
        const int y = ty0 + ly, x = tx0 + lx;
        if (y < h && x < w) {
            const double c = t1(y, x);
            uout[y * w + x] = c + g.odv(y, x) * (Fs[(ly + 2) * F2_W + lx + 2] - mg_au(g, t1, y, x, c, dt));
        }
    });
}
__device__ __forceinline__ void mg_up_tile_regular(double od4, int w, double *__restrict__ uout, double dt, int ty0, int tx0, double *V,
                                                   const double *Fs, const double *EC, double *T1) {
    // A: v = u + P e in place; both columns of pair j have the parent column j
    for_pairs<0, F2_H>([&](int ly, int j) {
        const double e = EC[(ly >> 1) * EC_W + j];
        double2 v = ld2(V + ly * F2_W + 2 * j);
        v.x = v.x + e;
        v.y = v.y + e;
        st2(V + ly * F2_W + 2 * j, v);
    });
    __syncthreads();
    for_pairs<1, F2_H - 1>([&](int ly, int j) { st2(T1 + ly * F2_W + 2 * j, sweep_at(V, Fs, ly, j, od4, dt)); });
    __syncthreads();
    const bool vec = (w & 1) == 0;
    {   // the tile: box rows 2..17, pairs 1..32 -- 16 x 32 pairs, two passes of 8 rows
        const int tid = threadIdx.x, j = 1 + (tid & 31), r = tid >> 5;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int ly = 2 + r + 8 * k;
            const double2 v = sweep_at(T1, Fs, ly, j, od4, dt);
            double *g = uout + (ty0 + ly - 2) * w + tx0 + 2 * j - 2;
            if (vec) st2(g, v);
            else { g[0] = v.x; g[1] = v.y; }
        }
    }
}

// is the tile at (ty0, tx0) regular (every face of the tile + 2 open)?  Level 0: from the geometry; stored levels: set-up flag
__device__ __forceinline__ bool tile_regular(const Geom0 &g, const unsigned char *, int, int ty0, int tx0) {
    return region_open(g, ty0 - 2, ty0 + FT_H + 1, tx0 - 2, tx0 + FT_W + 1);
}
__device__ __forceinline__ bool tile_regular(const GeomStored &, const unsigned char *flags, int t, int, int) { return flags[t] != 0; }

template <class G>
__global__ void __launch_bounds__(kThreads, 5)
k_mg_down(const G g, const unsigned char *__restrict__ flags, double od_reg, double *__restrict__ uout, const double *__restrict__ f,
          double *__restrict__ fc, int wc, double dt) {
    extern __shared__ __align__(16) double sm[];
    double *U1 = sm + 2 * kDownStage, *U2 = U1 + F2_H * F2_W;   // U2: tile + 1 layout (general path) or tile + 2 layout (regular path)
    const int tiles_x = (g.w + FT_W - 1) / FT_W, ntiles = tiles_x * ((g.h + FT_H - 1) / FT_H);
    const bool even = (g.w & 1) == 0;
    auto stage = [&](double *dst, int tt) {
        const int sy0 = (tt / tiles_x) * FT_H, sx0 = (tt % tiles_x) * FT_W;
        if (even && tile_regular(g, flags, tt, sy0, sx0)) stage_box2_vec(dst, f, g.w, sy0, sx0);
        else stage_box2(dst, f, g.h, g.w, sy0, sx0);
    };
    int t = blockIdx.x, st = 0;
    if (t < ntiles) stage(sm, t);
    cp_async_commit();
    for (; t < ntiles; t += gridDim.x, st ^= 1) {
        const int tn = t + gridDim.x;
        if (tn < ntiles) stage(sm + (st ^ 1) * kDownStage, tn);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const int ty0 = (t / tiles_x) * FT_H, tx0 = (t % tiles_x) * FT_W;
        const double *F = sm + st * kDownStage;
        if (tile_regular(g, flags, t, ty0, tx0)) mg_down_tile_regular(od_reg, g.w, uout, fc, wc, dt, ty0, tx0, F, U1, U2);
        else mg_down_tile(g, g.h, g.w, uout, fc, wc, dt, ty0, tx0, F, U1, U2);
        __syncthreads();   // the stage (and U1 / U2) are free again
    }
}

template <class G>
__global__ void __launch_bounds__(kThreads, 3)
k_mg_up(const G g, const unsigned char *__restrict__ flags, double od_reg, double *__restrict__ uout, const double *__restrict__ uin,
        const double *__restrict__ ec, int hc, int wc, const double *__restrict__ f, double dt) {
    extern __shared__ __align__(16) double sm[];
    double *T1 = sm + 2 * kUpStage;
    const int tiles_x = (g.w + FT_W - 1) / FT_W, ntiles = tiles_x * ((g.h + FT_H - 1) / FT_H);
    const bool even = (g.w & 1) == 0;
    auto stage = [&](double *dst, int tt) {
        const int ty0 = (tt / tiles_x) * FT_H, tx0 = (tt % tiles_x) * FT_W;
        if (even && tile_regular(g, flags, tt, ty0, tx0)) {
            stage_box2_vec(dst, uin, g.w, ty0, tx0);
            stage_box2_vec(dst + F2_H * F2_W, f, g.w, ty0, tx0);
        } else {
            stage_box2(dst, uin, g.h, g.w, ty0, tx0);
            stage_box2(dst + F2_H * F2_W, f, g.h, g.w, ty0, tx0);
        }
        const int cy0 = (ty0 >> 1) - 1, cx0 = (tx0 >> 1) - 1;
        double *E = dst + 2 * F2_H * F2_W;
        for (int i = threadIdx.x; i < EC_H * EC_W; i += kThreads) {
            const int y = cy0 + i / EC_W, x = cx0 + i % EC_W;
            const bool ok = y >= 0 && y < hc && x >= 0 && x < wc;
            cp_async8(E + i, ok ? ec + y * wc + x : ec, ok);
        }
    };
    int t = blockIdx.x, st = 0;
    if (t < ntiles) stage(sm, t);
    cp_async_commit();
    for (; t < ntiles; t += gridDim.x, st ^= 1) {
        const int tn = t + gridDim.x;
        if (tn < ntiles) stage(sm + (st ^ 1) * kUpStage, tn);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const int ty0 = (t / tiles_x) * FT_H, tx0 = (t % tiles_x) * FT_W;
        double *V = sm + st * kUpStage;
        if (tile_regular(g, flags, t, ty0, tx0)) mg_up_tile_regular(od_reg, g.w, uout, dt, ty0, tx0, V, V + F2_H * F2_W, V + 2 * F2_H * F2_W, T1);
        else mg_up_tile(g, g.h, g.w, uout, dt, ty0, tx0, V, V + F2_H * F2_W, V + 2 * F2_H * F2_W, T1);
        __syncthreads();
    }
}

// set-up: flags[t] = 1 when every face of tile t + 2 of a stored level has weight exactly 1 (then od = omega / (dt * 4) there)
__global__ void k_mg_tile_flags(const GeomStored g, unsigned char *__restrict__ flags) {
    const int tiles_x = (g.w + FT_W - 1) / FT_W;
    const int t = blockIdx.x, ty0 = (t / tiles_x) * FT_H, tx0 = (t % tiles_x) * FT_W;
    const bool outside = ty0 - 2 < 0 || ty0 + FT_H + 1 > g.h - 1 || tx0 - 2 < 0 || tx0 + FT_W + 1 > g.w - 1;   // block-uniform
    int b = 0;
    if (!outside) {
        for (int i = threadIdx.x; i < F2_H * F2_W; i += blockDim.x) {
            const int y = ty0 - 2 + i / F2_W, x = tx0 - 2 + i % F2_W;
            double n, s, w_, e;
            g.weights(y, x, n, s, w_, e);
            if (n != 1.0 || s != 1.0 || w_ != 1.0 || e != 1.0) b = 1;
        }
    }
    const int any = __syncthreads_or(b);
    if (threadIdx.x == 0) flags[t] = (outside || any) ? 0 : 1;
}

// ---- every level that fits kTailMax^2 cells: one CTA, one thread per cell, block barriers between the operations
struct TailArgs {
    int nlev;
    GeomStored g[kTailLevels];
    double *u[kTailLevels], *t[kTailLevels];
    const double *f0;            // rhs of the first tail level (the caller's src when the tail starts at level 0)
    double *f[kTailLevels];      // rhs of the deeper levels (f[0] unused)
    double dt;
};

__global__ void __launch_bounds__(kTailMax *kTailMax) k_mg_tail(const TailArgs a) {
    const int t = threadIdx.x;
    const int last = a.nlev - 1;
    for (int l = 0; l < last; ++l) {
        const GeomStored &g = a.g[l];
        const double *f = l == 0 ? a.f0 : a.f[l];
        if (t < g.h * g.w) a.u[l][t] = mg_pre_cell(g, f, t / g.w, t % g.w, a.dt);
        __syncthreads();
        const GeomStored &gc = a.g[l + 1];
        if (t < gc.h * gc.w) a.f[l + 1][t] = mg_restrict_cell(g, a.u[l], f, t / gc.w, t % gc.w, a.dt);
        __syncthreads();
    }
    {
        const GeomStored &g = a.g[last];
        const double *f = last == 0 ? a.f0 : a.f[last];
        const bool on = t < g.h * g.w;
        const int y = on ? t / g.w : 0, x = on ? t % g.w : 0;
        if (on) a.u[last][t] = mg_pre_cell(g, f, y, x, a.dt);
        __syncthreads();
        for (int k = 2; k < kCoarseSweeps; k += 2) {
            if (on) a.t[last][t] = mg_sweep_cell(g, a.u[last], f, y, x, a.dt);
            __syncthreads();
            if (on) a.u[last][t] = mg_sweep_cell(g, a.t[last], f, y, x, a.dt);
            __syncthreads();
        }
    }
    for (int l = last - 1; l >= 0; --l) {
        const GeomStored &g = a.g[l];
        const double *f = l == 0 ? a.f0 : a.f[l];
        const bool on = t < g.h * g.w;
        const int y = on ? t / g.w : 0, x = on ? t % g.w : 0;
        if (on) a.t[l][t] = mg_post1_cell(g, a.u[l], a.u[l + 1], a.g[l + 1].w, f, y, x, a.dt);
        __syncthreads();
        if (on) a.u[l][t] = mg_sweep_cell(g, a.t[l], f, y, x, a.dt);
        __syncthreads();
    }
}

// ---- setup: coarse face weights (halve-and-add of the two fine faces behind a coarse face) and od
template <class G>
__global__ void k_mg_coarsen(const G g, double *__restrict__ wyc, double *__restrict__ wxc, int hc, int wc) {
    const int n_y = (hc + 1) * wc, n_x = hc * (wc + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_y + n_x; i += gridDim.x * blockDim.x) {
        if (i < n_y) {
            const int Y = i / wc, X = i % wc, y = 2 * Y, x = 2 * X;
            const double a = y <= g.h ? g.face_y(y, x) : 0.0;
            const double b = (y <= g.h && x + 1 < g.w) ? g.face_y(y, x + 1) : 0.0;
            wyc[i] = 0.5 * (a + b);
        } else {
            const int j = i - n_y, Y = j / (wc + 1), X = j % (wc + 1), y = 2 * Y, x = 2 * X;
            const double a = x <= g.w ? g.face_x(y, x) : 0.0;
            const double b = (x <= g.w && y + 1 < g.h) ? g.face_x(y + 1, x) : 0.0;
            wxc[j] = 0.5 * (a + b);
        }
    }
}
__global__ void k_mg_store_level0(const Geom0 g, double *__restrict__ wy, double *__restrict__ wx) {
    const int n_y = (g.h + 1) * g.w, n_x = g.h * (g.w + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_y + n_x; i += gridDim.x * blockDim.x) {
        if (i < n_y) wy[i] = g.face_y(i / g.w, i % g.w);
        else wx[i - n_y] = g.face_x((i - n_y) / (g.w + 1), (i - n_y) % (g.w + 1));
    }
}
__global__ void k_mg_od(const double *__restrict__ wy, const double *__restrict__ wx, double *__restrict__ od, int h, int w, double dt) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h * w; i += gridDim.x * blockDim.x) {
        const int y = i / w, x = i % w;
        const double d = ((wy[y * w + x] + wy[(y + 1) * w + x]) + wx[y * (w + 1) + x]) + wx[y * (w + 1) + x + 1];
        od[i] = d > 0.0 ? kOmega / (dt * d) : 0.0;
    }
}

__global__ void __launch_bounds__(kThreads)
k_jacobi(const Geom0 g, double *__restrict__ dst, const double *__restrict__ src) {   // g.od holds 1 / (dt * k) here
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    if (x >= g.w) return;
    for (int y = blockIdx.y * 8 + (threadIdx.x >> 5); y < g.h; y += gridDim.y * 8) {
        const double inv = g.odv(y, x);
        dst[y * g.w + x] = inv != 0.0 ? src[y * g.w + x] * inv : 0.0;
    }
}

inline dim3 grid2d(int rows, int cols) {
    int gy = (rows + 7) / 8;
    if (gy > 16384) gy = 16384;
    return dim3((unsigned)((cols + 31) / 32), (unsigned)(gy < 1 ? 1 : gy));
}

}  // namespace

struct pano_mg {
    pano_ctx *ctx = nullptr;
    size_t h = 0, w = 0;
    double dt = 0;
    pano_rect obstacle{0, 0, 0, 0};
    int nlev = 0, tail = 0;                 // levels [tail, nlev) run in k_mg_tail
    int hs[kMaxLevels], ws[kMaxLevels];
    double *wy[kMaxLevels], *wx[kMaxLevels], *od[kMaxLevels], *u[kMaxLevels], *f[kMaxLevels], *t[kMaxLevels];
    double *pool = nullptr;
    unsigned char *flags[kMaxLevels];       // per-tile `regular` flags of the stored levels (null for level 0)
    unsigned char *flag_pool = nullptr;
    double od_reg = 0;                      // omega / (dt * 4)
    int occ_down = 1, occ_up = 1;           // resident blocks per SM of k_mg_down / k_mg_up (occupancy query)
    Geom0 g0;
};

static GeomStored stored(const pano_mg *m, int l) { return GeomStored{m->hs[l], m->ws[l], m->wy[l], m->wx[l], m->od[l]}; }

static int mg_build(pano_mg *m) {
    pano_ctx *ctx = m->ctx;
    // level dimensions
    int l = 0;
    m->hs[0] = (int)m->h; m->ws[0] = (int)m->w;
    while ((m->hs[l] > m->ws[l] ? m->hs[l] : m->ws[l]) > kCoarsestMax && l + 1 < kMaxLevels) {
        m->hs[l + 1] = (m->hs[l] + 1) / 2;
        m->ws[l + 1] = (m->ws[l] + 1) / 2;
        ++l;
    }
    m->nlev = l + 1;
    m->tail = m->nlev - 1;
    while (m->tail > 0 && m->hs[m->tail - 1] <= kTailMax && m->ws[m->tail - 1] <= kTailMax) --m->tail;
    if (m->nlev - m->tail > kTailLevels) m->tail = m->nlev - kTailLevels;
    RectI mr = pano_clip_rect(m->obstacle, m->h + 1, m->w + 1);
    m->g0.h = (int)m->h; m->g0.w = (int)m->w; m->g0.m = mr;
    m->g0.od[0] = 0.0;
    for (int k = 1; k <= 4; ++k) m->g0.od[k] = kOmega / (m->dt * (double)k);
    // one pool: stored geometry for levels >= 1 (and level 0 when the tail starts there), u/f/t for levels >= 1, t for level 0
    size_t total = 0;
    auto take = [&](size_t n) { size_t o = total; total += (n + 1) & ~(size_t)1; return o; };   // keep 16-byte alignment
    size_t o_wy[kMaxLevels], o_wx[kMaxLevels], o_od[kMaxLevels], o_u[kMaxLevels], o_f[kMaxLevels], o_t[kMaxLevels];
    for (int i = 0; i < m->nlev; ++i) {
        const size_t hh = m->hs[i], ww = m->ws[i];
        const bool geom = i > 0 || m->tail == 0;
        o_wy[i] = geom ? take((hh + 1) * ww) : 0;
        o_wx[i] = geom ? take(hh * (ww + 1)) : 0;
        o_od[i] = geom ? take(hh * ww) : 0;
        o_u[i] = i > 0 ? take(hh * ww) : 0;
        o_f[i] = i > 0 ? take(hh * ww) : 0;
        o_t[i] = take(hh * ww);
    }
    PANO_CUDA(cudaMalloc((void **)&m->pool, total * sizeof(double)));
    PANO_CUDA(cudaMemsetAsync(m->pool, 0, total * sizeof(double), ctx->stream));
    for (int i = 0; i < m->nlev; ++i) {
        const bool geom = i > 0 || m->tail == 0;
        m->wy[i] = geom ? m->pool + o_wy[i] : nullptr;
        m->wx[i] = geom ? m->pool + o_wx[i] : nullptr;
        m->od[i] = geom ? m->pool + o_od[i] : nullptr;
        m->u[i] = i > 0 ? m->pool + o_u[i] : nullptr;
        m->f[i] = i > 0 ? m->pool + o_f[i] : nullptr;
        m->t[i] = m->pool + o_t[i];
    }
    m->od_reg = kOmega / (m->dt * 4.0);
    auto ntiles = [&](int i) { return (size_t)((m->ws[i] + FT_W - 1) / FT_W) * (size_t)((m->hs[i] + FT_H - 1) / FT_H); };
    size_t nflags = 0;
    for (int i = 1; i < m->tail; ++i) nflags += ntiles(i);
    for (int i = 0; i < kMaxLevels; ++i) m->flags[i] = nullptr;
    if (nflags) {
        PANO_CUDA(cudaMalloc((void **)&m->flag_pool, nflags));
        size_t o = 0;
        for (int i = 1; i < m->tail; ++i) { m->flags[i] = m->flag_pool + o; o += ntiles(i); }
    }
    auto flat = [&](size_t n) { size_t b = (n + 255) / 256; return (int)(b < 1 ? 1 : (b > 4096 ? 4096 : b)); };
    if (m->tail == 0) {
        k_mg_store_level0<<<flat((m->h + 1) * (m->w + 1) * 2), 256, 0, ctx->stream>>>(m->g0, m->wy[0], m->wx[0]);
        PANO_TRY(pano_after_launch(ctx, "mg_store_level0"));
        k_mg_od<<<flat(m->h * m->w), 256, 0, ctx->stream>>>(m->wy[0], m->wx[0], m->od[0], m->hs[0], m->ws[0], m->dt);
        PANO_TRY(pano_after_launch(ctx, "mg_od"));
    }
    for (int i = 0; i + 1 < m->nlev; ++i) {
        const int hc = m->hs[i + 1], wc = m->ws[i + 1];
        const int g = flat((size_t)(hc + 1) * (wc + 1) * 2);
        if (i == 0 && m->tail != 0) k_mg_coarsen<Geom0><<<g, 256, 0, ctx->stream>>>(m->g0, m->wy[1], m->wx[1], hc, wc);
        else k_mg_coarsen<GeomStored><<<g, 256, 0, ctx->stream>>>(stored(m, i), m->wy[i + 1], m->wx[i + 1], hc, wc);
        PANO_TRY(pano_after_launch(ctx, "mg_coarsen"));
        k_mg_od<<<flat((size_t)hc * wc), 256, 0, ctx->stream>>>(m->wy[i + 1], m->wx[i + 1], m->od[i + 1], hc, wc, m->dt);
        PANO_TRY(pano_after_launch(ctx, "mg_od"));
    }
    for (int i = 1; i < m->tail; ++i) {
        k_mg_tile_flags<<<(unsigned)ntiles(i), 128, 0, ctx->stream>>>(stored(m, i), m->flags[i]);
        PANO_TRY(pano_after_launch(ctx, "mg_tile_flags"));
    }
    PANO_CUDA(cudaFuncSetAttribute(k_mg_down<Geom0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDownSmem));
    PANO_CUDA(cudaFuncSetAttribute(k_mg_down<GeomStored>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDownSmem));
    PANO_CUDA(cudaFuncSetAttribute(k_mg_up<Geom0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmem));
    PANO_CUDA(cudaFuncSetAttribute(k_mg_up<GeomStored>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmem));
    int a = 0, b = 0;
    PANO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_mg_down<Geom0>, kThreads, kDownSmem));
    PANO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_mg_down<GeomStored>, kThreads, kDownSmem));
    m->occ_down = a < b ? a : b;
    PANO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_mg_up<Geom0>, kThreads, kUpSmem));
    PANO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_mg_up<GeomStored>, kThreads, kUpSmem));
    m->occ_up = a < b ? a : b;
    if (m->occ_down < 1) m->occ_down = 1;
    if (m->occ_up < 1) m->occ_up = 1;
    return PANO_OK;
}

template <class G>
static int mg_level_op(pano_ctx *ctx, const G &g, int op, double *out, const double *uin, const double *f, const double *ec, int hc, int wc,
                       double dt) {
    const dim3 grid = op == OP_RESTRICT ? grid2d(hc, wc) : grid2d(g.h, g.w);
    switch (op) {
        case OP_PRE: k_mg_level<G, OP_PRE><<<grid, kThreads, 0, ctx->stream>>>(g, out, uin, f, ec, hc, wc, dt); break;
        case OP_SWEEP: k_mg_level<G, OP_SWEEP><<<grid, kThreads, 0, ctx->stream>>>(g, out, uin, f, ec, hc, wc, dt); break;
        case OP_POST1: k_mg_level<G, OP_POST1><<<grid, kThreads, 0, ctx->stream>>>(g, out, uin, f, ec, hc, wc, dt); break;
        default: k_mg_level<G, OP_RESTRICT><<<grid, kThreads, 0, ctx->stream>>>(g, out, uin, f, ec, hc, wc, dt); break;
    }
    return pano_after_launch(ctx, "mg_level");
}

// dst = V-cycle(src) on raw device pointers (h*w doubles each, distinct)
int pano_mg_apply_raw(pano_mg *m, double *dst, const double *src) {
    pano_ctx *ctx = m->ctx;
    auto U = [&](int l) { return l == 0 ? dst : m->u[l]; };
    auto F = [&](int l) { return l == 0 ? src : (const double *)m->f[l]; };
    auto op = [&](int l, int o, double *out, const double *uin, const double *ec) -> int {
        const int hc = l + 1 < m->nlev ? m->hs[l + 1] : 0, wc = l + 1 < m->nlev ? m->ws[l + 1] : 0;
        if (l == 0) return mg_level_op(ctx, m->g0, o, out, uin, F(0), ec, hc, wc, m->dt);
        return mg_level_op(ctx, stored(m, l), o, out, uin, F(l), ec, hc, wc, m->dt);
    };
    const bool fused = pano_option(ctx, "mg_fused", 1) != 0;
    auto nblocks = [&](int l, int per_sm) {   // persistent blocks: a few per SM, never more than there are tiles
        const size_t nt = (size_t)((m->ws[l] + FT_W - 1) / FT_W) * (size_t)((m->hs[l] + FT_H - 1) / FT_H);
        const size_t cap = (size_t)ctx->num_sms * per_sm;
        return (unsigned)(nt < cap ? nt : cap);
    };
    // persistent kernels: exactly as many blocks as are resident at once (more would run as a second, nearly empty wave)
    int down_per_sm = (int)pano_option(ctx, "mg_down_blocks", 0), up_per_sm = (int)pano_option(ctx, "mg_up_blocks", 0);
    if (down_per_sm <= 0) down_per_sm = m->occ_down;
    if (up_per_sm <= 0) up_per_sm = m->occ_up;
    for (int l = 0; l < m->tail; ++l) {                         // down: two sweeps from zero, residual + restriction
        if (fused) {                                            // one pass: f -> t (= u), f'
            if (l == 0) k_mg_down<Geom0><<<nblocks(0, down_per_sm), kThreads, kDownSmem, ctx->stream>>>(m->g0, nullptr, m->od_reg, m->t[0], F(0), m->f[1], m->ws[1], m->dt);
            else k_mg_down<GeomStored><<<nblocks(l, down_per_sm), kThreads, kDownSmem, ctx->stream>>>(stored(m, l), m->flags[l], m->od_reg, m->t[l], F(l), m->f[l + 1], m->ws[l + 1], m->dt);
            PANO_TRY(pano_after_launch(ctx, "mg_down"));
            continue;
        }
        PANO_TRY(op(l, OP_PRE, U(l), nullptr, nullptr));
        PANO_TRY(op(l, OP_RESTRICT, m->f[l + 1], U(l), nullptr));
    }
    TailArgs a;
    memset(&a, 0, sizeof(a));
    a.nlev = m->nlev - m->tail;
    a.dt = m->dt;
    a.f0 = F(m->tail);
    for (int i = 0; i < a.nlev; ++i) {
        const int l = m->tail + i;
        a.g[i] = stored(m, l);
        a.u[i] = U(l);
        a.t[i] = m->t[l];
        a.f[i] = m->f[l];
    }
    k_mg_tail<<<1, kTailMax * kTailMax, 0, ctx->stream>>>(a);
    PANO_TRY(pano_after_launch(ctx, "mg_tail"));
    for (int l = m->tail - 1; l >= 0; --l) {                    // up: prolongation fused into post-sweep 1, post-sweep 2
        if (fused) {                                            // one pass: t, u', f -> u
            if (l == 0) k_mg_up<Geom0><<<nblocks(0, up_per_sm), kThreads, kUpSmem, ctx->stream>>>(m->g0, nullptr, m->od_reg, U(0), m->t[0], U(1), m->hs[1], m->ws[1], F(0), m->dt);
            else k_mg_up<GeomStored><<<nblocks(l, up_per_sm), kThreads, kUpSmem, ctx->stream>>>(stored(m, l), m->flags[l], m->od_reg, U(l), m->t[l], U(l + 1), m->hs[l + 1], m->ws[l + 1], F(l), m->dt);
            PANO_TRY(pano_after_launch(ctx, "mg_up"));
            continue;
        }
        PANO_TRY(op(l, OP_POST1, m->t[l], U(l), U(l + 1)));
        PANO_TRY(op(l, OP_SWEEP, U(l), m->t[l], nullptr));
    }
    return PANO_OK;
}

static pano_mg *mg_cached(pano_ctx *ctx, size_t h, size_t w, double dt, pano_rect ob) {
    for (pano_mg *m : ctx->mg_cache)
        if (m->h == h && m->w == w && m->dt == dt && m->obstacle.y0 == ob.y0 && m->obstacle.y1 == ob.y1 && m->obstacle.x0 == ob.x0 &&
            m->obstacle.x1 == ob.x1)
            return m;
    return nullptr;
}

void pano_mg_free_all(pano_ctx *ctx) {
    for (pano_mg *m : ctx->mg_cache) {
        if (m->pool) cudaFree(m->pool);
        if (m->flag_pool) cudaFree(m->flag_pool);
        delete m;
    }
    ctx->mg_cache.clear();
}

static int jacobi_raw(pano_ctx *ctx, double *dst, const double *src, size_t h, size_t w, double dt, pano_rect ob) {
    Geom0 g;
    g.h = (int)h; g.w = (int)w;
    g.m = pano_clip_rect(ob, h + 1, w + 1);
    g.od[0] = 0.0;
    for (int k = 1; k <= 4; ++k) g.od[k] = 1.0 / (dt * (double)k);
    k_jacobi<<<grid2d((int)h, (int)w), kThreads, 0, ctx->stream>>>(g, dst, src);
    return pano_after_launch(ctx, "jacobi");
}

// pcg.rs:14-82 with a Preconditioner object, host-driven: every scalar is a device reduction read back, exactly like the
// reference's sigma / alpha / beta / residual_error live on its host.  The identity goes through the persistent kernels instead.
int pano_pcg_precond_raw(pano_ctx *ctx, int precond, pano_field *x, const pano_field *b, int max_iterations, double threshold,
                         pano_field *residual, pano_field *auxiliary, pano_field *search, double dt, pano_rect ob, pano_pcg_info *info) {
    if (x->dtype != PANO_F64) PANO_FAIL(PANO_ERR_UNIMPLEMENTED, "pano_pcg_solve: the Jacobi / multigrid preconditioners are f64 only");
    if (x->n >= ((size_t)1 << 31)) PANO_FAIL(PANO_ERR_SHAPE, "pano_pcg_solve: grid too large for 32-bit cell indices");
    pano_mg *mg = nullptr;
    if (precond == PANO_PRECOND_MULTIGRID) {
        mg = mg_cached(ctx, x->h, x->w, dt, ob);
        if (!mg) PANO_TRY(pano_mg_create(ctx, x->h, x->w, dt, ob, &mg));
    }
    auto apply = [&](pano_field *dst, const pano_field *src) -> int {
        if (mg) return pano_mg_apply_raw(mg, (double *)dst->d, (const double *)src->d);
        return jacobi_raw(ctx, (double *)dst->d, (const double *)src->d, x->h, x->w, dt, ob);
    };
    double bmax = 0, sigma = 0, zs = 0, err = 0;
    PANO_TRY(pano_field_fill(x, 0.0));                                          // :32
    PANO_TRY(pano_field_norm_max(b, &bmax));
    if (bmax < threshold) {                                                      // :35-38
        if (info) { info->iterations = -1; info->applies = 0; info->final_residual = bmax; info->rhs_max = bmax; }
        return PANO_OK;
    }
    PANO_TRY(pano_field_assign(residual, b));                                    // :40
    PANO_TRY(apply(auxiliary, residual));                                        // :41
    PANO_TRY(pano_field_assign(search, auxiliary));                              // :42
    PANO_TRY(pano_field_dot(auxiliary, residual, &sigma));                       // :46
    int it = max_iterations, applies = 0;
    err = bmax;
    for (int i = 0; i < max_iterations; ++i) {                                   // :48
        PANO_TRY(pano_laplacian_apply(auxiliary, search, dt, ob));               // :51
        ++applies;
        PANO_TRY(pano_field_dot(auxiliary, search, &zs));
        const double alpha = sigma / zs;                                         // :53
        PANO_TRY(pano_field_scaled_add(x, alpha, search));                       // :55
        PANO_TRY(pano_field_scaled_add(residual, -alpha, auxiliary));            // :56
        PANO_TRY(pano_field_norm_max(residual, &err));                           // :58
        if (err < threshold) { it = i; break; }                                  // :60-63
        PANO_TRY(apply(auxiliary, residual));                                    // :65
        double sigma_new = 0;
        PANO_TRY(pano_field_dot(auxiliary, residual, &sigma_new));               // :67
        const double beta = sigma_new / sigma;                                   // :68
        PANO_TRY(pano_field_xpby(search, auxiliary, beta));                      // :72-77
        sigma = sigma_new;                                                       // :79
    }
    if (info) { info->iterations = it; info->applies = applies; info->final_residual = err; info->rhs_max = bmax; }
    return PANO_OK;
}

extern "C" {

int pano_mg_create(pano_ctx *ctx, size_t h, size_t w, double timestep, pano_rect obstacle, pano_mg **out) {
    if (!ctx || !out) PANO_FAIL(PANO_ERR_INVALID, "pano_mg_create: null argument");
    if (h < 1 || w < 1) PANO_FAIL(PANO_ERR_SHAPE, "pano_mg_create: empty grid");
    if ((h + 1) * (w + 1) >= ((size_t)1 << 31)) PANO_FAIL(PANO_ERR_SHAPE, "pano_mg_create: grid too large for 32-bit cell indices");
    if (!(timestep > 0.0)) PANO_FAIL(PANO_ERR_INVALID, "pano_mg_create: timestep must be positive");
    PANO_TRY(pano_check_rect_within(obstacle, h, w, "pano_mg_create(obstacle)"));
    PANO_TRY(pano_activate(ctx));
    pano_mg *m = new pano_mg();
    m->ctx = ctx; m->h = h; m->w = w; m->dt = timestep; m->obstacle = obstacle;
    const int rc = mg_build(m);
    if (rc != PANO_OK) {
        if (m->pool) cudaFree(m->pool);
        if (m->flag_pool) cudaFree(m->flag_pool);
        delete m;
        return rc;
    }
    ctx->mg_cache.push_back(m);   // owned by the context (freed by pano_mg_destroy or with the context)
    *out = m;
    return PANO_OK;
}

int pano_mg_destroy(pano_mg *m) {
    if (!m) return PANO_OK;
    pano_ctx *ctx = m->ctx;
    PANO_TRY(pano_activate(ctx));
    PANO_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < ctx->mg_cache.size(); ++i)
        if (ctx->mg_cache[i] == m) { ctx->mg_cache.erase(ctx->mg_cache.begin() + i); break; }
    if (m->pool) cudaFree(m->pool);
    if (m->flag_pool) cudaFree(m->flag_pool);
    delete m;
    return PANO_OK;
}

int pano_mg_levels(const pano_mg *m, int *levels, int *tail_levels) {
    if (!m || !levels) PANO_FAIL(PANO_ERR_INVALID, "pano_mg_levels: null argument");
    *levels = m->nlev;
    if (tail_levels) *tail_levels = m->nlev - m->tail;
    return PANO_OK;
}

int pano_mg_apply(pano_mg *m, pano_field *dst, const pano_field *src) {
    if (!m) PANO_FAIL(PANO_ERR_INVALID, "pano_mg_apply: null preconditioner");
    PANO_TRY(pano_check_kind(dst, PANO_SIMPLEX2, "pano_mg_apply(dst)"));
    PANO_TRY(pano_check_kind(src, PANO_SIMPLEX2, "pano_mg_apply(src)"));
    PANO_TRY(pano_check_same(dst, src, "pano_mg_apply"));
    if (dst->ctx != m->ctx || dst->h != m->h || dst->w != m->w) PANO_FAIL(PANO_ERR_SHAPE, "pano_mg_apply: field does not match the preconditioner's grid");
    if (dst->dtype != PANO_F64) PANO_FAIL(PANO_ERR_UNIMPLEMENTED, "pano_mg_apply: f64 only");
    if (dst->d == src->d) PANO_FAIL(PANO_ERR_INVALID, "pano_mg_apply: dst aliases src (the trait takes &mut dst, &src)");
    PANO_TRY(pano_activate(m->ctx));
    return pano_mg_apply_raw(m, (double *)dst->d, (const double *)src->d);
}

int pano_jacobi_apply(pano_field *dst, const pano_field *src, double timestep, pano_rect obstacle) {
    PANO_TRY(pano_check_kind(dst, PANO_SIMPLEX2, "pano_jacobi_apply(dst)"));
    PANO_TRY(pano_check_kind(src, PANO_SIMPLEX2, "pano_jacobi_apply(src)"));
    PANO_TRY(pano_check_same(dst, src, "pano_jacobi_apply"));
    if (dst->dtype != PANO_F64) PANO_FAIL(PANO_ERR_UNIMPLEMENTED, "pano_jacobi_apply: f64 only");
    if (dst->d == src->d) PANO_FAIL(PANO_ERR_INVALID, "pano_jacobi_apply: dst aliases src");
    if (dst->n >= ((size_t)1 << 31)) PANO_FAIL(PANO_ERR_SHAPE, "pano_jacobi_apply: grid too large for 32-bit cell indices");
    PANO_TRY(pano_check_rect_within(obstacle, dst->h, dst->w, "pano_jacobi_apply(obstacle)"));
    PANO_TRY(pano_activate(dst->ctx));
    if (dst->n == 0) return PANO_OK;
    return jacobi_raw(dst->ctx, (double *)dst->d, (const double *)src->d, dst->h, dst->w, timestep, obstacle);
}

}  // extern "C"
