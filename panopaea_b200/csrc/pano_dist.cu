// pano_dist.cu -- the grid fluid step slab-decomposed over the GPUs of one node.
//
// One process per GPU.  Rank g owns cell rows [y0, y1) of the H x W grid, the vx rows and the vy
// face rows with the same indices (the last rank also owns face row H).  Every array carries
// kGhost ghost rows on both sides, and every array of a rank lives in ONE allocation (the
// "window") that the other ranks map through CUDA IPC (or, in single-process loop-back tests,
// address directly).  Nothing crosses the host:
//   * the four halo exchanges of a step outside the solver (advection inputs, vy after advection,
//     the right-hand side, the pressure) are a copy kernel that stores my boundary rows straight
//     into the neighbours' ghost rows over NVLink and then raises a {sequence} flag there, plus a
//     one-thread wait kernel in the consumer's stream;
//   * inside the solver the streaming CG kernel (pano_cg_stream.cu) stores its first/last row of
//     s' and r into the neighbours' ghost rows as it produces them, and the two reductions per
//     iteration get a cross-rank stage over peer-mapped units (pano_sm100.cuh).
// NCCL is not in the data path; torch.distributed is used by the Python harness only to hand the
// IPC handles around.  The step itself is pano_step.cu's, restricted to the slab:
// examples/dec_fluid.rs:46-141.
#include <cstdlib>
#include <ctime>

#include "pano_sm100.cuh"

using namespace pano_sm100;

int pano_advect_slab_launch(pano_ctx *ctx, double *q_dst, double *vy_dst, double *vx_dst, const double *q_src, const double *vy_src,
                            const double *vx_src, size_t h, size_t w, double dt, int ya, int yb, int ylo, size_t rows_q, unsigned int *err);
int pano_neg_divergence_slab_launch(pano_ctx *ctx, double *b, const double *vy, const double *vx, size_t h, size_t w, pano_rect obstacle,
                                    int ya, int yb);
int pano_project_slab_launch(pano_ctx *ctx, double *vy, double *vx, const double *p, size_t h, size_t w, double dt, int ya, int yb);
int pano_cg_stream_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, double *s1, size_t h, size_t w,
                          int max_iterations, double threshold, double timestep, RectI m, const PanoCgSlab *slab);

int pano_cg_sr_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *p, double *s0, double *r1, double *s1, size_t h,
                      size_t w, int max_iterations, double threshold, double timestep, RectI m, const PanoCgSrSlab *slab);
int pano_preload_cg_sr();

int pano_preload_fused();
int pano_preload_cg_stream();

namespace {

constexpr int kGhost = 12;         // ghost rows per side; the backtrace reach dt*max|v| + 2 (+ kExtend in the fused-halo step) must fit (checked on device)
constexpr int kExtend = 4;         // fused-halo step: rows beyond the slab that a rank advects itself instead of receiving them (even: vx rows pair up)
constexpr int kThreads = 256;
enum { EX_ADV = 0, EX_VY = 1, EX_B = 2, EX_P = 3, EX_COUNT = 4 };
// F_S0 is the search direction (p of the single-reduction kernel), F_S1 / F_S2 its s = A p buffers, F_R / F_R1 the residual's
enum { F_D0 = 0, F_D1, F_VY0, F_VY1, F_VX0, F_VX1, F_P, F_R, F_S0, F_S1, F_R1, F_S2, F_COUNT };

// element offsets (in doubles) of every array inside a rank's window; the same formula on every rank
struct Layout {
    size_t y0, y1, hl;
    size_t off[F_COUNT], pitch[F_COUNT], rows[F_COUNT];
    size_t xunits, flags, err;
    size_t total;
};

void slab_range(size_t H, int rank, int nranks, size_t *y0, size_t *y1) {
    *y0 = H * (size_t)rank / (size_t)nranks;
    *y1 = H * (size_t)(rank + 1) / (size_t)nranks;
}

Layout make_layout(size_t H, size_t W, int rank, int nranks) {
    Layout L;
    slab_range(H, rank, nranks, &L.y0, &L.y1);
    L.hl = L.y1 - L.y0;
    size_t off = 0;
    auto take = [&](size_t n) {
        size_t o = off;
        off += (n + 31) & ~(size_t)31;   // 256-byte granules keep every array TMA / vector aligned
        return o;
    };
    for (int f = 0; f < F_COUNT; ++f) {
        const bool is_vy = f == F_VY0 || f == F_VY1, is_vx = f == F_VX0 || f == F_VX1;
        L.pitch[f] = is_vx ? W + 1 : W;
        L.rows[f] = L.hl + 2 * kGhost + (is_vy ? 1 : 0);
        L.off[f] = take(L.rows[f] * L.pitch[f]);
    }
    L.xunits = take(2 * (size_t)(kXUnitsTotal + kXFlagUnits));   // ReduceUnit = 2 doubles; cross-rank totals, then the halo flags
    L.flags = take(2 * (size_t)EX_COUNT * 2);
    L.err = take(2);
    L.total = off;
    return L;
}

struct Segments {                  // row blocks copied by one push launch
    const double *src[8];
    double *dst[8];
    size_t n[8];
    int count;
};

// copy my boundary rows into the neighbours' ghost rows, then (last block) raise their flags
__global__ void __launch_bounds__(kThreads) k_push(Segments seg, ReduceUnit *flag_up, ReduceUnit *flag_dn, unsigned long long seq,
                                                   unsigned int *counter) {
    for (int s = 0; s < seg.count; ++s) {
        const double *src = seg.src[s];
        double *dst = seg.dst[s];
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < seg.n[s]; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {
            *counter = 0;
            __threadfence_system();
            if (flag_up) unit_store(flag_up, 0.0, seq);
            if (flag_dn) unit_store(flag_dn, 0.0, seq);
        }
    }
}

__global__ void k_wait(const ReduceUnit *f0, const ReduceUnit *f1, unsigned long long seq, unsigned int *err) {
    double v;
    if (f0 && !unit_poll(f0, seq, v, err)) return;
    if (f1 && !unit_poll(f1, seq, v, err)) return;
    __threadfence_system();
}

__global__ void k_fill_rows(double *p, size_t pitch, int y0, int y1, int x0, int x1, double v) {
    const int rw = x1 - x0;
    const size_t n = (size_t)(y1 - y0) * rw;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[(size_t)(y0 + (int)(i / rw)) * pitch + x0 + (int)(i % rw)] = v;
}

}  // namespace

struct pano_dist {
    pano_ctx *ctx = nullptr;
    int rank = 0, nranks = 1;
    size_t H = 0, W = 0;
    pano_step_params prm;
    Layout L;
    double *window = nullptr;
    double *peer[kMaxRanks] = {nullptr};
    bool peer_ipc[kMaxRanks] = {false};
    bool connected = false;
    int cur = 0;                      // which ping-pong buffer holds the current density / velocity
    unsigned long long step_no = 0;   // identical on all ranks
    unsigned long long solve_no = 0;  // CG launches so far (steps and stand-alone solves), identical on all ranks
    unsigned int *d_counter = nullptr;
    int max_ctas = 0;
};

namespace {

// virtual address of GLOBAL row 0 of local array f (only rows [y0 - kGhost, ...) are stored)
inline double *virt(const pano_dist *d, int f) {
    return d->window + d->L.off[f] - ((ptrdiff_t)d->L.y0 - kGhost) * (ptrdiff_t)d->L.pitch[f];
}
// address of global row g of array f inside rank r's window, through the peer mapping
inline double *peer_row(const pano_dist *d, int r, const Layout &Lr, int f, ptrdiff_t g) {
    return d->peer[r] + Lr.off[f] + (g - ((ptrdiff_t)Lr.y0 - kGhost)) * (ptrdiff_t)Lr.pitch[f];
}
inline ReduceUnit *flag_of(const pano_dist *d, int r, const Layout &Lr, int ex, int slot) {   // slot 0: raised by the upper neighbour
    return reinterpret_cast<ReduceUnit *>(d->peer[r] + Lr.flags) + ex * 2 + slot;
}

// Push `nrows` boundary rows of each listed field to both neighbours and wait for theirs.
// the second half of an exchange: a one-thread kernel that holds the stream until both neighbours' rows (flags) have arrived
int exchange_wait(pano_dist *d, int ex) {
    pano_ctx *ctx = d->ctx;
    const Layout &L = d->L;
    const bool has_up = d->rank > 0, has_dn = d->rank + 1 < d->nranks;
    if (!has_up && !has_dn) return PANO_OK;
    k_wait<<<1, 1, 0, ctx->stream>>>(has_up ? flag_of(d, d->rank, L, ex, 0) : nullptr, has_dn ? flag_of(d, d->rank, L, ex, 1) : nullptr,
                                     d->step_no, reinterpret_cast<unsigned int *>(d->window + L.err));
    return pano_after_launch(ctx, "dist_wait");
}

int exchange(pano_dist *d, int ex, int nfields, const int *fields, int nrows, bool wait_too = true) {
    pano_ctx *ctx = d->ctx;
    const Layout &L = d->L;
    const bool has_up = d->rank > 0, has_dn = d->rank + 1 < d->nranks;
    if (!has_up && !has_dn) return PANO_OK;
    Layout Lup, Ldn;
    if (has_up) Lup = make_layout(d->H, d->W, d->rank - 1, d->nranks);
    if (has_dn) Ldn = make_layout(d->H, d->W, d->rank + 1, d->nranks);
    Segments seg;
    seg.count = 0;
    size_t total = 0;
    for (int i = 0; i < nfields; ++i) {
        const int f = fields[i];
        const size_t P = L.pitch[f];
        if (has_up) {   // my first rows, global [y0, y0 + nrows), are the upper neighbour's ghost rows below its slab
            seg.src[seg.count] = peer_row(d, d->rank, L, f, (ptrdiff_t)L.y0);
            seg.dst[seg.count] = peer_row(d, d->rank - 1, Lup, f, (ptrdiff_t)L.y0);
            seg.n[seg.count] = (size_t)nrows * P;
            total += seg.n[seg.count++];
        }
        if (has_dn) {   // my last rows, global [y1 - nrows, y1), are the lower neighbour's ghost rows above its slab
            seg.src[seg.count] = peer_row(d, d->rank, L, f, (ptrdiff_t)L.y1 - nrows);
            seg.dst[seg.count] = peer_row(d, d->rank + 1, Ldn, f, (ptrdiff_t)L.y1 - nrows);
            seg.n[seg.count] = (size_t)nrows * P;
            total += seg.n[seg.count++];
        }
    }
    const unsigned long long seq = d->step_no;
    // my flag slots in the neighbours' windows: at the upper neighbour I am "the lower neighbour" (slot 1), and vice versa
    ReduceUnit *flag_up = has_up ? flag_of(d, d->rank - 1, Lup, ex, 1) : nullptr;
    ReduceUnit *flag_dn = has_dn ? flag_of(d, d->rank + 1, Ldn, ex, 0) : nullptr;
    size_t blocks = (total + kThreads * 8 - 1) / (kThreads * 8);
    if (blocks < 1) blocks = 1;
    if (blocks > 64) blocks = 64;
    k_push<<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(seg, flag_up, flag_dn, seq, d->d_counter);
    PANO_TRY(pano_after_launch(ctx, "dist_push"));
    if (!wait_too) return PANO_OK;
    return exchange_wait(d, ex);
}

int fill_owned(pano_dist *d, int f, pano_rect r, double value) {
    // rows of the rectangle that this rank owns (the vy face row y belongs to the owner of cell row y)
    const Layout &L = d->L;
    int64_t ya = r.y0 > (int64_t)L.y0 ? r.y0 : (int64_t)L.y0;
    int64_t yb = r.y1 < (int64_t)L.y1 ? r.y1 : (int64_t)L.y1;
    if (yb <= ya || r.x1 <= r.x0) return PANO_OK;
    const size_t cells = (size_t)(yb - ya) * (size_t)(r.x1 - r.x0);
    size_t blocks = (cells + kThreads - 1) / kThreads;
    if (blocks > 1024) blocks = 1024;
    k_fill_rows<<<(unsigned)blocks, kThreads, 0, d->ctx->stream>>>(virt(d, f), L.pitch[f], (int)ya, (int)yb, (int)r.x0, (int)r.x1, value);
    return pano_after_launch(d->ctx, "dist_fill");
}


// option "cg_single_reduction" (-1 auto = on for slabs, 0, 1; must agree on all ranks): the Chronopoulos-Gear kernel of pano_cg_sr.cu
bool use_single_reduction(pano_dist *d) { return pano_option(d->ctx, "cg_single_reduction", -1) != 0; }   // -1 auto: yes on slabs

// The persistent CG kernel on this rank's slab, right-hand side in array fB; x lands in F_P.
int launch_cg(pano_dist *d, int fB, bool mirror_x) {
    pano_ctx *ctx = d->ctx;
    const Layout &L = d->L;
    const pano_step_params &p = d->prm;
    const size_t H = d->H, W = d->W;
    const int ya = (int)L.y0;
    ++d->solve_no;
    if (use_single_reduction(d)) {
        // one reduction per iteration (pano_cg_sr.cu): r and s double-buffered, two halo rows of r, one of s
        PanoCgSrSlab s;
        memset(&s, 0, sizeof(s));
        s.row0 = kGhost;
        s.rows_total = (int)L.rows[F_P];
        s.gy0 = ya;
        s.gh = (int)H;
        s.rank = d->rank;
        s.nranks = d->nranks;
        s.xseq_base = d->solve_no << 32;
        s.max_ctas = d->max_ctas;
        const int fr[2] = {F_R, F_R1}, fs[2] = {F_S1, F_S2};
        if (d->rank > 0) {
            const Layout Lup = make_layout(H, W, d->rank - 1, d->nranks);
            for (int i = 0; i < 2; ++i) {   // my rows y0, y0+1 = its first ghost rows below its slab
                s.up_r[i] = peer_row(d, d->rank - 1, Lup, fr[i], (ptrdiff_t)L.y0);
                s.up_s[i] = peer_row(d, d->rank - 1, Lup, fs[i], (ptrdiff_t)L.y0);
            }
        }
        if (d->rank + 1 < d->nranks) {
            const Layout Ldn = make_layout(H, W, d->rank + 1, d->nranks);
            for (int i = 0; i < 2; ++i) {   // my rows y1-2, y1-1 = its last ghost rows above its slab
                s.dn_r[i] = peer_row(d, d->rank + 1, Ldn, fr[i], (ptrdiff_t)L.y1 - 2);
                s.dn_s[i] = peer_row(d, d->rank + 1, Ldn, fs[i], (ptrdiff_t)L.y1 - 1);
            }
            // my last row of x = the p[y-1] its projection reads for its first face row (fused-halo step: no exchange of p)
            if (mirror_x) s.dn_x = peer_row(d, d->rank + 1, Ldn, F_P, (ptrdiff_t)L.y1 - 1);
        }
        for (int r = 0; r < d->nranks; ++r) {
            const Layout Lr = r == d->rank ? L : make_layout(H, W, r, d->nranks);
            s.xunits_peer[r] = d->peer[r] + Lr.xunits;
        }
        s.xunits_local = d->window + L.xunits;
        const RectI m = pano_clip_rect(p.obstacle, H + 1, W + 1);
        return pano_cg_sr_launch(ctx, d->window + L.off[F_P], d->window + L.off[fB], d->window + L.off[F_R], d->window + L.off[F_S0],
                                 d->window + L.off[F_S1], d->window + L.off[F_R1], d->window + L.off[F_S2], L.hl, W, p.max_iterations,
                                 p.threshold, p.timestep, m, &s);
    }
    {
        PanoCgSlab s;
        memset(&s, 0, sizeof(s));
        s.row0 = kGhost;
        s.rows_total = (int)L.rows[F_P];
        s.gy0 = ya;
        s.gh = (int)H;
        s.rank = d->rank;
        s.nranks = d->nranks;
        s.xseq_base = d->solve_no << 32;
        s.max_ctas = d->max_ctas;
        if (d->rank > 0) {
            const Layout Lup = make_layout(H, W, d->rank - 1, d->nranks);
            s.up_r = peer_row(d, d->rank - 1, Lup, F_R, (ptrdiff_t)L.y0);     // global row y0 = first ghost row below its slab
            s.up_s0 = peer_row(d, d->rank - 1, Lup, F_S0, (ptrdiff_t)L.y0);
            s.up_s1 = peer_row(d, d->rank - 1, Lup, F_S1, (ptrdiff_t)L.y0);
        }
        if (d->rank + 1 < d->nranks) {
            const Layout Ldn = make_layout(H, W, d->rank + 1, d->nranks);
            s.dn_r = peer_row(d, d->rank + 1, Ldn, F_R, (ptrdiff_t)L.y1 - 1); // global row y1-1 = last ghost row above its slab
            s.dn_s0 = peer_row(d, d->rank + 1, Ldn, F_S0, (ptrdiff_t)L.y1 - 1);
            s.dn_s1 = peer_row(d, d->rank + 1, Ldn, F_S1, (ptrdiff_t)L.y1 - 1);
        }
        for (int r = 0; r < d->nranks; ++r) {
            const Layout Lr = r == d->rank ? L : make_layout(H, W, r, d->nranks);
            s.xunits_peer[r] = d->peer[r] + Lr.xunits;
        }
        s.xunits_local = d->window + L.xunits;
        const RectI m = pano_clip_rect(p.obstacle, H + 1, W + 1);
        PANO_TRY(pano_cg_stream_launch(ctx, d->window + L.off[F_P], d->window + L.off[fB], d->window + L.off[F_R], d->window + L.off[F_S0],
                                       d->window + L.off[F_S1], L.hl, W, p.max_iterations, p.threshold, p.timestep, m, &s));
    }
    return PANO_OK;
}
}  // namespace

extern "C" {

int pano_slab_range(size_t h, int rank, int nranks, size_t *y0, size_t *y1) {
    if (!y0 || !y1 || nranks < 1 || rank < 0 || rank >= nranks) PANO_FAIL(PANO_ERR_INVALID, "pano_slab_range: bad arguments");
    slab_range(h, rank, nranks, y0, y1);
    return PANO_OK;
}

int pano_dist_create(pano_ctx *ctx, size_t h, size_t w, int rank, int nranks, const pano_step_params *params, pano_dist **out) {
    if (!ctx || !params || !out) PANO_FAIL(PANO_ERR_INVALID, "pano_dist_create: null argument");
    *out = nullptr;
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks)
        PANO_FAIL(PANO_ERR_INVALID, "pano_dist_create: rank %d of %d (at most %d ranks)", rank, nranks, kMaxRanks);
    if (w % 2 != 0 || w < 2) PANO_FAIL(PANO_ERR_SHAPE, "pano_dist_create: the slab solver needs an even width (TMA rows are 16-byte aligned)");
    if (params->max_iterations <= 0)
        PANO_FAIL(PANO_ERR_INVALID, "pano_dist_create: max_iterations = %d; the slab solver runs at least one iteration", (int)params->max_iterations);
    if (params->precond != PANO_PRECOND_IDENTITY) PANO_FAIL(PANO_ERR_UNIMPLEMENTED, "pano_dist_create: only the identity preconditioner exists");
    PANO_TRY(pano_check_rect_within(params->inflow, h, w, "pano_dist_create(inflow)"));
    PANO_TRY(pano_check_rect_within(params->obstacle, h, w, "pano_dist_create(obstacle)"));
    for (int r = 0; r < nranks; ++r) {
        size_t a, b;
        slab_range(h, r, nranks, &a, &b);
        if (b - a < (size_t)kGhost) PANO_FAIL(PANO_ERR_SHAPE, "pano_dist_create: slab of rank %d has %zu rows, fewer than the %d ghost rows", r, b - a, kGhost);
    }
    PANO_TRY(pano_activate(ctx));
    {   // load every kernel of the step now (see pano_preload_fused)
        cudaFuncAttributes fa;
        PANO_CUDA(cudaFuncGetAttributes(&fa, k_push));
        PANO_CUDA(cudaFuncGetAttributes(&fa, k_wait));
        PANO_CUDA(cudaFuncGetAttributes(&fa, k_fill_rows));
        PANO_TRY(pano_preload_fused());
        PANO_TRY(pano_preload_cg_stream());
        PANO_TRY(pano_preload_cg_sr());
    }
    pano_dist *d = new pano_dist();
    d->ctx = ctx;
    d->rank = rank;
    d->nranks = nranks;
    d->H = h;
    d->W = w;
    d->prm = *params;
    d->L = make_layout(h, w, rank, nranks);
    cudaError_t e = cudaMalloc(&d->window, d->L.total * sizeof(double));
    if (e != cudaSuccess) {
        const size_t bytes = d->L.total * sizeof(double);
        delete d;
        PANO_FAIL(PANO_ERR_CUDA, "pano_dist_create: cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
    }
    e = cudaMemsetAsync(d->window, 0, d->L.total * sizeof(double), ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&d->d_counter, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemsetAsync(d->d_counter, 0, sizeof(unsigned int), ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {   // nothing half-built survives a failed create
        cudaFree(d->window);
        cudaFree(d->d_counter);
        delete d;
        PANO_FAIL(PANO_ERR_CUDA, "pano_dist_create: %s", cudaGetErrorString(e));
    }
    d->peer[rank] = d->window;
    d->connected = nranks == 1;
    *out = d;
    return PANO_OK;
}

int pano_dist_destroy(pano_dist *d) {
    if (!d) return PANO_OK;
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    for (int r = 0; r < d->nranks; ++r)
        if (r != d->rank && d->peer[r] && d->peer_ipc[r]) cudaIpcCloseMemHandle(d->peer[r]);
    cudaFree(d->window);
    cudaFree(d->d_counter);
    delete d;
    return PANO_OK;
}

int pano_dist_window(pano_dist *d, void **ptr, size_t *bytes) {
    if (!d || !ptr || !bytes) PANO_FAIL(PANO_ERR_INVALID, "pano_dist_window: null argument");
    *ptr = d->window;
    *bytes = d->L.total * sizeof(double);
    return PANO_OK;
}

int pano_dist_ipc_handle(pano_dist *d, void *handle_out) {
    if (!d || !handle_out) PANO_FAIL(PANO_ERR_INVALID, "pano_dist_ipc_handle: null argument");
    PANO_TRY(pano_activate(d->ctx));
    static_assert(sizeof(cudaIpcMemHandle_t) == PANO_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t hnd;
    PANO_CUDA(cudaIpcGetMemHandle(&hnd, d->window));
    memcpy(handle_out, &hnd, sizeof(hnd));
    return PANO_OK;
}

int pano_dist_connect(pano_dist *d, int kind, const void *peers) {
    if (!d || !peers) PANO_FAIL(PANO_ERR_INVALID, "pano_dist_connect: null argument");
    PANO_TRY(pano_activate(d->ctx));
    for (int r = 0; r < d->nranks; ++r) {
        if (r == d->rank) continue;
        if (kind == 0) {   // raw device pointers: ranks share this process (loop-back tests on one GPU, or peer-enabled GPUs)
            d->peer[r] = reinterpret_cast<double *const *>(peers)[r];
            if (!d->peer[r]) PANO_FAIL(PANO_ERR_COMM, "pano_dist_connect: null window for rank %d", r);
        } else {           // CUDA IPC handles gathered from the other processes
            cudaIpcMemHandle_t hnd;
            memcpy(&hnd, reinterpret_cast<const unsigned char *>(peers) + (size_t)r * PANO_IPC_HANDLE_BYTES, sizeof(hnd));
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) PANO_FAIL(PANO_ERR_COMM, "pano_dist_connect: cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
            d->peer[r] = (double *)p;
            d->peer_ipc[r] = true;
        }
    }
    d->connected = true;
    return PANO_OK;
}

int pano_dist_set_max_ctas(pano_dist *d, int max_ctas) {
    if (!d) PANO_FAIL(PANO_ERR_INVALID, "pano_dist_set_max_ctas: null handle");
    d->max_ctas = max_ctas;
    return PANO_OK;
}

// which: 0 density (h*w), 1 vy ((h+1)*w), 2 vx (h*(w+1)), 3 pressure (h*w).  Transfers the rows this rank owns:
// host points at the first owned row; *rows receives their number.
static int dist_rows(pano_dist *d, int which, int *f, size_t *g0, size_t *nrows) {
    const Layout &L = d->L;
    switch (which) {
        case 0: *f = d->cur ? F_D1 : F_D0; break;
        case 1: *f = d->cur ? F_VY1 : F_VY0; break;
        case 2: *f = d->cur ? F_VX1 : F_VX0; break;
        case 3: *f = F_P; break;
        default: PANO_FAIL(PANO_ERR_INVALID, "pano_dist: unknown field selector %d", which);
    }
    *g0 = L.y0;
    *nrows = L.hl + ((which == 1 && d->rank == d->nranks - 1) ? 1 : 0);
    return PANO_OK;
}

int pano_dist_upload(pano_dist *d, int which, const double *host_rows) {
    if (!d || !host_rows) PANO_FAIL(PANO_ERR_INVALID, "pano_dist_upload: null argument");
    PANO_TRY(pano_activate(d->ctx));
    int f;
    size_t g0, n;
    PANO_TRY(dist_rows(d, which, &f, &g0, &n));
    PANO_CUDA(cudaMemcpyAsync(peer_row(d, d->rank, d->L, f, (ptrdiff_t)g0), host_rows, n * d->L.pitch[f] * sizeof(double),
                              cudaMemcpyHostToDevice, d->ctx->stream));
    PANO_CUDA(cudaStreamSynchronize(d->ctx->stream));
    return PANO_OK;
}

int pano_dist_download(pano_dist *d, int which, double *host_rows, size_t *rows) {
    if (!d || !host_rows) PANO_FAIL(PANO_ERR_INVALID, "pano_dist_download: null argument");
    PANO_TRY(pano_activate(d->ctx));
    int f;
    size_t g0, n;
    PANO_TRY(dist_rows(d, which, &f, &g0, &n));
    PANO_CUDA(cudaMemcpyAsync(host_rows, peer_row(d, d->rank, d->L, f, (ptrdiff_t)g0), n * d->L.pitch[f] * sizeof(double),
                              cudaMemcpyDeviceToHost, d->ctx->stream));
    PANO_CUDA(cudaStreamSynchronize(d->ctx->stream));
    if (rows) *rows = n;
    return PANO_OK;
}

// after_advect (nullable) runs once the new density is final, long before the solve ends: the host-buffer entry point starts the
// density download on the copy stream there (as pano_fluid_step_host does on one GPU)
typedef int (*DistStepHook)(pano_dist *d, void *user);
static int dist_step_impl(pano_dist *d, DistStepHook after_advect, void *user);

// One step, enqueued asynchronously on the context's stream (collective: every rank must call it).
int pano_dist_step(pano_dist *d) { return dist_step_impl(d, nullptr, nullptr); }

static int dist_step_impl(pano_dist *d, DistStepHook after_advect, void *user) {
    if (!d) PANO_FAIL(PANO_ERR_INVALID, "pano_dist_step: null handle");
    if (!d->connected) PANO_FAIL(PANO_ERR_COMM, "pano_dist_step: call pano_dist_connect first");
    pano_ctx *ctx = d->ctx;
    PANO_TRY(pano_activate(ctx));
    const Layout &L = d->L;
    const pano_step_params &p = d->prm;
    const size_t H = d->H, W = d->W;
    const int ya = (int)L.y0, yb = (int)L.y1;
    const int cur = d->cur, nxt = cur ^ 1;
    const int fD = cur ? F_D1 : F_D0, fDn = nxt ? F_D1 : F_D0;
    const int fVY = cur ? F_VY1 : F_VY0, fVYn = nxt ? F_VY1 : F_VY0;
    const int fVX = cur ? F_VX1 : F_VX0, fVXn = nxt ? F_VX1 : F_VX0;
    unsigned int *err = reinterpret_cast<unsigned int *>(d->window + L.err);
    ++d->step_no;
    static const bool dbg = getenv("PANO_DIST_DEBUG") != nullptr;
    struct timespec ts0;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    auto mark = [&](const char *what) {
        if (!dbg) return;
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        fprintf(stderr, "[pano_dist rank %d] %-12s +%.3f ms\n", d->rank, what, (ts.tv_sec - ts0.tv_sec) * 1e3 + (ts.tv_nsec - ts0.tv_nsec) * 1e-6);
    };

    const bool fused_halos = use_single_reduction(d) && pano_option(ctx, "dist_fused_halos", 1) != 0;
    const int ylo = ya - kGhost;                     // first stored row of every array (storage, not data, where negative)
    const size_t rows_q = L.hl + 2 * kGhost;
    PANO_TRY(pano_phase_mark(ctx, 0));
    // inflow  (dec_fluid.rs:48-57): the part of the rectangle this rank owns
    PANO_TRY(fill_owned(d, fD, p.inflow, p.inflow_density));
    PANO_TRY(fill_owned(d, fVY, p.inflow, p.inflow_vy));
    // ghost rows of everything the advection gathers from.  Option "dist_overlap" = 1 (fused halos): the stream does not wait for
    // the neighbours' rows here -- the rows at least kGhost away from both slab edges gather from owned rows only, so they are
    // advected while the halos cross NVLink; the wait and the two edge strips follow.  Measured on 2 B200 (2048 x 8192 slabs,
    // profiles/r02_dist_overlap_2gpu.txt): 10.778 ms per step against 10.762 without -- the exchange is ~5 us of transfer behind
    // ~25 us of launches, and the two strip launches cost more than the wait they hide.  Default 0; kept and tested.
    const bool overlap = fused_halos && pano_option(ctx, "dist_overlap", 0) != 0 && (yb - ya) > 4 * kGhost;
    {
        const int fields[3] = {fD, fVY, fVX};
        PANO_TRY(exchange(d, EX_ADV, 3, fields, kGhost, !overlap));
    }
    mark("ex_adv");
    PANO_TRY(pano_phase_mark(ctx, 1));
    const int fB = fD;                               // b reuses the old density buffer, as `temp` does in the reference
    if (fused_halos) {
        // ONE exchange per step.  Every rank advects kExtend rows beyond its slab from the ghost rows it has just received
        // (same inputs, same code as their owner: the same bits), so the new vy face row y1, the two ghost rows of b the
        // solver's halo recomputation reads, ... are all local; the solver itself mirrors its last row of x into the lower
        // neighbour's ghost row, which is the p[y0-1] of that neighbour's projection.
        const int A = ya - kExtend > 0 ? ya - kExtend : 0, B = yb + kExtend < (int)H ? yb + kExtend : (int)H;
        if (overlap) {
            const bool has_up = d->rank > 0, has_dn = d->rank + 1 < d->nranks;
            const int ia = has_up ? ya + kGhost : A, ib = has_dn ? yb - kGhost : B;
            // the interior pass may only SEE rows this rank owns (the ghost rows are still in flight): a backtrace that leaves
            // them raises the slab error word instead of reading stale rows (the edge strips bound the supported backtrace
            // more tightly anyway: kGhost - kExtend - 2 rows)
            const int wlo_i = has_up ? ya : ylo, whi_i = has_dn ? yb - 2 : ylo + (int)rows_q;
            PANO_TRY(pano_advect_slab_launch(ctx, virt(d, fDn), virt(d, fVYn), virt(d, fVXn), virt(d, fD), virt(d, fVY), virt(d, fVX), H, W,
                                             p.timestep, ia, ib, wlo_i, (size_t)(whi_i - wlo_i), err));
            PANO_TRY(exchange_wait(d, EX_ADV));
            if (ia > A)
                PANO_TRY(pano_advect_slab_launch(ctx, virt(d, fDn), virt(d, fVYn), virt(d, fVXn), virt(d, fD), virt(d, fVY), virt(d, fVX), H, W,
                                                 p.timestep, A, ia, ylo, rows_q, err));
            if (ib < B)
                PANO_TRY(pano_advect_slab_launch(ctx, virt(d, fDn), virt(d, fVYn), virt(d, fVXn), virt(d, fD), virt(d, fVY), virt(d, fVX), H, W,
                                                 p.timestep, ib, B, ylo, rows_q, err));
        } else {
            PANO_TRY(pano_advect_slab_launch(ctx, virt(d, fDn), virt(d, fVYn), virt(d, fVXn), virt(d, fD), virt(d, fVY), virt(d, fVX), H, W,
                                             p.timestep, A, B, ylo, rows_q, err));
        }
        mark("advect");
        if (after_advect) PANO_TRY(after_advect(d, user));
        PANO_TRY(pano_phase_mark(ctx, 2));
        const int A2 = ya - 2 > 0 ? ya - 2 : 0, B2 = yb + 2 < (int)H ? yb + 2 : (int)H;
        PANO_TRY(pano_neg_divergence_slab_launch(ctx, virt(d, fB), virt(d, fVYn), virt(d, fVXn), H, W, p.obstacle, A2, B2));
        PANO_TRY(pano_phase_mark(ctx, 3));
        PANO_TRY(launch_cg(d, fB, true));
        mark("cg");
        PANO_TRY(pano_phase_mark(ctx, 4));
        PANO_TRY(pano_project_slab_launch(ctx, virt(d, fVYn), virt(d, fVXn), virt(d, F_P), H, W, p.timestep, ya, yb));
        PANO_TRY(pano_phase_mark(ctx, 5));
        mark("project");
        d->cur = nxt;
        return PANO_OK;
    }
    // advect + advect_mac on the owned rows, into the other ping-pong buffers  (:59-63)
    PANO_TRY(pano_advect_slab_launch(ctx, virt(d, fDn), virt(d, fVYn), virt(d, fVXn), virt(d, fD), virt(d, fVY), virt(d, fVX), H, W,
                                     p.timestep, ya, yb, ylo, rows_q, err));
    mark("advect");
    if (after_advect) PANO_TRY(after_advect(d, user));
    // the face row y1 of the new vy belongs to the lower neighbour
    {
        const int fields[1] = {fVYn};
        PANO_TRY(exchange(d, EX_VY, 1, fields, 1));
    }
    PANO_TRY(pano_phase_mark(ctx, 2));
    // b = -div  (:69-83)
    PANO_TRY(pano_neg_divergence_slab_launch(ctx, virt(d, fB), virt(d, fVYn), virt(d, fVXn), H, W, p.obstacle, ya, yb));
    {
        const int fields[1] = {fB};
        PANO_TRY(exchange(d, EX_B, 1, fields, use_single_reduction(d) ? 2 : 1));   // the single-reduction kernel reads a two-row halo of b
    }
    mark("ex_b");
    PANO_TRY(pano_phase_mark(ctx, 3));
    // pressure solve  (:91-119): streaming CG on the slab, halo rows and reductions over NVLink from inside the kernel
    PANO_TRY(launch_cg(d, fB, false));
    mark("cg");
    PANO_TRY(pano_phase_mark(ctx, 4));
    // p[y0 - 1] lives on the upper neighbour
    {
        const int fields[1] = {F_P};
        PANO_TRY(exchange(d, EX_P, 1, fields, 1));
    }
    // projection + walls  (:124-141)
    PANO_TRY(pano_project_slab_launch(ctx, virt(d, fVYn), virt(d, fVXn), virt(d, F_P), H, W, p.timestep, ya, yb));
    PANO_TRY(pano_phase_mark(ctx, 5));
    mark("project");
    d->cur = nxt;
    return PANO_OK;
}

// The solve alone, again on the right-hand side of the last step (it still sits in the old density buffer).
int pano_dist_solve(pano_dist *d) {
    if (!d) PANO_FAIL(PANO_ERR_INVALID, "pano_dist_solve: null handle");
    if (!d->connected) PANO_FAIL(PANO_ERR_COMM, "pano_dist_solve: call pano_dist_connect first");
    if (d->step_no == 0) PANO_FAIL(PANO_ERR_INVALID, "pano_dist_solve: no step has produced a right-hand side yet");
    PANO_TRY(pano_activate(d->ctx));
    return launch_cg(d, (d->cur ^ 1) ? F_D1 : F_D0, true);
}

// wait for the enqueued steps; info (nullable) describes the last solve
struct DistHostCopy {
    double *density_rows;
};
static int dist_start_density_download(pano_dist *d, void *user) {
    // the advection has just written the NEXT density buffer; cur flips at the end of the step
    pano_ctx *ctx = d->ctx;
    const Layout &L = d->L;
    const int f = (d->cur ^ 1) ? F_D1 : F_D0;
    PANO_CUDA(cudaEventRecord(ctx->ev_advect, ctx->stream));
    PANO_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_advect, 0));
    PANO_CUDA(cudaMemcpyAsync(static_cast<DistHostCopy *>(user)->density_rows, peer_row(d, d->rank, L, f, (ptrdiff_t)L.y0),
                              L.hl * L.pitch[f] * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
    PANO_CUDA(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
    return PANO_OK;
}

int pano_dist_step_host(pano_dist *d, double *density_rows, double *vy_rows, double *vx_rows, pano_pcg_info *info) {
    if (!d || !density_rows || !vy_rows || !vx_rows) PANO_FAIL(PANO_ERR_INVALID, "pano_dist_step_host: null argument");
    pano_ctx *ctx = d->ctx;
    PANO_TRY(pano_activate(ctx));
    double *host[3] = {density_rows, vy_rows, vx_rows};
    for (int which = 0; which < 3; ++which) {          // this rank's rows in, back to back on the step's stream
        int f;
        size_t g0, n;
        PANO_TRY(dist_rows(d, which, &f, &g0, &n));
        PANO_CUDA(cudaMemcpyAsync(peer_row(d, d->rank, d->L, f, (ptrdiff_t)g0), host[which], n * d->L.pitch[f] * sizeof(double),
                                  cudaMemcpyHostToDevice, ctx->stream));
    }
    DistHostCopy hook{density_rows};
    PANO_TRY(dist_step_impl(d, dist_start_density_download, &hook));     // the density goes home under the solve
    PANO_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
    for (int which = 1; which < 3; ++which) {          // the projected velocity follows the last kernel
        int f;
        size_t g0, n;
        PANO_TRY(dist_rows(d, which, &f, &g0, &n));    // (cur has flipped: the new buffers)
        PANO_CUDA(cudaMemcpyAsync(host[which], peer_row(d, d->rank, d->L, f, (ptrdiff_t)g0), n * d->L.pitch[f] * sizeof(double),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    }
    return pano_dist_sync(d, info);
}

int pano_dist_sync(pano_dist *d, pano_pcg_info *info) {
    if (!d) PANO_FAIL(PANO_ERR_INVALID, "pano_dist_sync: null handle");
    pano_ctx *ctx = d->ctx;
    PANO_TRY(pano_activate(ctx));
    PANO_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(PanoCgControl), cudaMemcpyDeviceToHost, ctx->stream));
    PANO_CUDA(cudaMemcpyAsync(ctx->h_scalars, d->window + d->L.err, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PANO_CUDA(cudaStreamSynchronize(ctx->stream));
    const unsigned int derr = *reinterpret_cast<unsigned int *>(ctx->h_scalars);
    if (derr != 0u) cudaMemsetAsync(d->window + d->L.err, 0, sizeof(unsigned int), ctx->stream);   // reported once; the handle stays usable
    if (derr == 1u) {
        ctx->h_cg->error = 0;
        cudaMemsetAsync(&ctx->d_cg->error, 0, sizeof(unsigned int), ctx->stream);
        PANO_FAIL(PANO_ERR_TIMEOUT, "pano_dist: a bounded device-side wait expired (rank %d)", d->rank);
    }
    PANO_TRY(pano_check_device_error(ctx, "pano_dist_sync"));
    if (derr == 2u)
        PANO_FAIL(PANO_ERR_SHAPE, "pano_dist: an advection backtrace left the %d ghost rows (rank %d): dt*max|v| is too large for the slab", kGhost, d->rank);
    if (info) {
        info->iterations = ctx->h_cg->iterations;
        info->applies = ctx->h_cg->applies;
        info->final_residual = ctx->h_cg->final_residual;
        info->rhs_max = ctx->h_cg->rhs_max;
    }
    return PANO_OK;
}

}  // extern "C"
