// pano_cg_stream.cu -- the pressure solve for grids that do not fit on chip: a persistent,
// warp-specialised, TMA-pipelined conjugate-gradient kernel (one CTA per SM).
//
// Same algorithm and arithmetic as k_cg_generic (pano_cg.cu; pcg.rs:14-82 + dec_fluid.rs:100-119):
//   P1: s' = r + beta*s on every tile INCLUDING its one-cell halo, z = A s' (not stored), z.s'
//   P2: z recomputed from s', x += alpha s', r -= alpha z, r.r, max|r|
// What changes is the data movement:
//   * a producer warp streams (TH+2)x(TW+2) halo boxes of r / s and TH x TW boxes of r / x into a
//     4-stage shared-memory ring with cp.async.bulk.tensor (TMA), completion on mbarriers;
//     out-of-range box elements are zero-filled by the TMA unit, which is exactly what a closed
//     (Neumann) wall edge needs
//   * 8 consumer warps run the stencil out of shared memory with a vertical register window
//     (3 shared loads per cell) and write results with coalesced global stores
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
constexpr int BH = TH + 2, BW = TW + 2;         // halo box
constexpr int kBoxElems = BH * BW;              // 2244 (even)
constexpr int kHaloBoxBytes = kBoxElems * 8;    // 17952 = TMA transaction size
constexpr int kHaloSlot = 18048;                // rounded up to a multiple of 128
constexpr int kIntBoxBytes = TH * TW * 8;       // 16384
constexpr int kStageBytes = 2 * kHaloSlot + kIntBoxBytes;   // 52480
constexpr int kStages = 4;
constexpr int kConsumers = 256, kConsumerWarps = 8;
constexpr int kThreads = kConsumers + 32;       // + one producer warp
constexpr int kMaxCtas = 192;
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
    ReduceUnit *units;            // [2 banks][3 values][kMaxCtas]
    unsigned long long seq_base;
    PanoCgControl *ctl;
    int zigzag;
};

struct Tail {                     // small shared-memory area behind the stage ring
    uint64_t full[kStages], empty[kStages], go;
    double vals[3][kMaxCtas];
    double out[4];
    double wsum[3][kConsumerWarps];
    int cont;                     // 1: producer continues with the next phase, 0: stop
    int ok;
};

__device__ __forceinline__ void consumer_sync() { named_bar_sync(1, kConsumers); }

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

// Grid-wide all-reduce of up to three block totals (kinds: 0/1 sum, 2 max unless max_mask says otherwise).
// Called by all consumer threads.  Returns false if a bounded wait expired.
__device__ __forceinline__ bool grid_allreduce(const StreamArgs &a, Tail *tl, unsigned long long n, int nvals, double v0,
                                               double v1, double v2, unsigned max_mask, double *out) {
    const int G = gridDim.x, tid = threadIdx.x;
    const unsigned long long seq = a.seq_base + n;
    ReduceUnit *bank = a.units + (n & 1) * 3 * kMaxCtas;
    if (tid == 0) {
        __threadfence();
        fence_proxy_async();
        unit_store(bank + 0 * kMaxCtas + blockIdx.x, v0, seq);
        if (nvals > 1) unit_store(bank + 1 * kMaxCtas + blockIdx.x, v1, seq);
        if (nvals > 2) unit_store(bank + 2 * kMaxCtas + blockIdx.x, v2, seq);
    }
    bool ok = true;
    if (tid < G) {
        volatile unsigned int *err = &a.ctl->error;
        for (int k = 0; k < nvals && ok; ++k) {
            double v;
            ok = unit_poll(bank + k * kMaxCtas + tid, seq, v, err);
            tl->vals[k][tid] = v;
        }
        __threadfence();
        if (!ok) tl->ok = 0;
    }
    consumer_sync();
    const int wid = tid >> 5, lane = tid & 31;
    if (wid < nvals) {
        double r = ((max_mask >> wid) & 1u) ? warp_fixed_max(tl->vals[wid], G, lane) : warp_fixed_sum(tl->vals[wid], G, lane);
        if (lane == 0) tl->out[wid] = r;
    }
    consumer_sync();
    out[0] = tl->out[0];
    if (nvals > 1) out[1] = tl->out[1];
    if (nvals > 2) out[2] = tl->out[2];
    return tl->ok != 0;
}

// ------------------------------------------------------------------------------ tile kernels
// Consumer warp `wid` owns columns 32*(wid&1) + lane and rows 8*(wid>>1) .. +7 of the tile.
// S: halo box (BH x BW), cell (ty, tx) at S[(ty+1)*BW + tx+1].
template <bool kFast, bool kFirst>
__device__ __forceinline__ void tile_p1(const StreamArgs &a, const double *S, double *s_dst, int ty0, int tx0,
                                        double &acc_zs, double &acc_bb, double &acc_bmax) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = (wid & 1) * 32 + lane, row0 = (wid >> 1) * 8;
    const int gx = tx0 + col;
    const double *p = S + (row0 + 1) * BW + col + 1;
    double up = p[-BW], c = p[0];
#pragma unroll
    for (int k = 0; k < 8; ++k, p += BW) {
        const double dn = p[BW], wv = p[-1], ev = p[1];
        const int gy = ty0 + row0 + k;
        double z;
        if (kFast) {
            z = pano::laplacian_cell<double>(c, up, dn, wv, ev, true, true, true, true, a.dt);
        } else {
            const bool valid = gy < a.h && gx < a.w;
            if (!valid) {
                up = c;
                c = dn;
                continue;
            }
            const bool oN = gy > 0 && !in_rect(a.m, gy, gx), oS = gy < a.h - 1 && !in_rect(a.m, gy + 1, gx);
            const bool oW = gx > 0 && !in_rect(a.m, gy, gx), oE = gx < a.w - 1 && !in_rect(a.m, gy, gx + 1);
            z = pano::laplacian_cell<double>(c, up, dn, wv, ev, oN, oS, oW, oE, a.dt);
        }
        acc_zs = acc_zs + z * c;
        if (kFirst) {
            const double ab = c < 0 ? -c : c;
            acc_bmax = ab > acc_bmax ? ab : acc_bmax;
            acc_bb = acc_bb + c * c;
        } else {
            s_dst[(size_t)gy * a.w + gx] = c;
        }
        up = c;
        c = dn;
    }
}

template <bool kFast, bool kFirst>
__device__ __forceinline__ void tile_p2(const StreamArgs &a, const double *S, const double *R, const double *X, int ty0,
                                        int tx0, double alpha, double &acc_rr, double &acc_rmax) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = (wid & 1) * 32 + lane, row0 = (wid >> 1) * 8;
    const int gx = tx0 + col;
    const double nalpha = -alpha;
    const double *p = S + (row0 + 1) * BW + col + 1;
    double up = p[-BW], c = p[0];
#pragma unroll
    for (int k = 0; k < 8; ++k, p += BW) {
        const double dn = p[BW], wv = p[-1], ev = p[1];
        const int gy = ty0 + row0 + k;
        double z;
        if (kFast) {
            z = pano::laplacian_cell<double>(c, up, dn, wv, ev, true, true, true, true, a.dt);
        } else {
            const bool valid = gy < a.h && gx < a.w;
            if (!valid) {
                up = c;
                c = dn;
                continue;
            }
            const bool oN = gy > 0 && !in_rect(a.m, gy, gx), oS = gy < a.h - 1 && !in_rect(a.m, gy + 1, gx);
            const bool oW = gx > 0 && !in_rect(a.m, gy, gx), oE = gx < a.w - 1 && !in_rect(a.m, gy, gx + 1);
            z = pano::laplacian_cell<double>(c, up, dn, wv, ev, oN, oS, oW, oE, a.dt);
        }
        const size_t gi = (size_t)gy * a.w + gx;
        const int ti = (row0 + k) * TW + col;
        const double xo = kFirst ? 0.0 : X[ti];
        const double ro = kFirst ? c : R[ti];            // iteration 0: r = s = b (pcg.rs:40-42)
        a.x[gi] = xo + alpha * c;                        // pcg.rs:55
        const double rn = ro + nalpha * z;               // pcg.rs:56
        a.r[gi] = rn;
        if (kFirst) a.s0[gi] = c;
        const double ar = rn < 0 ? -rn : rn;
        acc_rmax = ar > acc_rmax ? ar : acc_rmax;
        acc_rr = acc_rr + rn * rn;
        up = c;
        c = dn;
    }
}

__device__ __forceinline__ bool tile_is_fast(const StreamArgs &a, int ty0, int tx0) {
    if (ty0 < 1 || ty0 + TH > a.h - 1 || tx0 < 1 || tx0 + TW > a.w - 1) return false;
    if (a.m.y1 > a.m.y0 && a.m.x1 > a.m.x0 && ty0 < a.m.y1 && ty0 + TH > a.m.y0 - 1 && tx0 < a.m.x1 && tx0 + TW > a.m.x0 - 1)
        return false;
    return true;
}

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
        for (int it = 0; it < a.max_iter; ++it) {
            const bool first = it == 0;
            for (int phase = 0; phase < 2; ++phase) {
                fence_proxy_async();
                for (int jj = 0; jj < n_my; ++jj, ++n) {
                    const int j = (phase == 1 && a.zigzag) ? n_my - 1 - jj : jj;
                    const int t = blockIdx.x + j * G;
                    const int tx0 = (t % a.tiles_x) * TW, ty0 = (t / a.tiles_x) * TH;
                    const int st = n % kStages;
                    if (!mbar_wait(&tl->empty[st], ((n / kStages) & 1) ^ 1, err)) return;
                    unsigned char *base = smem + st * kStageBytes;
                    if (first) {
                        mbar_arrive_expect_tx(&tl->full[st], kHaloBoxBytes);
                        tma_load_2d(base, &a.m_b_halo, &tl->full[st], tx0 - 1, ty0 - 1);
                    } else if (phase == 0) {
                        mbar_arrive_expect_tx(&tl->full[st], 2 * kHaloBoxBytes);
                        tma_load_2d(base, cur ? &a.m_s0_halo : &a.m_s1_halo, &tl->full[st], tx0 - 1, ty0 - 1);   // s (old)
                        tma_load_2d(base + kHaloSlot, &a.m_r_halo, &tl->full[st], tx0 - 1, ty0 - 1);
                    } else {
                        mbar_arrive_expect_tx(&tl->full[st], kHaloBoxBytes + 2 * kIntBoxBytes);
                        tma_load_2d(base, cur ? &a.m_s1_halo : &a.m_s0_halo, &tl->full[st], tx0 - 1, ty0 - 1);   // s'
                        tma_load_2d(base + kHaloSlot, &a.m_r_int, &tl->full[st], tx0, ty0);
                        tma_load_2d(base + 2 * kHaloSlot, &a.m_x_int, &tl->full[st], tx0, ty0);
                    }
                }
                // the next phase reads what other CTAs wrote in this one: wait for the grid-wide reduction
                if (!mbar_wait(&tl->go, ngo & 1, err)) return;
                ++ngo;
                if (!*(volatile int *)&tl->cont) return;
            }
            cur ^= 1;
        }
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

    for (it = 0; it < a.max_iter; ++it) {
        const bool first = it == 0;
        // ------------------------------------------------------------------ P1
        double acc_zs = 0, acc_bb = 0, acc_bmax = 0;
        for (int jj = 0; jj < n_my; ++jj, ++n) {
            const int t = blockIdx.x + jj * G;
            const int tx0 = (t % a.tiles_x) * TW, ty0 = (t / a.tiles_x) * TH;
            const int st = n % kStages;
            if (!mbar_wait(&tl->full[st], (n / kStages) & 1, err)) return;
            double *S = reinterpret_cast<double *>(smem + st * kStageBytes);
            const bool fast = tile_is_fast(a, ty0, tx0);
            if (first) {
                if (fast) tile_p1<true, true>(a, S, nullptr, ty0, tx0, acc_zs, acc_bb, acc_bmax);
                else tile_p1<false, true>(a, S, nullptr, ty0, tx0, acc_zs, acc_bb, acc_bmax);
            } else {
                // s' = r + beta * s over the whole halo box, in place (16-byte shared accesses)
                double2 *S2 = reinterpret_cast<double2 *>(S);
                const double2 *R2 = reinterpret_cast<const double2 *>(smem + st * kStageBytes + kHaloSlot);
                for (int i = tid; i < kBoxElems / 2; i += kConsumers) {
                    double2 sv = S2[i];
                    const double2 rv = R2[i];
                    sv.x = rv.x + beta * sv.x;
                    sv.y = rv.y + beta * sv.y;
                    S2[i] = sv;
                }
                consumer_sync();
                if (fast) tile_p1<true, false>(a, S, s_cur, ty0, tx0, acc_zs, acc_bb, acc_bmax);
                else tile_p1<false, false>(a, S, s_cur, ty0, tx0, acc_zs, acc_bb, acc_bmax);
                fence_proxy_async();   // this stage was written with st.shared; the TMA unit overwrites it next
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&tl->empty[st]);
        }
        {
            const double v0 = consumer_sum(acc_zs, tl->wsum[0]);
            double v1 = 0, v2 = 0;
            if (first) {
                v1 = consumer_sum(acc_bb, tl->wsum[1]);
                v2 = consumer_max(acc_bmax, tl->wsum[2]);
            }
            if (!grid_allreduce(a, tl, nred++, first ? 3 : 1, v0, v1, v2, 0x4u, red)) {
                if (tid == 0) { tl->cont = 0; mbar_arrive(&tl->go); }
                return;
            }
        }
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
        for (int jj = 0; jj < n_my; ++jj, ++n) {
            const int j = a.zigzag ? n_my - 1 - jj : jj;
            const int t = blockIdx.x + j * G;
            const int tx0 = (t % a.tiles_x) * TW, ty0 = (t / a.tiles_x) * TH;
            const int st = n % kStages;
            if (!mbar_wait(&tl->full[st], (n / kStages) & 1, err)) return;
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
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&tl->empty[st]);
        }
        {
            const double v0 = consumer_sum(acc_rr, tl->wsum[0]);
            const double v1 = consumer_max(acc_rmax, tl->wsum[1]);
            if (!grid_allreduce(a, tl, nred++, 2, v0, v1, 0.0, 0x2u, red)) {
                if (tid == 0) { tl->cont = 0; mbar_arrive(&tl->go); }
                return;
            }
        }
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
    const size_t ncell = (size_t)a.h * a.w;
    const size_t stride = (size_t)G * kConsumers, i0 = (size_t)blockIdx.x * kConsumers + tid;
    if (early) {
        for (size_t i = i0; i < ncell; i += stride) a.x[i] = 0.0;
    } else {
        // converged: the last applied direction is in s_cur; exhausted: the swap already happened, it is
        // in s_old, and the reference still performs the search update (pcg.rs:72-77) before leaving
        const double *s_fin = converged ? s_cur : s_old;
        if (!converged || s_fin != a.s0) {
            for (size_t i = i0; i < ncell; i += stride) {
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

// TMA needs 16-byte aligned rows: even width, 16-byte aligned base pointers.
bool pano_cg_stream_supported(size_t h, size_t w, const void *x, const void *b, const void *r, const void *s0, const void *s1) {
    if (w % 2 != 0 || w < 2 || h < 1) return false;
    const void *ps[] = {x, b, r, s0, s1};
    for (const void *p : ps)
        if (((uintptr_t)p & 15u) != 0) return false;
    return true;
}

int pano_cg_stream_launch(pano_ctx *ctx, double *x, const double *b, double *r, double *s0, double *s1, size_t h, size_t w,
                          int max_iterations, double threshold, double timestep, RectI m) {
    PANO_CUDA(cudaFuncSetAttribute(k_cg_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    static_assert(sizeof(Tail) <= kTailBytes, "Tail does not fit");
    StreamArgs a;
    const uint64_t pitch = (uint64_t)w * 8;
    PANO_TRY(pano_make_tensor_map_2d(&a.m_b_halo, b, 8, w, h, pitch, BW, BH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_r_halo, r, 8, w, h, pitch, BW, BH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_r_int, r, 8, w, h, pitch, TW, TH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_s0_halo, s0, 8, w, h, pitch, BW, BH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_s1_halo, s1, 8, w, h, pitch, BW, BH));
    PANO_TRY(pano_make_tensor_map_2d(&a.m_x_int, x, 8, w, h, pitch, TW, TH));
    a.x = x; a.b = b; a.r = r; a.s0 = s0; a.s1 = s1;
    a.h = (int)h; a.w = (int)w;
    a.dt = timestep; a.threshold = threshold; a.max_iter = max_iterations;
    a.m = m;
    a.tiles_x = ((int)w + TW - 1) / TW;
    a.tiles_y = ((int)h + TH - 1) / TH;
    a.ctl = ctx->d_cg;
    a.zigzag = pano_option(ctx, "cg_zigzag", 1) != 0;
    if (!ctx->d_units) PANO_CUDA(cudaMalloc(&ctx->d_units, 2 * 3 * kMaxCtas * sizeof(ReduceUnit)));
    if (ctx->launch_epoch == 0) PANO_CUDA(cudaMemsetAsync(ctx->d_units, 0, 2 * 3 * kMaxCtas * sizeof(ReduceUnit), ctx->stream));
    a.units = (ReduceUnit *)ctx->d_units;
    a.seq_base = (++ctx->launch_epoch) << 32;
    int G = ctx->num_sms;
    const int ntiles = a.tiles_x * a.tiles_y;
    if (G > ntiles) G = ntiles;
    if (G > kMaxCtas) G = kMaxCtas;
    PANO_CUDA(cudaMemsetAsync(ctx->d_cg, 0, sizeof(PanoCgControl), ctx->stream));
    void *kargs[] = {(void *)&a};
    PANO_CUDA(cudaLaunchCooperativeKernel((const void *)k_cg_stream, dim3((unsigned)G), dim3(kThreads), kargs, kSmemBytes, ctx->stream));
    return pano_after_launch(ctx, "cg_stream");
}
