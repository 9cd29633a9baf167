// pano_sm100.cuh -- sm_100a building blocks shared by the persistent kernels:
//   * mbarrier + TMA (cp.async.bulk.tensor) wrappers, written as inline PTX
//   * bounded waits (a stuck wait raises an error flag instead of hanging the GPU)
//   * the grid-wide "publish + poll" all-reduce that replaces counter barriers
#pragma once

#include <cuda.h>   // CUtensorMap (types only; the encode entry point is fetched at run time)

#include "pano_internal.cuh"

namespace pano_sm100 {

constexpr long long kSpinLimit = 1LL << 27;   // polls before a wait gives up (seconds), never reached in a healthy run

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Returns false if the bounded wait expired or somebody else raised the error flag.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, volatile unsigned int *err) {
    long long spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > kSpinLimit || ((spins & 0x3ff) == 0 && *err)) {
            atomicCAS((unsigned int *)err, 0u, 1u);   // keep the first error code (View2W raises 2)
            return false;
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// 2-D tiled load global -> shared, completion signalled on `bar` (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
// orders generic-proxy accesses (ld/st) against async-proxy accesses (TMA) of the same memory
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------- grid all-reduce
// One 16-byte unit per (value, CTA): {double value, u64 sequence}.  A CTA publishes its partials
// with single 16-byte stores; every CTA polls every unit until the sequence matches, then reduces
// the values in a fixed order, so all CTAs obtain bit-identical results.  Publishing, the barrier
// and the data movement are one L2 round trip (no atomics, no second read of a partials array).
struct __align__(16) ReduceUnit {
    double v;
    unsigned long long seq;
};

__device__ __forceinline__ void unit_store(ReduceUnit *u, double v, unsigned long long seq) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(u), "l"(__double_as_longlong(v)), "l"(seq) : "memory");
}
__device__ __forceinline__ void unit_load(const ReduceUnit *u, double &v, unsigned long long &seq) {
    long long bits;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(bits), "=l"(seq) : "l"(u) : "memory");
    v = __longlong_as_double(bits);
}
// polls one unit; false on timeout / raised error flag
__device__ __forceinline__ bool unit_poll(const ReduceUnit *u, unsigned long long seq, double &v, volatile unsigned int *err) {
    unsigned long long got;
    long long spins = 0;
    for (;;) {
        unit_load(u, v, got);
        if (got == seq) return true;
        if (++spins > kSpinLimit || ((spins & 0x3ff) == 0 && *err)) {
            atomicCAS((unsigned int *)err, 0u, 1u);   // keep the first error code (View2W raises 2)
            return false;
        }
    }
}

// up to three units polled CONCURRENTLY (their loads are in flight together: one L2 round trip instead of three; the reductions
// of the single-reduction CG kernels carry three values).  u[k] must be valid for k < n.
__device__ __forceinline__ bool unit_poll_n(const ReduceUnit *const *u, int n, unsigned long long seq, double *v, volatile unsigned int *err) {
    long long spins = 0;
    unsigned pending = (1u << n) - 1u;
    for (;;) {
        double t[3];
        unsigned long long got[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k < n && ((pending >> k) & 1u)) unit_load(u[k], t[k], got[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k < n && ((pending >> k) & 1u) && got[k] == seq) { v[k] = t[k]; pending &= ~(1u << k); }
        if (pending == 0u) return true;
        if (++spins > kSpinLimit || ((spins & 0x3ff) == 0 && *err)) {
            atomicCAS((unsigned int *)err, 0u, 1u);
            return false;
        }
    }
}

// the same poll with ld.acquire.sys: the poller synchronises with a releasing writer on another GPU without paying a
// system-scope fence afterwards (SASS: LD + CCTL.IVALL; MEMBAR.SYS costs ~3.1 us on a B200 NVLink system)
__device__ __forceinline__ bool unit_poll_acquire_sys(const ReduceUnit *u, unsigned long long seq, volatile unsigned int *err) {
    unsigned long long got, bits;
    long long spins = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(bits), "=l"(got) : "l"(u) : "memory");
        if (got == seq) return true;
        if (++spins > kSpinLimit || ((spins & 0x3ff) == 0 && *err)) {
            atomicCAS((unsigned int *)err, 0u, 1u);   // keep the first error code (View2W raises 2)
            return false;
        }
    }
}

// fixed-order sum / max of n values held in shared memory, executed by one full warp
__device__ __forceinline__ double warp_fixed_sum(const double *vals, int n, int lane) {
    double a = 0;
    for (int i = lane; i < n; i += 32) a += vals[i];
    return warp_sum(a);
}
__device__ __forceinline__ double warp_fixed_max(const double *vals, int n, int lane) {
    double a = 0;
    for (int i = lane; i < n; i += 32) a = vals[i] > a ? vals[i] : a;
    return warp_max(a);
}

// ---------------------------------------------------------------------------------- the all-reduce
// Grid-wide, deterministic all-reduce of up to three per-CTA values (sum or max per value).
//   G <= 32 : all-to-all -- every CTA polls every CTA's unit (one L2 round trip, 1.2 us at G = 32)
//   G  > 32 : root       -- CTA 0 polls all partials, reduces them in a fixed order and publishes the
//             result; the others poll that one unit (two round trips but no hot-spotting:
//             measured 1.43 us at G = 148 against 3.5 us all-to-all and 2.5 us for an atomic counter)
//   G  > 32, `inbox` given, one GPU : push -- every CTA stores its partials into every CTA's private inbox and
//             polls its own (one round trip, no shared polled lines); see the branch below
// Every CTA ends up with bit-identical results.  `fenced`: bulk data written with ordinary stores
// must be visible to the other CTAs afterwards (streaming kernel); the SM-resident kernel moves
// everything through units and needs no fence.
constexpr int kMaxCtas = 192;
constexpr int kUnitsPerBank = 4 * kMaxCtas;       // 3 x kMaxCtas partials + results
constexpr int kUnitsTotal = 2 * kUnitsPerBank;    // double-buffered by exchange parity
// "push" exchange (grids above 32 CTAs, one GPU): [2 banks][reader CTA][3 values][writer CTA] units, 3.4 MB
constexpr size_t kInboxUnits = (size_t)2 * kMaxCtas * 3 * kMaxCtas;
static_assert(kInboxUnits * 16 == kPanoInboxBytes, "inbox size (pano_internal.cuh)");

// Multi-GPU extension: after CTA 0 has reduced its GPU's partials it writes the GPU total into a slot of
// EVERY rank's cross-rank unit array (peer-mapped memory, st.volatile = system scope, over NVLink),
// polls its own array until all ranks have written, and combines them in a fixed lane = rank order, so
// every GPU obtains the bit-identical value.  One NVLink store latency per reduction, no NCCL call.
constexpr int kMaxRanks = 8;
// Four banks over (launch & 1, exchange & 1): between two solver launches the ranks are ordered only by nearest-neighbour
// halo exchanges, so a fast rank may store its total of (step+1, n=0) while a rank several hops away has not yet read the
// value of (step, n=0) -- the launch bit keeps those two apart (seq_base carries the launch number in its high word).
constexpr int kXBanks = 4;
constexpr int kXUnitsTotal = kXBanks * kMaxRanks * 3;   // [bank][rank][value]
// "Halo flags" (XRank::hflags, option cg_xflags): the cross-GPU exchange above needs a system-scope fence on either side
// of it in the root CTA, because the root vouches for the halo rows OTHER CTAs stored into the neighbours' memory
// (fence cumulativity) -- measured 2 x 3.1 us per reduction (scripts/probe/xgpu_probe.cu).  With halo flags every CTA
// vouches for itself: after its own system-scope fence it stores a {0, sequence} unit into a per-CTA slot in each
// neighbour's memory, and the neighbour's root polls those slots with ld.acquire.sys.  The roots then exchange the GPU
// totals with plain volatile stores and polls, no fence around them.  [bank][from up / from down][CTA] units.
constexpr int kXFlagUnits = kXBanks * 2 * kMaxCtas;
struct XRank;
__device__ __forceinline__ int xbank_of(const XRank *xr, unsigned long long n);
struct XRank {
    int rank, nranks;
    unsigned long long seq_base;      // identical on all ranks (incremented once per collective launch)
    ReduceUnit *local;                // this rank's cross-rank unit array
    ReduceUnit *peer[kMaxRanks];      // every rank's array as mapped into this process (peer[rank] == local)
    ReduceUnit *hflags;               // this rank's halo-flag array (written by the neighbours' CTAs), or null: fenced exchange
    ReduceUnit *hflags_up, *hflags_dn;   // the neighbours' arrays as mapped here, or null at the domain walls
    int g_up, g_dn;                   // CTAs of the upper / lower neighbour's kernel (flags to wait for), 0 at the walls
};

__device__ __forceinline__ int xbank_of(const XRank *xr, unsigned long long n) {
    return (int)(n & 1) + 2 * (int)((xr->seq_base >> 32) & 1);
}

// Fences of the `fenced` exchange.  Every use is a release (before publishing a unit) or an acquire (after polling
// one) around relaxed volatile accesses, for which PTX's fence.acq_rel is sufficient; __threadfence*() emit the
// sequentially consistent fence.sc (SASS MEMBAR.SC.* instead of MEMBAR.ALL.*).  `light` selects fence.acq_rel.
__device__ __forceinline__ void fence_gpu(bool light) {
    if (light) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    else __threadfence();
}
__device__ __forceinline__ void fence_sys(bool light) {
    if (light) asm volatile("fence.acq_rel.sys;" ::: "memory");
    else __threadfence_system();
}
constexpr int kFenceLight = 1;        // fence.acq_rel instead of fence.sc
constexpr int kFenceSysIfRemote = 2;  // multi-GPU: system scope only in CTAs that stored into a peer's memory this phase
constexpr int kFenceRootGpuOnly = 4;  // DIAGNOSTIC (formally a race): the root's two fences around the cross-GPU exchange at gpu scope

// One thread, after a fence that covers the CTA's stores into the neighbours' memory: this CTA's halo flags of exchange n.
__device__ __forceinline__ void send_halo_flags(const XRank *xr, unsigned long long n) {
    const unsigned long long xseq = xr->seq_base + n;
    const int xb = xbank_of(xr, n);
    if (xr->hflags_up) unit_store(xr->hflags_up + (xb * 2 + 1) * kMaxCtas + blockIdx.x, 0.0, xseq);
    if (xr->hflags_dn) unit_store(xr->hflags_dn + (xb * 2 + 0) * kMaxCtas + blockIdx.x, 0.0, xseq);
}

struct NoWork {
    __device__ __forceinline__ void operator()() const {}
};

// `mid` is work that does not depend on the reduction's result (e.g. the x update of CG): a CTA that only waits
// runs it between publishing its partials and polling the result; the root, which everybody waits for, runs it
// after it has published the result, in the head start it has over the others.
template <class SyncFn, class MidFn = NoWork>
__device__ __forceinline__ bool grid_allreduce_units(ReduceUnit *units, unsigned long long seq, unsigned long long n, int nvals,
                                                     double v0, double v1, double v2, unsigned max_mask,
                                                     double (*vals)[kMaxCtas], double *out_sh, int *ok_sh,
                                                     volatile unsigned int *err, bool fenced, SyncFn sync, double *out,
                                                     const XRank *xr = nullptr, MidFn mid = MidFn(), ReduceUnit *inbox = nullptr,
                                                     int fmode = 0, bool remote = true, bool flags_sent = false) {
    const int G = gridDim.x, tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const unsigned parity = (unsigned)(n & 1);
    const bool multi = xr != nullptr && xr->nranks > 1;
    ReduceUnit *bank = units + parity * kUnitsPerBank;
    ReduceUnit *result = bank + 3 * kMaxCtas;
    const bool light = (fmode & kFenceLight) != 0;
    if (fenced && tid == 0) {
        // halo rows stored into the neighbours' memory must have landed; a CTA that stored none only publishes
        // tiles that its own GPU reads
        if (multi && (remote || !(fmode & kFenceSysIfRemote))) fence_sys(light);
        else fence_gpu(light);
        // this CTA's halo rows of the ending phase have landed: tell the neighbours' roots (unless send_halo_flags
        // already ran, right after the CTA's last halo tile of the phase)
        if (multi && xr->hflags && !flags_sent) send_halo_flags(xr, n);
        fence_proxy_async();
    }
    if (fenced) sync();   // the publishing threads below must not run ahead of thread 0's fence
    if (inbox != nullptr && !multi && G > 32) {
        // push: thread t hands this CTA's partials (v0..v2 must hold the CTA totals in EVERY thread) to CTA t's
        // private inbox, then polls the unit CTA t pushed into this CTA's inbox.  Nobody shares a polled line with
        // another reader, so the exchange is one store + one poll; all CTAs then reduce the same values in the same
        // order as the root would.  Two banks suffice: a CTA can be at most one exchange ahead of any other, because
        // it needs everybody's units of exchange n+1 before it can start n+2.
        if (tid < G) {
            ReduceUnit *dst = inbox + ((size_t)(parity * kMaxCtas + (unsigned)tid) * 3) * kMaxCtas + blockIdx.x;
            unit_store(dst, v0, seq);
            if (nvals > 1) unit_store(dst + kMaxCtas, v1, seq);
            if (nvals > 2) unit_store(dst + 2 * kMaxCtas, v2, seq);
        }
        mid();
        if (tid < G) {
            const ReduceUnit *src = inbox + ((size_t)(parity * kMaxCtas + blockIdx.x) * 3) * kMaxCtas + tid;
            bool ok = true;
            for (int k = 0; k < nvals && ok; ++k) {
                double v;
                ok = unit_poll(src + k * kMaxCtas, seq, v, err);
                vals[k][tid] = v;
            }
            if (fenced) fence_gpu(light);
            if (!ok) *ok_sh = 0;
        }
        sync();
        if (wid < nvals) {
            const bool is_max = (max_mask >> wid) & 1u;
            const double r = is_max ? warp_fixed_max(vals[wid], G, lane) : warp_fixed_sum(vals[wid], G, lane);
            if (lane == 0) out_sh[wid] = r;
        }
        sync();
        out[0] = out_sh[0];
        if (nvals > 1) out[1] = out_sh[1];
        if (nvals > 2) out[2] = out_sh[2];
        return *ok_sh != 0;
    }
    if (tid < nvals) unit_store(bank + tid * kMaxCtas + blockIdx.x, tid == 0 ? v0 : (tid == 1 ? v1 : v2), seq);
    const bool root_mode = G > 32 || multi;
    if (!root_mode || blockIdx.x == 0) {
        if (tid < G) {
            const ReduceUnit *us[3] = {bank + tid, bank + kMaxCtas + tid, bank + 2 * kMaxCtas + tid};
            double v[3] = {0.0, 0.0, 0.0};
            const bool ok = unit_poll_n(us, nvals, seq, v, err);   // the values' units in flight together
            for (int k = 0; k < nvals; ++k) vals[k][tid] = v[k];
            if (fenced) fence_gpu(light);
            if (!ok) *ok_sh = 0;
        }
        sync();
        if (multi && xr->hflags && tid == 32 * 7) {
            // "local done": every CTA of THIS GPU has published (hence fenced) its part of the ending pass.  The producers of
            // k_cg_sr start streaming the tiles that do not touch another GPU's rows on this, while the totals still travel.
            fence_gpu(light);
            unit_store(result + 3, 0.0, seq);
        }
        if (wid < nvals) {
            const bool is_max = (max_mask >> wid) & 1u;
            double r = is_max ? warp_fixed_max(vals[wid], G, lane) : warp_fixed_sum(vals[wid], G, lane);
            if (multi) {   // cross-rank stage (root CTA only, one warp per value)
                const unsigned long long xseq = xr->seq_base + n;
                const int xb = xbank_of(xr, n);
                const int slot = (xb * kMaxRanks + xr->rank) * 3 + wid;
                const bool root_sys = !(fmode & kFenceRootGpuOnly), fence_here = fenced && xr->hflags == nullptr;
                if (fence_here) { if (root_sys) fence_sys(light); else fence_gpu(light); }
                if (lane < xr->nranks) unit_store(xr->peer[lane] + slot, r, xseq);
                double v = 0.0;
                if (lane < xr->nranks && !unit_poll(xr->local + (xb * kMaxRanks + lane) * 3 + wid, xseq, v, err)) *ok_sh = 0;
                if (fence_here) { if (root_sys) fence_sys(light); else fence_gpu(light); }
                r = is_max ? warp_max(v) : warp_sum(v);   // fixed butterfly over lane = rank: same bits on every GPU
            }
            if (lane == 0) {
                out_sh[wid] = r;
                if (root_mode && !(multi && xr->hflags)) unit_store(result + wid, r, seq);
            }
        } else if (multi && xr->hflags && wid < nvals + 4) {
            // four more warps: the halo flags of both neighbours' CTAs (acquire: their rows are visible to this GPU now)
            const unsigned long long xseq = xr->seq_base + n;
            const int xb = xbank_of(xr, n);
            const ReduceUnit *from_up = xr->hflags + (xb * 2 + 0) * kMaxCtas, *from_dn = xr->hflags + (xb * 2 + 1) * kMaxCtas;
            for (int t = tid - 32 * nvals; t < xr->g_up + xr->g_dn; t += 128)
                if (!unit_poll_acquire_sys(t < xr->g_up ? from_up + t : from_dn + (t - xr->g_up), xseq, err)) *ok_sh = 0;
        }
        if (multi && xr->hflags) {   // the result goes out once the totals AND the flags are in
            sync();
            if (wid < nvals && lane == 0) {
                fence_gpu(light);    // release, cumulative over what the flag pollers acquired (ordered by the barrier above)
                unit_store(result + wid, out_sh[wid], seq);
            }
        }
        mid();
    } else {
        mid();
        if (tid < nvals) {
            double v;
            if (!unit_poll(result + tid, seq, v, err)) *ok_sh = 0;
            if (fenced) fence_gpu(light);
            out_sh[tid] = v;
        }
    }
    sync();
    out[0] = out_sh[0];
    if (nvals > 1) out[1] = out_sh[1];
    if (nvals > 2) out[2] = out_sh[2];
    return *ok_sh != 0;
}

}  // namespace pano_sm100

// host side: fetch cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
int pano_make_tensor_map_2d(CUtensorMap *map, const void *base, size_t elem_bytes, uint64_t width, uint64_t height,
                            uint64_t row_pitch_bytes, uint32_t box_w, uint32_t box_h);
