// Microbenchmark: latency of the cross-GPU stage of the CG kernel's all-reduce (pano_sm100.cuh), isolated.
// One process, NG GPUs with peer access; one warp per GPU plays the root CTA: lane r stores a 16-byte {value, seq} unit
// into GPU r's inbox (st.volatile over NVLink), polls the own inbox until every rank's unit of this round arrived, and
// repeats.  Variants (what surrounds the exchange):
//   0: nothing                     1: fence.acq_rel.sys before the store and after the poll
//   2: __threadfence_system() x2   3: fence.acq_rel.sys before the store only, ld.acquire.sys polls
//   4: 1 + each round first stores 8 KB into the peer (a halo row), i.e. the fence has posted writes to wait for
// usage: xgpu_probe [ngpus=2] [rounds=2000]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
struct __align__(16) Unit { double v; unsigned long long seq; };
struct Args { Unit *inbox[8]; double *row[8]; int rank, n; int rounds; int mode; long long *cycles; double *out; };
__device__ __forceinline__ void st_vol(Unit *u, double v, unsigned long long s) { asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(u), "l"(__double_as_longlong(v)), "l"(s) : "memory"); }
__device__ __forceinline__ void ld_vol(const Unit *u, double &v, unsigned long long &s) { long long b; asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(b), "=l"(s) : "l"(u) : "memory"); v = __longlong_as_double(b); }
__device__ __forceinline__ void ld_acq(const Unit *u, double &v, unsigned long long &s) { long long b; asm volatile("ld.acquire.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(b), "=l"(s) : "l"(u) : "memory"); v = __longlong_as_double(b); }
__global__ void k(Args a) {
    const int lane = threadIdx.x;
    double acc = 1.0 + a.rank;
    long long t0 = clock64();
    for (int n = 1; n <= a.rounds; ++n) {
        Unit *mine = a.inbox[a.rank] + (n & 1) * 8;
        if (a.mode == 4) {   // a halo row for the next rank, 8 KB
            double *dst = a.row[(a.rank + 1) % a.n];
            for (int i = lane; i < 1024; i += 32) dst[i] = acc + i;
        }
        if (a.mode == 1 || a.mode == 3 || a.mode == 4) asm volatile("fence.acq_rel.sys;" ::: "memory");
        if (a.mode == 2) __threadfence_system();
        if (lane < a.n) st_vol(a.inbox[lane] + (n & 1) * 8 + a.rank, acc, (unsigned long long)n);
        double v = 0; unsigned long long s = 0;
        if (lane < a.n) {
            long long spins = 0;
            do { if (a.mode == 3) ld_acq(mine + lane, v, s); else ld_vol(mine + lane, v, s); } while (s != (unsigned long long)n && ++spins < (1LL << 26));
        }
        if (a.mode == 1 || a.mode == 4) asm volatile("fence.acq_rel.sys;" ::: "memory");
        if (a.mode == 2) __threadfence_system();
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc = v * 1e-3 + a.rank;
    }
    long long t1 = clock64();
    if (lane == 0) { a.cycles[0] = t1 - t0; a.out[0] = acc; }
}
int main(int argc, char **argv) {
    int ng = argc > 1 ? atoi(argv[1]) : 2, rounds = argc > 2 ? atoi(argv[2]) : 2000, have = 0;
    CK(cudaGetDeviceCount(&have));
    if (have < ng) { printf("need %d GPUs, have %d\n", ng, have); return 0; }
    std::vector<Unit *> inbox(ng); std::vector<double *> row(ng); std::vector<long long *> cyc(ng); std::vector<double *> out(ng); std::vector<cudaStream_t> st(ng);
    for (int g = 0; g < ng; ++g) {
        CK(cudaSetDevice(g));
        for (int p = 0; p < ng; ++p) if (p != g) { cudaError_t e = cudaDeviceEnablePeerAccess(p, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { printf("no peer access %d->%d\n", g, p); return 0; } cudaGetLastError(); }
        CK(cudaMalloc(&inbox[g], 16 * sizeof(Unit))); CK(cudaMalloc(&row[g], 8192)); CK(cudaMalloc(&cyc[g], 8)); CK(cudaMalloc(&out[g], 8)); CK(cudaStreamCreate(&st[g]));
    }
    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    for (int mode = 0; mode <= 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            for (int g = 0; g < ng; ++g) { CK(cudaSetDevice(g)); CK(cudaMemset(inbox[g], 0, 16 * sizeof(Unit))); CK(cudaDeviceSynchronize()); }
            for (int g = 0; g < ng; ++g) {
                CK(cudaSetDevice(g));
                Args a; for (int p = 0; p < 8; ++p) { a.inbox[p] = p < ng ? inbox[p] : nullptr; a.row[p] = p < ng ? row[p] : nullptr; }
                a.rank = g; a.n = ng; a.rounds = rounds; a.mode = mode; a.cycles = cyc[g]; a.out = out[g];
                k<<<1, 32, 0, st[g]>>>(a);
                CK(cudaGetLastError());
            }
            for (int g = 0; g < ng; ++g) { CK(cudaSetDevice(g)); CK(cudaStreamSynchronize(st[g])); }
            if (rep == 1) {
                printf("mode %d:", mode);
                for (int g = 0; g < ng; ++g) { long long c; CK(cudaSetDevice(g)); CK(cudaMemcpy(&c, cyc[g], 8, cudaMemcpyDeviceToHost)); printf(" gpu%d %.2f us/round", g, (double)c / rounds / (clk * 1e-3)); }
                printf("  (%d GPUs, %d rounds, SM clock %d kHz nominal)\n", ng, rounds, clk);
            }
        }
    }
    return 0;
}
