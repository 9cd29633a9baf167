// Microbenchmark: latency of the grid-wide publish+poll all-reduce, in several flavours.
//   mode 0: st.volatile / ld.volatile 16-byte units (what the kernels use)
//   mode 1: st.relaxed.gpu / ld.relaxed.gpu 16-byte units
//   mode 2: atomicAdd counter barrier + partials array (the classic)
//   mode 3: mode 1 but every warp (not only warp 0..) polls a slice and no CTA-level reduce (pure exchange)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
struct __align__(16) Unit { double v; unsigned long long seq; };
__device__ __forceinline__ void st_vol(Unit *u, double v, unsigned long long s) { asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(u), "l"(__double_as_longlong(v)), "l"(s) : "memory"); }
__device__ __forceinline__ void ld_vol(const Unit *u, double &v, unsigned long long &s) { long long b; asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(b), "=l"(s) : "l"(u) : "memory"); v = __longlong_as_double(b); }
__device__ __forceinline__ void st_rlx(Unit *u, double v, unsigned long long s) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(u), "l"(__double_as_longlong(v)), "l"(s) : "memory"); }
__device__ __forceinline__ void ld_rlx(const Unit *u, double &v, unsigned long long &s) { long long b; asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(b), "=l"(s) : "l"(u) : "memory"); v = __longlong_as_double(b); }
__device__ __forceinline__ double warp_sum(double v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(Unit *units, unsigned long long *counter, double *partials, int iters, double *out, long long *cycles) {
    __shared__ double vals[256]; __shared__ double res;
    const int G = gridDim.x, tid = threadIdx.x;
    double acc = 1.0 + blockIdx.x;
    long long t0 = clock64();
    for (int n = 1; n <= iters; ++n) {
        Unit *bank = units + (n & 1) * 256;
        if (MODE == 0 || MODE == 1 || MODE == 3) {
            if (tid == 0) { if (MODE == 0) st_vol(bank + blockIdx.x, acc, n); else st_rlx(bank + blockIdx.x, acc, n); }
            if (tid < G) {
                double v; unsigned long long s;
                do { if (MODE == 0) ld_vol(bank + tid, v, s); else ld_rlx(bank + tid, v, s); } while (s != (unsigned long long)n);
                vals[tid] = v;
            }
            __syncthreads();
            if (tid < 32) { double a = 0; for (int i = tid; i < G; i += 32) a += vals[i]; a = warp_sum(a); if (tid == 0) res = a; }
            __syncthreads();
            acc = res * 1e-3 + blockIdx.x;
        } else if (MODE == 4) {
            // root protocol: everybody publishes, CTA 0 polls + reduces + publishes the result, the others poll the result
            Unit *resu = units + 512 + (n & 1);
            if (tid == 0) st_vol(bank + blockIdx.x, acc, n);
            if (blockIdx.x == 0) {
                if (tid < G) { double v; unsigned long long s; do { ld_vol(bank + tid, v, s); } while (s != (unsigned long long)n); vals[tid] = v; }
                __syncthreads();
                if (tid < 32) { double a = 0; for (int i = tid; i < G; i += 32) a += vals[i]; a = warp_sum(a); if (tid == 0) { res = a; st_vol(resu, a, n); } }
            } else {
                if (tid == 0) { double v; unsigned long long s; do { ld_vol(resu, v, s); } while (s != (unsigned long long)n); res = v; }
            }
            __syncthreads();
            acc = res * 1e-3 + blockIdx.x;
        } else {
            if (tid == 0) { partials[(n & 1) * 256 + blockIdx.x] = acc; __threadfence(); atomicAdd(counter, 1ULL);
                while (*(volatile unsigned long long *)counter < (unsigned long long)n * G) {} __threadfence(); }
            __syncthreads();
            if (tid < 32) { double a = 0; for (int i = tid; i < G; i += 32) a += __ldcg(partials + (n & 1) * 256 + i); a = warp_sum(a); if (tid == 0) res = a; }
            __syncthreads();
            acc = res * 1e-3 + blockIdx.x;
        }
    }
    long long t1 = clock64();
    if (tid == 0) { out[blockIdx.x] = acc; cycles[blockIdx.x] = t1 - t0; }
}
int main(int argc, char **argv) {
    int iters = 2000;
    Unit *units; unsigned long long *counter; double *partials, *out; long long *cycles;
    cudaMalloc(&units, 520 * sizeof(Unit)); cudaMalloc(&counter, 8); cudaMalloc(&partials, 512 * 8); cudaMalloc(&out, 256 * 8); cudaMalloc(&cycles, 256 * 8);
    for (int G : {8, 32, 64, 148}) for (int mode : {0, 2, 4}) {
        cudaMemset(units, 0, 520 * sizeof(Unit)); cudaMemset(counter, 0, 8);
        void *args[] = {&units, &counter, &partials, &iters, &out, &cycles};
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const void *fn = mode == 0 ? (const void *)k<0> : mode == 4 ? (const void *)k<4> : (const void *)k<2>;
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(G), dim3(512), args, 0, 0);
        cudaEventRecord(e1); cudaError_t e2 = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long hc[256]; cudaMemcpy(hc, cycles, G * 8, cudaMemcpyDeviceToHost);
        printf("G=%3d mode=%d: %s %s  %.3f us per all-reduce (events), %.0f cycles per all-reduce (clock64, CTA 0)\n", G, mode, cudaGetErrorString(e), cudaGetErrorString(e2), ms * 1e3 / iters, (double)hc[0] / iters);
    }
    return 0;
}
