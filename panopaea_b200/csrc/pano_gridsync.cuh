// pano_gridsync.cuh -- grid-wide barrier and deterministic partial reductions shared by the cooperative CG kernels
// (pano_cg.cu: any 2-D shape / dtype; pano_grid3.cu: the 7-point solve).
#pragma once

#include "pano_internal.cuh"

namespace {

constexpr long long kSpinLimit = 40LL * 1000 * 1000;   // bounded wait: a few seconds, then PANO_ERR_TIMEOUT

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Grid-wide barrier on a monotonically increasing arrival counter.  Returns false when the
// bounded wait expired somewhere (ctl->error set): every CTA then leaves the kernel.
__device__ __forceinline__ bool grid_barrier(PanoCgControl *ctl, unsigned long long target, int *s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&ctl->barrier, 1ULL);
        long long spins = 0;
        int ok = 1;
        while (ld_acquire_u64(&ctl->barrier) < target) {
            if (*(volatile unsigned int *)&ctl->error) { ok = 0; break; }
            if (++spins > kSpinLimit) {
                atomicCAS(&ctl->error, 0u, 1u);
                ok = 0;
                break;
            }
        }
        __threadfence();
        if (*(volatile unsigned int *)&ctl->error) ok = 0;
        *s_flag = ok;
    }
    __syncthreads();
    return *s_flag != 0;
}

template <class T>
__device__ __forceinline__ T sum_partials(const double *p, int n, T *scratch) {
    T acc = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (T)__ldcg(p + i);
    return block_sum(acc, scratch);
}
template <class T>
__device__ __forceinline__ T max_partials(const double *p, int n, T *scratch) {
    T acc = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        T v = (T)__ldcg(p + i);
        acc = v > acc ? v : acc;
    }
    return block_max(acc, scratch);
}

}  // namespace
