// Stand-alone probe: one TMA 2-D f64 halo-box load, issued (mode 0) by a lone lane whose siblings exited,
// (mode 1) by lane 0 of a full warp that stays alive.  Prints a checksum of the box.
#include <cstdio>
#include <cstdlib>
#include "../../panopaea_b200/csrc/pano_sm100.cuh"
using namespace pano_sm100;
void pano_set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
constexpr int BW = 68, BH = 34;
struct Args { CUtensorMap map; double *out; int x0, y0, mode; };
__global__ void __launch_bounds__(288, 1) probe(const __grid_constant__ Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 18560);
    const int tid = threadIdx.x, wid = tid >> 5;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (wid == 8) {
        if (a.mode == 0) {
            if ((tid & 31) != 0) return;
            mbar_arrive_expect_tx(bar, BW * BH * 8);
            tma_load_2d(smem, &a.map, bar, a.x0, a.y0);
        } else {
            if ((tid & 31) == 0) {
                mbar_arrive_expect_tx(bar, BW * BH * 8);
                tma_load_2d(smem, &a.map, bar, a.x0, a.y0);
            }
            __syncwarp();
        }
        return;
    }
    unsigned int err = 0;
    volatile unsigned int *perr = &err;
    while (!mbar_try_wait(bar, 0)) {}
    const double *S = reinterpret_cast<const double *>(smem);
    for (int i = tid; i < BW * BH; i += 256) a.out[i] = S[i];
    (void)perr;
}
int pano_make_tensor_map_2d(CUtensorMap *map, const void *base, size_t elem_bytes, uint64_t width, uint64_t height, uint64_t row_pitch_bytes, uint32_t box_w, uint32_t box_h) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    const cuuint64_t dims[2] = {width, height}; const cuuint64_t strides[1] = {row_pitch_bytes};
    const cuuint32_t box[2] = {box_w, box_h}; const cuuint32_t estr[2] = {1, 1};
    CUresult rc = ((EncodeFn)fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)rc); return rc;
}
int main(int argc, char **argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 0, w = argc > 2 ? atoi(argv[2]) : 256, h = argc > 3 ? atoi(argv[3]) : 128;
    int x0 = argc > 4 ? atoi(argv[4]) : -1, y0 = argc > 5 ? atoi(argv[5]) : -1;
    double *d, *out; cudaMalloc(&d, (size_t)w * h * 8 + 256); cudaMalloc(&out, BW * BH * 8);
    double *hbuf = (double *)malloc((size_t)w * h * 8);
    for (int i = 0; i < w * h; ++i) hbuf[i] = 1.0 + i;
    cudaMemcpy(d, hbuf, (size_t)w * h * 8, cudaMemcpyHostToDevice);
    Args a; a.out = out; a.x0 = x0; a.y0 = y0; a.mode = mode;
    if (pano_make_tensor_map_2d(&a.map, d, 8, w, h, (uint64_t)w * 8, BW, BH)) return 1;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    probe<<<1, 288, 32768>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d sync: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    double ho[BW * BH]; cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
    double want = 0, got = 0; int bad = 0;
    for (int r = 0; r < BH; ++r) for (int c = 0; c < BW; ++c) {
        int gy = y0 + r, gx = x0 + c; double v = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? 1.0 + (double)gy * w + gx : 0.0;
        want += v; got += ho[r * BW + c]; if (ho[r * BW + c] != v) ++bad;
    }
    printf("mode %d box(%d,%d): want %.1f got %.1f mismatches %d\n", mode, x0, y0, want, got, bad);
    return bad != 0;
}
