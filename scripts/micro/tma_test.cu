// Minimal TMA 3-D box load test: variant selected by argv[1].
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

#ifndef BW
#define BW 132
#endif
#ifndef BH
#define BH 20
#endif
#ifndef BC
#define BC 3
#endif

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, int c0, int c1, int c2,
                                            unsigned long long* bar) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(d), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(b)
        : "memory");
}

__global__ void k(const __grid_constant__ CUtensorMap tmap, float* out, int x0, int y0, int z0) {
    __shared__ __align__(128) float sm[BC * BH * BW];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, BC * BH * BW * 4);
        tma_load_3d(sm, &tmap, x0, y0, z0, &bar);
    }
    mbar_wait(&bar, 0);
    __syncthreads();
    for (int i = threadIdx.x; i < BC * BH * BW; i += blockDim.x) out[i] = sm[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    int W = 512, H = 512, P = 6;
    if (variant == 1) { W = 24; H = 20; }
    float* h = (float*)malloc(sizeof(float) * W * H * P);
    for (int i = 0; i < W * H * P; ++i) h[i] = (float)(i % 9973);
    float *d, *o;
    cudaMalloc(&d, sizeof(float) * W * H * P);
    cudaMalloc(&o, sizeof(float) * BC * BH * BW);
    cudaMemcpy(d, h, sizeof(float) * W * H * P, cudaMemcpyHostToDevice);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: err=%d q=%d p=%p\n", (int)e, (int)q, p);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P};
    cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {BW, BH, BC};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, variant == 2 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    int x0 = (variant == 3) ? 124 : -4, y0 = (variant == 3) ? 14 : -2, z0 = 3; if (variant == 4) { x0 = 126; }
    k<<<1, 256>>>(map, o, x0, y0, z0);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    float* ho = (float*)malloc(sizeof(float) * BC * BH * BW);
    cudaMemcpy(ho, o, sizeof(float) * BC * BH * BW, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < BC; ++c) for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) {
        int gx = x0 + x, gy = y0 + y, gz = z0 + c;
        float want = (gx < 0 || gy < 0 || gx >= W || gy >= H || gz >= P) ? 0.f : h[((size_t)gz * H + gy) * W + gx];
        if (ho[(c * BH + y) * BW + x] != want) { if (bad < 5) printf("mismatch c%d y%d x%d got %f want %f\n", c, y, x, ho[(c * BH + y) * BW + x], want); ++bad; }
    }
    printf("variant %d: %d mismatches\n", variant, bad);
    return bad != 0;
}
