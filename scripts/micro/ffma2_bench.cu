// Microbenchmark: FFMA vs FFMA2 (packed f32x2, sm_100a) issue throughput per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float a[8]; unsigned long long p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = pk(a[i], a[i] + 1); }
    unsigned long long ps = pk(s, s), pc = pk(0.5f, 0.25f);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], s, 0.5f);
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], ps, pc);
        }
    }
    float r = 0; for (int i = 0; i < 8; ++i) { r += a[i]; r += __uint_as_float((unsigned)(p[i] & 0xffffffffu)); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 4, 512>>>(out, iters, 0.999f); else k<1><<<148 * 4, 512>>>(out, iters, 0.999f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double inst = 148.0 * 4 * 512 * (double)iters * 32;  // thread-instructions
        printf("%s: %.3f ms, %.1f Gthread-inst/s, %.1f lanes/clk/SM @1.965GHz, %.2f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms,
               inst / ms / 1e6, inst / (ms * 1e-3) / 148 / 1.965e9, inst * (mode ? 4 : 2) / ms / 1e9);
    }
    return 0;
}
