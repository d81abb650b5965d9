// Shot / read noise of the synthetic-RAW model (isp/unprocess_np.py:145-181: random_noise_levels_* draw
// one (shot, read) pair per image on the host; add_read_and_shot_noise adds N(0, sqrt(x * shot + read)) per
// element) with the brightness scaling that precedes it (adjust_random_brightness :131-138) folded in:
//     out = gain * x + sqrt(gain * x * shot + read) * z,   z ~ N(0, 1).
// HBM-bound streaming kernel: 8 B per element when z is generated in the kernel (Philox4x32-10, one counter
// per group of four elements, Box-Muller), 12 B when the caller supplies z (parity runs: the result is then
// a deterministic function of its inputs).  A negative variance gives NaN, as numpy's sqrt does.
#include "aisp_common.cuh"

namespace aisp {

__device__ __forceinline__ uint2 mulhilo32(unsigned a, unsigned b) {
    const unsigned long long p = (unsigned long long)a * b;
    return make_uint2((unsigned)p, (unsigned)(p >> 32));
}

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint2 a = mulhilo32(0xD2511F53u, c.x), b = mulhilo32(0xCD9E8D57u, c.z);
        c = make_uint4(b.y ^ c.y ^ k.x, b.x, a.y ^ c.w ^ k.y, a.x);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ float2 box_muller(unsigned a, unsigned b) {
    const float u = ((float)a + 0.5f) * 2.3283064365386963e-10f;   // (0, 1)
    const float v = ((float)b + 0.5f) * 2.3283064365386963e-10f;
    const float rad = sqrtf(-2.0f * __logf(u));
    float s, c;
    __sincosf(6.283185307179586f * v, &s, &c);
    return make_float2(rad * c, rad * s);
}

template <bool HAVE_Z>
__global__ void __launch_bounds__(kThreads)
shot_read_noise_kernel(const float* __restrict__ img, const float* __restrict__ z, float* __restrict__ out,
                       const float* __restrict__ shot, const float* __restrict__ read, const float* __restrict__ gain,
                       long long n /* elements per image */, unsigned long long seed, unsigned long long offset) {
    pdl_prologue();
    const int b = blockIdx.y;
    const float sh = shot[b], rd = read[b], gn = gain ? gain[b] : 1.0f;
    const float* x = img + (size_t)b * n;
    float* o = out + (size_t)b * n;
    const long long groups = (n + 3) / 4;
    const bool vec = ((n & 3) == 0) && (((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(out) |
                                          reinterpret_cast<uintptr_t>(z)) & 15u) == 0);
    for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < groups; q += (long long)gridDim.x * kThreads) {
        float zz[4];
        if (HAVE_Z) {
            if (vec) {
                const float4 t = ldg_stream4(z + (size_t)b * n + 4 * q);
                zz[0] = t.x; zz[1] = t.y; zz[2] = t.z; zz[3] = t.w;
            } else {
#pragma unroll
                for (int v = 0; v < 4; ++v) zz[v] = (4 * q + v < n) ? z[(size_t)b * n + 4 * q + v] : 0.f;
            }
        } else {
            const unsigned long long ctr = offset + (unsigned long long)q;
            const uint4 r = philox4x32_10(make_uint4((unsigned)ctr, (unsigned)(ctr >> 32), (unsigned)b, 0u),
                                          make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
            const float2 g0 = box_muller(r.x, r.y), g1 = box_muller(r.z, r.w);
            zz[0] = g0.x; zz[1] = g0.y; zz[2] = g1.x; zz[3] = g1.y;
        }
        float xv[4];
        if (vec) {
            const float4 t = ldg_stream4(x + 4 * q);
            xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
        } else {
#pragma unroll
            for (int v = 0; v < 4; ++v) xv[v] = (4 * q + v < n) ? x[4 * q + v] : 0.f;
        }
        float y[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float s = xv[v] * gn;
            y[v] = s + sqrtf(s * sh + rd) * zz[v];
        }
        if (vec) {
            stg_stream4(o + 4 * q, make_float4(y[0], y[1], y[2], y[3]));
        } else {
#pragma unroll
            for (int v = 0; v < 4; ++v)
                if (4 * q + v < n) o[4 * q + v] = y[v];
        }
    }
}

cudaError_t launch_shot_read_noise(const float* img, const float* z, float* out, const float* shot, const float* read,
                                   const float* gain, int B, long long n, unsigned long long seed,
                                   unsigned long long offset, cudaStream_t st) {
    const long long groups = (n + 3) / 4;
    long long gx = (groups + kThreads - 1) / kThreads;
    if (gx > 148 * 8) gx = 148 * 8;   // grid-stride: eight CTAs per SM
    dim3 grid((unsigned)gx, (unsigned)B);
    if (z)
        launch_pdl(shot_read_noise_kernel<true>, grid, kThreads, st, img, z, out, shot, read, gain, n, seed, offset);
    else
        launch_pdl(shot_read_noise_kernel<false>, grid, kThreads, st, img, z, out, shot, read, gain, n, seed, offset);
    return cudaGetLastError();
}

}  // namespace aisp
