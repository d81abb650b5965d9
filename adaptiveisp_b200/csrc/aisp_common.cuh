// Shared device-side vocabulary of the sm_100a ISP kernels: op classification, per-step derived
// constants, streaming 128-bit loads/stores and the block-level gradient reduction.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aisp_b200.h"

#define AISP_LN2F 0.69314718055994530942f
#define AISP_PIF 3.14159265358979323846f

namespace aisp {

constexpr int kThreads = 256;          // every kernel here runs 8 warps per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kConst = 32;             // floats of derived constants per step (>= AISP_PSTRIDE + 3)

__host__ __device__ __forceinline__ bool is_pointwise(int op) {
    return op == AISP_OP_EXPOSURE || op == AISP_OP_GAMMA || op == AISP_OP_CCM || op == AISP_OP_TONE ||
           op == AISP_OP_CONTRAST || op == AISP_OP_SATPLUS || op == AISP_OP_WNB || op == AISP_OP_WB ||
           op == AISP_OP_COLOR;
}
__host__ __device__ __forceinline__ bool is_sharpen(int op) {
    return op == AISP_OP_SHARPEN || op == AISP_OP_SHARPEN_V2 || op == AISP_OP_USM;
}

// ---------------------------------------------------------------------------------------------
// streaming global access: every image byte is touched once per pass, so bypass L1 allocation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg_stream1(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <int VEC>
struct Pack;
template <>
struct Pack<4> {
    float v[4];
    __device__ __forceinline__ void load(const float* p) {
        float4 t = ldg_stream4(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float* p) const { stg_stream4(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <>
struct Pack<1> {
    float v[1];
    __device__ __forceinline__ void load(const float* p) { v[0] = ldg_stream1(p); }
    __device__ __forceinline__ void store(float* p) const { stg_stream1(p, v[0]); }
};

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): every kernel here is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and starts with griddepcontrol.wait, so its
// launch/scheduling latency overlaps the tail of the previous kernel on the stream; the wait
// returns only when the previous grid has completed and flushed, so ordering is unchanged.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_smem(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                   Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// cp.async (LDGSTS) with zero-fill: global -> shared without staging registers
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int n = valid ? 16 : 0;   // src-size 0: nothing is read, the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int n = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// torch.clip semantics: NaN stays NaN (fminf/fmaxf would swallow it)
__device__ __forceinline__ float clip01(float y) {
    // two NaN-propagating FMNMX instead of two compare+select pairs
    float r;
    asm("min.NaN.f32 %0, %1, 0f3F800000;" : "=f"(r) : "f"(y));
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(r));
    return r;
}
// torch.maximum / torch.minimum / Tensor.max(dim) semantics: NaN wins (fmaxf / fminf drop it)
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float min_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
// ---------------------------------------------------------------------------------------------
// Blackwell packed fp32 (PTX ISA 8.6, sm_100+; SASS FADD2 / FMUL2 / FFMA2): one instruction, one issue
// slot, two IEEE round-to-nearest fp32 results (lane-wise identical to the scalar instruction).  A pair
// lives in an aligned 64-bit register pair: 64 / 128-bit loads deliver pairs for free, scalar producers
// are packed for free when the register allocator places their destinations side by side.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 splat2(float v) { return pack2(v, v); }
__device__ __forceinline__ float lo2(f32x2 v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi2(f32x2 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// Sum / difference that must NOT be contracted with a multiply feeding it: ptxas (CUDA 12.9) fuses mul.rn.f32x2 +
// add.rn.f32x2 into FFMA2 despite the explicit .rn and -fmad=false (it honours both for the scalar forms), and it
// also folds a literal x * 1.0 + y back into that pattern.  Written as x * ONE + y with ONE read from constant memory
// (opaque to ptxas) the sum is the same correctly rounded value, costs the same one instruction (FFMA2 with a uniform
// broadcast operand), and has no free multiply to contract.  Checked in SASS: FMUL2, FMUL2, FFMA2 -- not FMUL2, FFMA2.
static __constant__ float kOpaqueOne[2] = {1.0f, -1.0f};
__device__ __forceinline__ f32x2 add2_sep(f32x2 x, f32x2 y) { return fma2(x, splat2(kOpaqueOne[0]), y); }
__device__ __forceinline__ f32x2 sub2_sep(f32x2 x, f32x2 y) { return fma2(y, splat2(kOpaqueOne[1]), x); }   // x - y
__device__ __forceinline__ f32x2 lds2(const float* p) {   // 8-byte aligned pair from shared memory
    return *reinterpret_cast<const f32x2*>(p);
}

// clamp backward: gradient passes iff lo <= y <= hi, inclusive (NaN -> 0)
__device__ __forceinline__ float pass01(float y) { return (__saturatef(y) == y) ? 1.f : 0.f; }   // see mask01

// ---------------------------------------------------------------------------------------------
// per-step derived constants (one thread per step computes them once per CTA / finalize block)
//   raw : AISP_PSTRIDE parameters as the regressor produced them
//   c   : kConst floats consumed by the per-pixel code
// ---------------------------------------------------------------------------------------------
__device__ inline void derive_consts(int op, const float* raw, float* c) {
    switch (op) {
    case AISP_OP_EXPOSURE:  // img * exp(p * ln2)            isp/filters.py:224
        c[0] = expf(raw[0] * AISP_LN2F);
        break;
    case AISP_OP_CCM: {  // rows / row-sum, no epsilon        isp/filters.py:706-707
        for (int i = 0; i < 3; ++i) {
            float s = (raw[3 * i] + raw[3 * i + 1]) + raw[3 * i + 2];
            for (int j = 0; j < 3; ++j) c[3 * i + j] = raw[3 * i + j] / s;
            c[9 + i] = 1.0f / s;
        }
        break;
    }
    case AISP_OP_TONE: {  // steps / (sum + 1e-30)            isp/filters.py:340,345
        float s = 0.f;
        for (int k = 0; k < 8; ++k) { c[k] = raw[k]; s += raw[k]; }
        c[8] = 8.0f / (s + 1e-30f);
        break;
    }
    case AISP_OP_COLOR: {  // per-channel curves              isp/filters.py:297,302
        // torch.sum over the strided knot dim runs 4 interleaved accumulators (k % 4) and combines
        // them left to right; same order here so that saturated pixels hit the same side of 1.0
        for (int k = 0; k < 24; ++k) c[k] = raw[k];
        for (int ch = 0; ch < 3; ++ch) {
            float a4[4];
            for (int j = 0; j < 4; ++j) a4[j] = raw[3 * j + ch] + raw[3 * (j + 4) + ch];
            const float s = ((a4[0] + a4[1]) + a4[2]) + a4[3];
            c[24 + ch] = 8.0f / (s + 1e-30f);
        }
        break;
    }
    case AISP_OP_USM: {  // 5-tap gaussian and its sigma-derivative   isp/sharpen.py:15-23
        float sigma = raw[0];
        float e[5], sum = 0.f, m2 = 0.f;
        for (int i = 0; i < 5; ++i) {
            float x = (float)(i - 2);
            float t = x / sigma;
            e[i] = expf(-0.5f * (t * t));
            sum += e[i];
        }
        for (int i = 0; i < 5; ++i) { c[i] = e[i] / sum; }
        for (int i = 0; i < 5; ++i) { float x = (float)(i - 2); m2 += c[i] * x * x; }
        float inv3 = 1.0f / (sigma * sigma * sigma);
        for (int i = 0; i < 5; ++i) { float x = (float)(i - 2); c[5 + i] = c[i] * (x * x - m2) * inv3; }
        c[10] = raw[1];  // amount
        break;
    }
    default:  // E handled above; G, W, Ct, S+, BW, Shr, ShrV2, NLM use the raw values
        for (int k = 0; k < 3; ++k) c[k] = raw[k];
        break;
    }
}

// ---------------------------------------------------------------------------------------------
// raw per-sample partial sums -> parameter gradients (chain rule through the derived constants)
//   a : AISP_ACC_STRIDE reduced accumulators (double), c : derived constants, gp : PSTRIDE outputs
// ---------------------------------------------------------------------------------------------
__device__ inline void finalize_grads(int op, const double* a, const float* c, const float* raw, float* gp) {
    for (int k = 0; k < AISP_PSTRIDE; ++k) gp[k] = 0.f;
    switch (op) {
    case AISP_OP_EXPOSURE:  // a0 = sum gy*x ; dy/dp = x * 2^p * ln2
        gp[0] = (float)(a[0] * (double)c[0] * (double)AISP_LN2F);
        break;
    case AISP_OP_GAMMA:  // a0 = sum gy*y*log2(max(x,.001))
        gp[0] = (float)(a[0] * (double)AISP_LN2F);
        break;
    case AISP_OP_WB:
        for (int k = 0; k < 3; ++k) gp[k] = (float)a[k];
        break;
    case AISP_OP_CCM:  // a[3i+j] = sum gy_i*x_j ; through M/rowsum
        for (int i = 0; i < 3; ++i) {
            double dot = a[3 * i] * c[3 * i] + a[3 * i + 1] * c[3 * i + 1] + a[3 * i + 2] * c[3 * i + 2];
            for (int k = 0; k < 3; ++k) gp[3 * i + k] = (float)((a[3 * i + k] - dot) * (double)c[9 + i]);
        }
        break;
    case AISP_OP_TONE:  // a[k] = 8 * sum gy*seg_k ; a[8] = sum gy*y ; dy/dp_k = sc*seg_k - y*sc/8
        for (int k = 0; k < 8; ++k) gp[k] = (float)((double)c[8] * 0.125 * (a[k] - a[8]));
        break;
    case AISP_OP_COLOR:  // a[3k+ch], a[24+ch]
        for (int k = 0; k < 8; ++k)
            for (int ch = 0; ch < 3; ++ch)
                gp[3 * k + ch] = (float)((double)c[24 + ch] * 0.125 * (a[3 * k + ch] - a[24 + ch]));
        break;
    case AISP_OP_USM:  // a0 = sum gy*(d blur/d sigma), a1 = sum gy*(x - blur)
        gp[0] = (float)(-(double)raw[1] * a[0]);
        gp[1] = (float)a[1];
        break;
    default:  // Ct, S+, BW, Shr, ShrV2, NLM: a0 is the gradient
        gp[0] = (float)a[0];
        break;
    }
}

// ---------------------------------------------------------------------------------------------
// block reduction of N per-thread partial sums -> one row of the scratch (N <= AISP_ACC_STRIDE)
// warp shuffles first, then a fixed-order sum over the 8 warps: deterministic.
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void block_reduce_store(float (&acc)[N], float* red /*[kWarps][32]*/, float* dst) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp * AISP_ACC_STRIDE + k] = v;
    }
    __syncthreads();
    if (threadIdx.x < AISP_ACC_STRIDE) {
        float s = 0.f;
        if (threadIdx.x < N) {
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += red[w * AISP_ACC_STRIDE + threadIdx.x];
        }
        dst[threadIdx.x] = s;
    }
}

enum { FAMILY_POINTWISE = 0, FAMILY_SHARPEN = 1, FAMILY_NLM = 2 };

// Gradient that reaches the BLOCK MEANS of an output image (the critic pools the retouched image,
// value.py:63, so train.py:341-342 sends a gradient into the pooled image as well as into the full one).
// d mean / d pixel = 1 / (bh * bw): the backward kernels add g[b, ch, y >> bhs, x >> bws] * inv_area to the
// upstream gradient as they load it -- no up-sampled gradient image is ever materialised.  W, bh, bw are
// powers of two here (512 -> 64 is); g == nullptr: no pooled gradient.
struct PooledGrad {
    const float* g;   // [B,3,oh,ow]
    int ws, bhs, bws; // log2(W), log2(bh), log2(bw)
    int oh, ow;
    float inv_area;
};
inline PooledGrad no_pooled_grad() { return PooledGrad{nullptr, 0, 0, 0, 0, 0, 0.f}; }
__device__ __forceinline__ float pooled_at(const PooledGrad& pg, int b, int ch, int y, int x) {
    return __ldg(pg.g + (((size_t)b * 3 + ch) * pg.oh + (y >> pg.bhs)) * pg.ow + (x >> pg.bws)) * pg.inv_area;
}

// Filter-bank launches (aisp_bank_*): F filters applied to the same batch.  "Virtual sample"
// v = image * F + slot indexes out / params / grad_out / grad_params / scratch; the image (and the
// compact NLM stash) is indexed by v / F.  A family's launch only covers its own `n` slots: the
// sample coordinate of the grid runs over image * n + j and `slots` (4 bits per entry) maps j to
// the slot.  The op code of slot f is nibble f of `opsn` (the bank's op list lives on the host, so
// there is no device-side ops array).  A plain batch is {F = 1, n = 0}: v = grid coordinate.
constexpr int kMaxBankFilters = 16;
struct BankMap {
    int F;
    int n;
    unsigned long long slots;
    unsigned long long opsn;
};
inline BankMap plain_batch() { return BankMap{1, 0, 0ull, 0ull}; }
__device__ __forceinline__ int bank_sample(const BankMap& bm, int g) {
    return bm.n == 0 ? g : (g / bm.n) * bm.F + (int)((bm.slots >> (4 * (g % bm.n))) & 15ull);
}
__device__ __forceinline__ int bank_op(const BankMap& bm, int v) { return (int)((bm.opsn >> (4 * (v % bm.F))) & 15ull); }
// op of sample v (first step of its sequence when S > 1)
__device__ __forceinline__ int sample_op(const int32_t* __restrict__ ops, const BankMap& bm, int v, int S = 1) {
    return ops ? ops[(size_t)v * S] : bank_op(bm, v);
}
__host__ __device__ __forceinline__ bool in_family(int op, int family) {
    return family == FAMILY_POINTWISE ? is_pointwise(op)
         : family == FAMILY_SHARPEN   ? is_sharpen(op)
                                      : op == AISP_OP_NLM;
}

// Second stage shared by every family: grid = B, block = kThreads.  Sums `nrows` scratch rows of
// sample b in fp64 (fixed order), applies finalize_grads and writes grad_params[b, :].
__global__ void finalize_kernel(const float* __restrict__ partial, int nrows, const float* __restrict__ params,
                                const int32_t* __restrict__ ops, int family, float* __restrict__ grad_params,
                                BankMap bm);

// launch geometry shared between kernels and aisp_bwd_scratch_bytes
constexpr int kPwChunkPx = 4096;       // pixels per CTA in the per-pixel kernels
constexpr int kShTileW = 128, kShTileH = 16;   // sharpen / USM tile
constexpr int kNlmTileW = 28, kNlmTileH = 32;  // NLM tile (28 = 32 lanes - 2*2 box halo)

}  // namespace aisp
