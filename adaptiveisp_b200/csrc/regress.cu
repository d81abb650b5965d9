// Feature -> parameter regressors of every filter of a bank in ONE launch, forward and backward
// (`Filter.filter_param_regressor` of isp/filters.py:215-708: tanh_range / exp / sigmoid / the white-balance
// luminance normalisation), writing the packed [B,F,AISP_PSTRIDE] rows the image kernels read.
//
// The arithmetic is a few dozen flops per (sample, filter); what it replaces is launch count: the
// reference (and a per-module PyTorch statement of it) spends 6-10 elementwise launches per filter and as
// many again in backward -- ~200 launches of a few hundred bytes each per agent step, which is what makes
// the eager caller CPU-bound and pads a captured step with ~2 us kernels.  One thread per (sample, filter).
#include "aisp_common.cuh"

namespace aisp {

// The ranges of config.py:28-38 the regressors read, as (lo, span) pairs: the reference forms r - l in
// Python floats (float64) and only then meets the fp32 tensor, so the span is rounded ONCE on the host.
struct RegressCfg {
    float exposure_lo, exposure_span, gamma_lo, gamma_span;
    float tone_lo, tone_span, color_lo, color_span, color_shift;
    float usm_lo, usm_span, sharpen_lo, sharpen_span, ccm_lo, ccm_span;
};

// tanh_range(l, r, initial)(x) = (tanh(x + shift) * 0.5 + 0.5) * (r - l) + l        isp/filters.py:21-34
__device__ __forceinline__ float tanh_range_f(float x, float lo, float span, float shift, float* dydx) {
    const float t = tanhf(x + shift);
    if (dydx) *dydx = 0.5f * span * (1.0f - t * t);
    return (t * 0.5f + 0.5f) * span + lo;
}

// p[0..n) = regress_op(f[0..n)); with gp != nullptr also gf[0..n) = J^T gp
__device__ inline void regress_one(int op, const RegressCfg& c, const float* f, float* p, const float* gp, float* gf) {
    switch (op) {
    case AISP_OP_EXPOSURE: {   // tanh_range(-r, r, initial=0): shift = atanh(0) = 0
        float d;
        p[0] = tanh_range_f(f[0], c.exposure_lo, c.exposure_span, 0.f, &d);
        if (gp) gf[0] = gp[0] * d;
        break;
    }
    case AISP_OP_GAMMA: {      // exp(tanh_range(-ln g, ln g)(f))
        float d;
        const float v = expf(tanh_range_f(f[0], c.gamma_lo, c.gamma_span, 0.f, &d));
        p[0] = v;
        if (gp) gf[0] = gp[0] * v * d;
        break;
    }
    case AISP_OP_WB: {         // red feature zeroed; gains / (1e-5 + lum(gains))    isp/filters.py:253-272
        const float w[3] = {0.27f, 0.67f, 0.06f};
        float g[3], d[3];
        for (int i = 0; i < 3; ++i) g[i] = expf(tanh_range_f(i == 0 ? f[0] * 0.f : f[i], -0.5f, 1.0f, 0.f, &d[i]));
        const float n = 1.0f / (1e-5f + 0.27f * g[0] + 0.67f * g[1] + 0.06f * g[2]);
        for (int i = 0; i < 3; ++i) p[i] = g[i] * n;
        if (gp) {
            const float dot = gp[0] * g[0] + gp[1] * g[1] + gp[2] * g[2];
            for (int i = 0; i < 3; ++i) {
                const float gg = gp[i] * n - dot * n * n * w[i];
                gf[i] = (i == 0) ? 0.f : gg * g[i] * d[i];
            }
        }
        break;
    }
    case AISP_OP_TONE:
        for (int k = 0; k < 8; ++k) {
            float d;
            p[k] = tanh_range_f(f[k], c.tone_lo, c.tone_span, 0.f, &d);
            if (gp) gf[k] = gp[k] * d;
        }
        break;
    case AISP_OP_COLOR:
        for (int k = 0; k < 24; ++k) {
            float d;
            p[k] = tanh_range_f(f[k], c.color_lo, c.color_span, c.color_shift, &d);
            if (gp) gf[k] = gp[k] * d;
        }
        break;
    case AISP_OP_CONTRAST: {
        const float t = tanhf(f[0]);
        p[0] = t;
        if (gp) gf[0] = gp[0] * (1.0f - t * t);
        break;
    }
    case AISP_OP_WNB:
    case AISP_OP_SATPLUS:
    case AISP_OP_NLM: {        // torch.sigmoid
        const float s = 1.0f / (1.0f + expf(-f[0]));
        p[0] = s;
        if (gp) gf[0] = gp[0] * s * (1.0f - s);
        break;
    }
    case AISP_OP_USM:
        for (int k = 0; k < 2; ++k) {
            float d;
            p[k] = tanh_range_f(f[k], c.usm_lo, c.usm_span, 0.f, &d);
            if (gp) gf[k] = gp[k] * d;
        }
        break;
    case AISP_OP_SHARPEN:
    case AISP_OP_SHARPEN_V2: {
        float d;
        p[0] = tanh_range_f(f[0], c.sharpen_lo, c.sharpen_span, 0.f, &d);
        if (gp) gf[0] = gp[0] * d;
        break;
    }
    case AISP_OP_CCM:
        for (int k = 0; k < 9; ++k) {
            float d;
            p[k] = tanh_range_f(f[k], c.ccm_lo, c.ccm_span, 0.f, &d);
            if (gp) gf[k] = gp[k] * d;
        }
        break;
    default: break;
    }
}

__device__ __forceinline__ int nparams_of(int op) {
    const int n[AISP_OP_COUNT] = {1, 1, 9, 1, 1, 8, 1, 1, 1, 3, 2, 24, 1};
    return (op >= 0 && op < AISP_OP_COUNT) ? n[op] : 0;
}

// raw [B,Ntot] (filter f's features at columns offs[f] ..), packed [B,F,PSTRIDE]; BWD: gpacked -> graw
template <bool BWD>
__global__ void __launch_bounds__(128)
regress_kernel(const float* __restrict__ raw, const int32_t* __restrict__ fops, const int32_t* __restrict__ offs, int B,
               int F, int Ntot, RegressCfg cfg, float* __restrict__ packed, const float* __restrict__ gpacked,
               float* __restrict__ graw) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * F) return;
    const int b = t / F, f = t - b * F;
    const int op = fops[f], n = nparams_of(op), off = offs[f];
    float in[AISP_PSTRIDE], p[AISP_PSTRIDE], gp[AISP_PSTRIDE], gf[AISP_PSTRIDE];
    for (int k = 0; k < AISP_PSTRIDE; ++k) {
        in[k] = (k < n) ? raw[(size_t)b * Ntot + off + k] : 0.f;
        p[k] = 0.f;
        gf[k] = 0.f;
        gp[k] = (BWD && k < n) ? gpacked[((size_t)b * F + f) * AISP_PSTRIDE + k] : 0.f;
    }
    regress_one(op, cfg, in, p, BWD ? gp : nullptr, gf);
    if (BWD) {
        for (int k = 0; k < n; ++k) graw[(size_t)b * Ntot + off + k] = gf[k];
    } else {
        for (int k = 0; k < AISP_PSTRIDE; ++k) packed[((size_t)b * F + f) * AISP_PSTRIDE + k] = p[k];
    }
}

static RegressCfg make_cfg(const float* v) {
    RegressCfg c;
    c.exposure_lo = v[0]; c.exposure_span = v[1]; c.gamma_lo = v[2]; c.gamma_span = v[3];
    c.tone_lo = v[4]; c.tone_span = v[5]; c.color_lo = v[6]; c.color_span = v[7]; c.color_shift = v[8];
    c.usm_lo = v[9]; c.usm_span = v[10]; c.sharpen_lo = v[11]; c.sharpen_span = v[12]; c.ccm_lo = v[13]; c.ccm_span = v[14];
    return c;
}

cudaError_t launch_regress(const float* raw, const int32_t* fops, const int32_t* offs, int B, int F, int Ntot,
                           const float* cfg15, float* packed, const float* gpacked, float* graw, cudaStream_t st) {
    const RegressCfg c = make_cfg(cfg15);
    const int threads = 128, blocks = (B * F + threads - 1) / threads;
    if (gpacked)
        regress_kernel<true><<<blocks, threads, 0, st>>>(raw, fops, offs, B, F, Ntot, c, nullptr, gpacked, graw);
    else
        regress_kernel<false><<<blocks, threads, 0, st>>>(raw, fops, offs, B, F, Ntot, c, packed, nullptr, nullptr);
    return cudaGetLastError();
}

}  // namespace aisp
