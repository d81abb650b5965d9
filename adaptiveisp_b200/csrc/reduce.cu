// Block-mean down-sampling of an image batch in one streaming pass.
//
// The consumers of a retouched batch in the reference each re-read the full-resolution image:
// AdaptiveAvgPool2d(64,64) in Agent.forward (agent.py:97) and in Value.forward (value.py:63), the
// per-image mean of the truncation / pool-refill tests (train.py:288-290,374) and the NaN/Inf guard
// (train.py:374).  All of them are functions of the 64x64 block-mean image (the mean of equal-size
// block means is the image mean; a block mean is non-finite iff the block holds a non-finite
// value), so ONE pass producing that image replaces five or six passes over HBM.
//
// Thread <-> one output element; a warp covers 32 adjacent output columns, i.e. 32*bw contiguous
// input floats per row (128-bit loads when bw % 4 == 0).  Read-only, 12 B/px.
#include "pointwise_math.cuh"   // find_stencil (sequence classification)

namespace aisp {

// `ops` != nullptr restricts the pass to the samples of some kernel families: `families` is a bit set over
// (1 << FAMILY_*) of the sample's sequence (the family of its stencil step, FAMILY_POINTWISE without one) --
// used to complete the block means that the per-pixel / sharpen kernels emitted from their store path
// with the samples they could not serve (NLM: its 28-column tiles do not align with pooling blocks).
__global__ void __launch_bounds__(kThreads)
block_mean_kernel(const float* __restrict__ img, float* __restrict__ down, int H, int W, int oh, int ow, int bh,
                  int bw, int vec, const int32_t* __restrict__ ops, const int32_t* __restrict__ seq_len, int S,
                  int families) {
    const int plane = blockIdx.z;                                   // b * 3 + c
    if (ops) {
        const int b = plane / 3;
        int len = seq_len ? min(max(seq_len[b], 0), S) : S;
        const int pos = find_stencil(ops + (size_t)b * S, len, &len);
        const int fam = pos < 0 ? FAMILY_POINTWISE : (ops[(size_t)b * S + pos] == AISP_OP_NLM ? FAMILY_NLM : FAMILY_SHARPEN);
        if (!((families >> fam) & 1)) return;
    }
    const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
    const int oy = blockIdx.y * kWarps + (threadIdx.x >> 5);
    if (ox >= ow || oy >= oh) return;
    const float* src = img + (size_t)plane * H * W + (size_t)oy * bh * W + (size_t)ox * bw;
    float acc = 0.f;
    if (vec && bw == 8) {  // the 512 -> 64 case: two 128-bit loads per row, four rows in flight
        int r = 0;
        for (; r + 4 <= bh; r += 4) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[2 * k] = ldg_stream4(src + (size_t)(r + k) * W);
                v[2 * k + 1] = ldg_stream4(src + (size_t)(r + k) * W + 4);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                acc += ((v[2 * k].x + v[2 * k].y) + (v[2 * k].z + v[2 * k].w)) +
                       ((v[2 * k + 1].x + v[2 * k + 1].y) + (v[2 * k + 1].z + v[2 * k + 1].w));
        }
        for (; r < bh; ++r) {
            const float4 a = ldg_stream4(src + (size_t)r * W), b = ldg_stream4(src + (size_t)r * W + 4);
            acc += ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
        }
    } else if (vec) {
        for (int r = 0; r < bh; ++r) {
            const float* row = src + (size_t)r * W;
            float s = 0.f;
            for (int c = 0; c < bw; c += 4) {
                const float4 v = ldg_stream4(row + c);
                s += (v.x + v.y) + (v.z + v.w);
            }
            acc += s;
        }
    } else {
        for (int r = 0; r < bh; ++r) {
            const float* row = src + (size_t)r * W;
            float s = 0.f;
            for (int c = 0; c < bw; ++c) s += __ldg(row + c);
            acc += s;
        }
    }
    down[((size_t)plane * oh + oy) * ow + ox] = acc / (float)(bh * bw);
}

cudaError_t launch_block_mean_masked(const float* img, float* down, int B, int H, int W, int oh, int ow,
                                     const int32_t* ops, const int32_t* seq_len, int S, int families, cudaStream_t st) {
    const int bh = H / oh, bw = W / ow;
    const int vec = ((bw & 3) == 0) && ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(img) & 15u) == 0);
    dim3 grid((ow + 31) / 32, (oh + kWarps - 1) / kWarps, B * 3);
    block_mean_kernel<<<grid, kThreads, 0, st>>>(img, down, H, W, oh, ow, bh, bw, vec, ops, seq_len, S, families);
    return cudaGetLastError();
}

cudaError_t launch_block_mean(const float* img, float* down, int B, int H, int W, int oh, int ow, cudaStream_t st) {
    return launch_block_mean_masked(img, down, B, H, W, oh, ow, nullptr, nullptr, 1, 7, st);
}

// ---------------------------------------------------------------------------------------------
// The critic's image statistics (value.py:64-75) of the pooled image [B,3,h,w] in one launch, one CTA per
// image: mean and unbiased variance of the luminance 0.27 r + 0.67 g + 0.06 b + 1e-5, and the mean of
// (max - min) / (min(max + min, 2 - max - min) + 0.01) over the clipped channels.  The reference spends
// ~20 ATen launches on these 12k numbers per image.  Sums are fp32 per thread, fp64 across the CTA (fixed
// order); the variance is taken around the mean in a second sweep, like torch.var.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
value_stats_kernel(const float* __restrict__ down, int n, float* __restrict__ stats) {
    __shared__ double red[2][kWarps];
    __shared__ double s_mean;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* r = down + (size_t)b * 3 * n;
    const float* g = r + n;
    const float* bl = g + n;
    float sl = 0.f, ss = 0.f;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const float x = r[i], y = g[i], z = bl[i];
        sl += ((x * 0.27f + y * 0.67f) + z * 0.06f) + 1e-5f;
        const float cx = clip01(x), cy = clip01(y), cz = clip01(z);
        const float mx = max_nan(cx, max_nan(cy, cz)), mn = min_nan(cx, min_nan(cy, cz));
        ss += (mx - mn) / (min_nan(mx + mn, (2.0f - mx) - mn) + 1e-2f);
    }
    double dl = sl, ds = ss;
    for (int o = 16; o > 0; o >>= 1) { dl += __shfl_xor_sync(0xffffffffu, dl, o); ds += __shfl_xor_sync(0xffffffffu, ds, o); }
    if (lane == 0) { red[0][warp] = dl; red[1][warp] = ds; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tl = 0.0, ts = 0.0;
        for (int w = 0; w < kWarps; ++w) { tl += red[0][w]; ts += red[1][w]; }
        s_mean = tl / n;
        stats[b * 3 + 0] = (float)(tl / n);
        stats[b * 3 + 2] = (float)(ts / n);
    }
    __syncthreads();
    const float mean = (float)s_mean;
    float sv = 0.f;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const float d = (((r[i] * 0.27f + g[i] * 0.67f) + bl[i] * 0.06f) + 1e-5f) - mean;
        sv = fmaf(d, d, sv);
    }
    double dv = sv;
    for (int o = 16; o > 0; o >>= 1) dv += __shfl_xor_sync(0xffffffffu, dv, o);
    __syncthreads();
    if (lane == 0) red[0][warp] = dv;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tv = 0.0;
        for (int w = 0; w < kWarps; ++w) tv += red[0][w];
        stats[b * 3 + 1] = (float)(tv / (n > 1 ? n - 1 : 1));
    }
}

cudaError_t launch_value_stats(const float* down, int B, int n, float* stats, cudaStream_t st) {
    value_stats_kernel<<<B, kThreads, 0, st>>>(down, n, stats);
    return cudaGetLastError();
}

}  // namespace aisp
