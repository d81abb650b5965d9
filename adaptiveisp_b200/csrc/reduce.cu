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

}  // namespace aisp
