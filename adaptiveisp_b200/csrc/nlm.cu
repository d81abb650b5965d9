// Non-local-means denoise (gray distance, 11x11 search, 5x5 patch, circular boundaries) for sm_100a.
// Reference: DenoiseFilter.process isp/filters.py:582-586 -> NonLocalMeansGray isp/denoise.py:93-119.
//
// This stage is NOT HBM-bound: per output pixel it needs 121 patch distances (each a 5x5 box sum of
// squared luma differences), 121 sqrt and 121 exp.  The kernel is organised so that the FP32 pipe,
// the MUFU pipe and the LSU/shuffle path are all loaded about equally:
//
//   CTA <-> (sample, 28 x 32 tile); warp w owns 4 rows, lane l owns one column.
//   The tile (+7 halo, wrapped circularly) of clipped RGB and of luma Y is staged in shared memory.
//   For a fixed horizontal offset dx a thread pulls one column of Y (18 values) and one column of
//   RGB (14 x 3 values) into registers and reuses them for all 11 vertical offsets dy.
//   Squared differences are summed 5-high in registers (shared partial sums across the 4 rows) and
//   5-wide across lanes with a 3-shuffle tree -- lane l ends up with the box sum centred on its
//   output column x0 + l (lanes 0..27 produce outputs; lanes 28..31 only feed the tree).
//
// When `dout_dh` is requested the same pass also accumulates sum(w*d) and sum(w*d*x), which give
// d out / d h in closed form, so training never runs a second 121-shift pass.
#include <cstdlib>
#include "pointwise_math.cuh"   // fwd_px / stage_consts for the fused per-pixel prologue and epilogue (-fmad=false)

namespace aisp {

constexpr int kNlmHalo = 7;                       // search radius 5 + patch radius 2
constexpr int kNlmRows = 4;                       // output rows per thread
// Lane layout: LW lanes form one row group (LW = 32: a warp is one group of 28 output columns;
// LW = 16: a warp is two groups of 12 output columns, stacked vertically).  The 16-lane layout serves
// the remainder column of an image whose width leaves <= 12 columns after the 28-wide tiles (8 for
// W = 512): those columns then cost half a tile column instead of a whole one.
template <int LW>
struct NlmGeo {
    static constexpr int NG = 32 / LW;                       // row groups per warp
    static constexpr int TW = LW - 4;                        // output columns per group
    static constexpr int TH = kNlmRows * kWarps * NG;        // output rows per CTA
    static constexpr int SW = LW + 10 + (LW == 16 ? 2 : 0);  // staged columns x0-7 .. x0+LW+2 (+2 pad: the two row groups
                                                             // of a warp then sit 16 banks apart)
    static constexpr int SH = TH + 2 * kNlmHalo;             // staged rows    y0-7 .. y0+TH+6
};
static_assert(NlmGeo<32>::TW == kNlmTileW && NlmGeo<32>::TH == kNlmTileH, "geometry shared with the header");

__device__ __forceinline__ int wrap(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// variant: 0 = DenoiseFilter.process (isp/filters.py:582-586: input clipped, gray distance);
//   1 = the bare NonLocalMeansGray module (isp/denoise.py:93-119): distances on the luma of the CLIPPED
//       image (rgb_to_luminance clips, :14), averages of the UNclipped one;
//   2 + c = channel c of the bare NonLocalMeans module (isp/denoise.py:68-90): per-channel distances and
//       weights on the unclipped image -- the host runs c = 0, 1, 2 (each pass writes its own channel).
//   Variants >= 1 are modules, not Filters: no lerp term.
// SEQ: the sample's op SEQUENCE [per-pixel prologue] -> NLM -> [per-pixel epilogue] in one launch (params
// [B,S,PSTRIDE], ops [B,S]): the prologue runs on every pixel as it is staged (tile + halo, wrapped
// circularly: per-pixel filters commute with the wrap), the epilogue on the finished outputs in
// registers.  SEQ == false is the plain single-filter kernel (ops [B], params [B,PSTRIDE]).
template <bool WITH_GRAD, int LW, bool SEQ>
__global__ void __launch_bounds__(kThreads, WITH_GRAD ? 3 : 4)
nlm_kernel(const float* __restrict__ img, float* __restrict__ out, float* __restrict__ dout_dh,
           float* __restrict__ wsum_out, const float* __restrict__ params, const int32_t* __restrict__ ops, int H,
           int W, int x_off, BankMap bm, const int32_t* __restrict__ seq_len, int S, int clip_each, int variant) {
    pdl_prologue();
    using Geo = NlmGeo<LW>;
    constexpr int kNlmSmH = Geo::SH, kNlmSmW = Geo::SW;
    __shared__ float sY[kNlmSmH][kNlmSmW];
    __shared__ float sC[3][kNlmSmH][kNlmSmW];
    __shared__ float sraw[SEQ ? AISP_MAX_STEPS : 1][kConst];
    __shared__ float ssc[SEQ ? AISP_MAX_STEPS : 1][kConst];
    __shared__ int ssop[SEQ ? AISP_MAX_STEPS : 1];
    __shared__ int rownz[kNlmSmH];               // staged row holds a non-zero value (see the zero shortcut below)
    const int b = bank_sample(bm, blockIdx.z);   // filter-bank launches: see BankMap
    int pos = 0, len = 1;
    if (SEQ) {
        len = seq_len ? min(max(seq_len[b], 0), S) : S;
        pos = find_stencil(ops + (size_t)b * S, len, &len);
        if (pos < 0 || ops[(size_t)b * S + pos] != AISP_OP_NLM) return;   // another family owns this sample
        stage_consts(params, ops, b, S, len, sraw, ssc, ssop, bm);
    } else {
        if (sample_op(ops, bm, b) != AISP_OP_NLM) return;
    }
    const int x0 = x_off + blockIdx.x * Geo::TW, y0 = blockIdx.y * Geo::TH;
    const size_t plane = (size_t)H * W;
    const float* src = img + (size_t)(b / bm.F) * 3 * plane;
    const size_t sb = (size_t)(b / bm.F);        // stashes stay compact: one NLM slot per image

    // stage clipped RGB and luma of the wrapped tile + halo       (isp/filters.py:583, denoise.py:11-17)
    // The staged region is walked as ONE flat index range, fully unrolled: a thread's (up to) eight
    // elements x three planes are independent loads that are all in flight together -- one or two L2 round
    // trips per CTA instead of one per (row, column-pass) of a warp-per-row walk.  Row wrap / column wrap
    // by conditional adds when the image is larger than the tile footprint; the generic modulo only serves
    // tiny images.
    const int lane32 = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool wide = (W >= kNlmSmW) && (H >= kNlmSmH);
    constexpr int kStageN = kNlmSmH * kNlmSmW;
    constexpr int kStageIt = (kStageN + kThreads - 1) / kThreads;
    if (threadIdx.x < kNlmSmH) rownz[threadIdx.x] = 0;
    __syncthreads();
    {
        float vr[kStageIt], vg[kStageIt], vb[kStageIt];
#pragma unroll
        for (int it = 0; it < kStageIt; ++it) {
            const int e = it * kThreads + threadIdx.x;
            vr[it] = vg[it] = vb[it] = 0.f;
            if (e < kStageN) {
                const int row = e / kNlmSmW, col = e - row * kNlmSmW;
                int gy = y0 - kNlmHalo + row, gx = x0 - kNlmHalo + col;
                gy = wide ? (gy < 0 ? gy + H : (gy >= H ? gy - H : gy)) : wrap(gy, H);
                gx = wide ? (gx < 0 ? gx + W : (gx >= W ? gx - W : gx)) : wrap(gx, W);
                const float* rp = src + (size_t)gy * W + gx;
                vr[it] = __ldg(rp);
                vg[it] = __ldg(rp + plane);
                vb[it] = __ldg(rp + 2 * plane);
            }
        }
#pragma unroll
        for (int it = 0; it < kStageIt; ++it) {
            const int e = it * kThreads + threadIdx.x;
            const int row = e / kNlmSmW, col = e - row * kNlmSmW;
            float r = vr[it], g = vg[it], bl = vb[it];
            if (SEQ) {
                for (int k = 0; k < pos; ++k) {
                    fwd_px<true>(ssop[k], ssc[k], r, g, bl);
                    if (clip_each) { r = clip01(r); g = clip01(g); bl = clip01(bl); }
                }
            }
            const float cr = clip01(r), cg = clip01(g), cb = clip01(bl);
            float yy = (0.299f * cr + 0.587f * cg) + 0.114f * cb;
            if (variant == 0) { r = cr; g = cg; bl = cb; }                       // the filter averages what it clipped
            else if (variant >= 2) yy = (variant == 2) ? r : (variant == 3 ? g : bl);
            const bool in = e < kStageN;
            if (in) {
                sC[0][row][col] = r;
                sC[1][row][col] = g;
                sC[2][row][col] = bl;
                sY[row][col] = yy;
            }
            // "this staged row holds a non-zero value" (NaN counts): the lanes of a warp that staged the same
            // row vote, the lowest of them ORs the flag in (a warp's 32 consecutive elements span 2-3 rows)
            const bool nz = in && ((r != 0.f) | (g != 0.f) | (bl != 0.f));
            const unsigned any = __ballot_sync(0xffffffffu, nz);
            const unsigned grp = __match_any_sync(0xffffffffu, in ? row : -1);
            if (in && (any & grp) && lane32 == __ffs(grp) - 1) atomicOr(&rownz[row], 1);
        }
    }
    __syncthreads();

    const int lane = lane32 % LW;                               // column within the row group
    const int r0 = (warp * Geo::NG + lane32 / LW) * kNlmRows;   // first output row of this thread, relative to y0
    const float h = params[((size_t)b * (SEQ ? S : 1) + pos) * AISP_PSTRIDE];
    const float hh = fmaxf(h, 0.f) + 1e-8f;              // relu(h) + EPS   (denoise.py:112)
    const float negk = -1.4426950408889634f / hh;        // exp(-d/hh) = 2^(d * negk)

    // own luma column: rows r0-2 .. r0+5 at column x0-2+lane
    float yo[kNlmRows + 4];
#pragma unroll
    for (int j = 0; j < kNlmRows + 4; ++j) yo[j] = sY[r0 + j + 5][lane + 5];

    float ac[3][kNlmRows], bc[3][kNlmRows];
    f32x2 wsum2[kNlmRows / 2], wd2[kNlmRows / 2];
    const f32x2 negk2 = pack2(negk, negk);
#pragma unroll
    for (int i = 0; i < kNlmRows; ++i) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { ac[c][i] = 0.f; bc[c][i] = 0.f; }
    }
#pragma unroll
    for (int i = 0; i < kNlmRows / 2; ++i) { wsum2[i] = pack2(0.f, 0.f); wd2[i] = pack2(0.f, 0.f); }

    // Zero shortcut.  LOD frames are letterboxed with EXACT-zero bars (dataset.py:834-838,887-888: a 3:2
    // frame leaves a third of the 512 rows black), and wherever the whole 15 x 15 footprint of a warp's
    // rows is zero every patch distance is 0, every weight exp(0) = 1 and the result is sum(0) / 121 = 0
    // -- bit for bit what the 121-shift loop would produce -- so the warp skips the loop.  The test is
    // warp-uniform (the loop's shuffles need every lane) and costs one flag per staged row.
    bool live = false;
    for (int t = 0; t < kNlmRows + 2 * kNlmHalo; ++t) live |= (rownz[r0 + t] != 0);
    live = __any_sync(0xffffffffu, live);

    // source offsets run +5 .. -5 so that terms are accumulated in the reference's order
    // (x_shift outer, y_shift inner, shifted(p) = x(p - shift); denoise.py:106-109)
    for (int dx = live ? 5 : -6; dx >= -5; --dx) {
        float ys[kNlmRows + 14];  // luma rows r0-7 .. r0+10 at column x0-2+lane+dx
#pragma unroll
        for (int t = 0; t < kNlmRows + 14; ++t) ys[t] = sY[r0 + t][lane + dx + 5];
        const int ccol = min(lane + dx + kNlmHalo, kNlmSmW - 1);  // output column x0+lane, shifted
        float cs[3][kNlmRows + 10];  // RGB rows r0-5 .. r0+8
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int t = 0; t < kNlmRows + 10; ++t) cs[c][t] = sC[c][r0 + t + 2][ccol];

#pragma unroll
        for (int dy = 5; dy >= -5; --dy) {
            // 8 squared differences feed 4 five-high sums that share partial sums (12 FMA-pipe ops: one
            // FMUL, eleven FFMA that each square-and-add).  The file is compiled with -fmad=false (the fused
            // per-pixel prologue must round like ATen), so the multiply-adds are spelled out.
            float t[kNlmRows + 4];
#pragma unroll
            for (int j = 0; j < kNlmRows + 4; ++j) t[j] = yo[j] - ys[j + dy + 5];
            const float m34 = fmaf(t[3], t[3], t[4] * t[4]);
            const float c234 = fmaf(t[2], t[2], m34);
            const float c2345 = fmaf(t[5], t[5], c234);
            const float c345 = fmaf(t[5], t[5], m34);
            float v[kNlmRows];
            v[0] = fmaf(t[0], t[0], fmaf(t[1], t[1], c234));
            v[1] = fmaf(t[1], t[1], c2345);
            v[2] = fmaf(t[6], t[6], c2345);
            v[3] = fmaf(t[7], t[7], fmaf(t[6], t[6], c345));
            // Rows are handled in pairs so that the adds / multiplies that do not touch the (row-
            // misaligned) RGB column run as Blackwell packed-fp32 instructions (FADD2 / FMUL2: two
            // lanes of work per issue slot -- the kernel is issue-bound, not FMA-pipe-bound).
#pragma unroll
            for (int i = 0; i < kNlmRows; i += 2) {
                // 5-wide sum across lanes l .. l+4  (box centred on output column x0 + lane)
                const f32x2 vv = pack2(v[i], v[i + 1]);
                const f32x2 p2 = add2(vv, pack2(__shfl_down_sync(0xffffffffu, v[i], 1, LW),
                                                __shfl_down_sync(0xffffffffu, v[i + 1], 1, LW)));
                const f32x2 p4 = add2(p2, pack2(__shfl_down_sync(0xffffffffu, lo2(p2), 2, LW),
                                                __shfl_down_sync(0xffffffffu, hi2(p2), 2, LW)));
                const f32x2 box = add2(p4, pack2(__shfl_down_sync(0xffffffffu, v[i], 4, LW),
                                                 __shfl_down_sync(0xffffffffu, v[i + 1], 4, LW)));
                // box >= 0: the reference's relu is a no-op
                const f32x2 dist = pack2(sqrt_approx(lo2(box)), sqrt_approx(hi2(box)));
                const f32x2 arg = mul2(dist, negk2);
                const float w0 = ex2_approx(lo2(arg)), w1 = ex2_approx(hi2(arg));
                const f32x2 ww = pack2(w0, w1);
                wsum2[i >> 1] = add2(wsum2[i >> 1], ww);
                const int t = i + dy + 5;
                ac[0][i] = fmaf(w0, cs[0][t], ac[0][i]);
                ac[1][i] = fmaf(w0, cs[1][t], ac[1][i]);
                ac[2][i] = fmaf(w0, cs[2][t], ac[2][i]);
                ac[0][i + 1] = fmaf(w1, cs[0][t + 1], ac[0][i + 1]);
                ac[1][i + 1] = fmaf(w1, cs[1][t + 1], ac[1][i + 1]);
                ac[2][i + 1] = fmaf(w1, cs[2][t + 1], ac[2][i + 1]);
                if (WITH_GRAD) {
                    const f32x2 wdi = mul2(ww, dist);
                    wd2[i >> 1] = add2(wd2[i >> 1], wdi);
                    const float e0 = lo2(wdi), e1 = hi2(wdi);
                    bc[0][i] = fmaf(e0, cs[0][t], bc[0][i]);
                    bc[1][i] = fmaf(e0, cs[1][t], bc[1][i]);
                    bc[2][i] = fmaf(e0, cs[2][t], bc[2][i]);
                    bc[0][i + 1] = fmaf(e1, cs[0][t + 1], bc[0][i + 1]);
                    bc[1][i + 1] = fmaf(e1, cs[1][t + 1], bc[1][i + 1]);
                    bc[2][i + 1] = fmaf(e1, cs[2][t + 1], bc[2][i + 1]);
                }
            }
        }
    }

    const int gx = x0 + lane;
    if (lane >= Geo::TW || gx >= W) return;
    const float inv_h2 = (h > 0.f) ? 1.0f / (hh * hh) : 0.f;  // relu'(h)
    float wsum[kNlmRows], wd[kNlmRows];
#pragma unroll
    for (int i = 0; i < kNlmRows / 2; ++i) {
        wsum[2 * i] = lo2(wsum2[i]); wsum[2 * i + 1] = hi2(wsum2[i]);
        wd[2 * i] = lo2(wd2[i]); wd[2 * i + 1] = hi2(wd2[i]);
    }
    if (!live) {   // all 121 weights are exactly 1
#pragma unroll
        for (int i = 0; i < kNlmRows; ++i) wsum[i] = 121.0f;
    }
#pragma unroll
    for (int i = 0; i < kNlmRows; ++i) {
        const int gy = y0 + r0 + i;
        if (gy >= H) continue;
        const float iw = 1.0f / wsum[i];
        if (wsum_out) wsum_out[sb * plane + (size_t)gy * W + gx] = wsum[i];
        float yv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // 0 * x: the (1 - mask) * img term of the reference's lerp (isp/filters.py:115), NaN iff the
            // filter's input pixel is inf / NaN.  Plain kernel: the UNclipped centre pixel (an L2 hit, the
            // tile was just staged); after a fused prologue the staged (clipped) value stands in for it.
            float y = ac[c][i] * iw;
            if (variant == 0) {
                const float xin = (SEQ && pos > 0) ? sC[c][r0 + i + kNlmHalo][lane + kNlmHalo]
                                                   : __ldg(src + (size_t)c * plane + (size_t)gy * W + gx);
                y = fmaf(0.f, xin, y);
            }
            yv[c] = clip01(y);
            if (WITH_GRAD && (variant < 2 || c == variant - 2))
                dout_dh[sb * 3 * plane + (size_t)c * plane + (size_t)gy * W + gx] =
                    pass01(y) * (bc[c][i] - y * wd[i]) * iw * inv_h2;
        }
        if (SEQ) {
            for (int k = pos + 1; k < len; ++k) {
                fwd_px<true>(ssop[k], ssc[k], yv[0], yv[1], yv[2]);
                if (clip_each) { yv[0] = clip01(yv[0]); yv[1] = clip01(yv[1]); yv[2] = clip01(yv[2]); }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (variant < 2 || c == variant - 2) out[(size_t)b * 3 * plane + (size_t)c * plane + (size_t)gy * W + gx] = yv[c];
    }
}

// ---------------------------------------------------------------------------------------------
// Two-columns-per-lane layout (round 2): lane l owns the ADJACENT staged columns 2l, 2l+1 of a 64-column
// row group and two rows, so every per-pixel quantity of a thread is a natural {even column, odd column}
// pair and the whole arithmetic runs as Blackwell packed fp32 (sub/mul/fma/add.f32x2: one issue slot per
// two pixels) -- the kernel is bound by issue slots, not by the FMA pipe.  The 5-wide box sum needs three
// shuffles per PAIR of pixels (pair sums P from lanes l+1, l+2 and the even column of lane l+2) instead of
// three per pixel, lanes 0..29 of 32 produce outputs (60 of 64 columns), and the luma / RGB columns are
// not cached per horizontal offset but slide through registers along dy (one new row per step).
//   CTA <-> (sample, 60 x 16 tile); warp w owns rows 2w, 2w+1.
//   Shared memory holds the staged tile twice, the second copy shifted by one float, so that the pair
//   starting at ANY column is an aligned 64-bit load (even horizontal offsets read copy A, odd ones B).
// ---------------------------------------------------------------------------------------------
constexpr int kN2Rows = 2;                               // output rows per thread
constexpr int kN2TW = 60, kN2TH = kN2Rows * kWarps;      // outputs per CTA
constexpr int kN2SW = kN2TW + 2 * kNlmHalo;              // 74 staged columns  x0-7 .. x0+66
constexpr int kN2SH = kN2TH + 2 * kNlmHalo;              // 30 staged rows     y0-7 .. y0+22
constexpr int kN2RS = 76;                                // row stride in floats (even; copy A stores column s at s+1)
constexpr int kN2Plane = kN2SH * kN2RS;
constexpr int kN2Copy = 4 * kN2Plane;                    // planes: luma, R, G, B
constexpr int kN2SmemBytes = 2 * kN2Copy * (int)sizeof(float);

template <bool WITH_GRAD, bool SEQ>
__global__ void __launch_bounds__(kThreads, 3)
nlm2_kernel(const float* __restrict__ img, float* __restrict__ out, float* __restrict__ dout_dh,
            float* __restrict__ wsum_out, const float* __restrict__ params, const int32_t* __restrict__ ops, int H,
            int W, BankMap bm, const int32_t* __restrict__ seq_len, int S, int clip_each, int variant, int vec8) {
    pdl_prologue();
    extern __shared__ __align__(16) float n2sm[];
    __shared__ float sraw[SEQ ? AISP_MAX_STEPS : 1][kConst];
    __shared__ float ssc[SEQ ? AISP_MAX_STEPS : 1][kConst];
    __shared__ int ssop[SEQ ? AISP_MAX_STEPS : 1];
    __shared__ int rownz[kN2SH];                 // staged row holds a non-zero value (zero shortcut: see nlm_kernel)
    const int b = bank_sample(bm, blockIdx.z);
    int pos = 0, len = 1;
    if (SEQ) {
        len = seq_len ? min(max(seq_len[b], 0), S) : S;
        pos = find_stencil(ops + (size_t)b * S, len, &len);
        if (pos < 0 || ops[(size_t)b * S + pos] != AISP_OP_NLM) return;   // another family owns this sample
        stage_consts(params, ops, b, S, len, sraw, ssc, ssop, bm);
    } else {
        if (sample_op(ops, bm, b) != AISP_OP_NLM) return;
    }
    const int x0 = blockIdx.x * kN2TW, y0 = blockIdx.y * kN2TH;
    const size_t plane = (size_t)H * W;
    const float* src = img + (size_t)(b / bm.F) * 3 * plane;
    const size_t sb = (size_t)(b / bm.F);        // stashes stay compact: one NLM slot per image
    const float h = params[((size_t)b * (SEQ ? S : 1) + pos) * AISP_PSTRIDE];
    const float hh = fmaxf(h, 0.f) + 1e-8f;              // relu(h) + EPS   (denoise.py:112)
    const float negk = -1.4426950408889634f / hh;        // exp(-d/hh) = 2^(d * negk)

    // stage the wrapped tile + halo (flat, fully unrolled: see nlm_kernel), both copies
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool wide = (W >= kN2SW) && (H >= kN2SH);
    constexpr int kStageN = kN2SH * kN2SW;
    constexpr int kStageIt = (kStageN + kThreads - 1) / kThreads;
    if (threadIdx.x < kN2SH) rownz[threadIdx.x] = 0;
    __syncthreads();
    {
        float vr[kStageIt], vg[kStageIt], vb[kStageIt];
#pragma unroll
        for (int it = 0; it < kStageIt; ++it) {
            const int e = it * kThreads + threadIdx.x;
            vr[it] = vg[it] = vb[it] = 0.f;
            if (e < kStageN) {
                const int row = e / kN2SW, col = e - row * kN2SW;
                int gy = y0 - kNlmHalo + row, gx = x0 - kNlmHalo + col;
                gy = wide ? (gy < 0 ? gy + H : (gy >= H ? gy - H : gy)) : wrap(gy, H);
                gx = wide ? (gx < 0 ? gx + W : (gx >= W ? gx - W : gx)) : wrap(gx, W);
                const float* rp = src + (size_t)gy * W + gx;
                vr[it] = __ldg(rp);
                vg[it] = __ldg(rp + plane);
                vb[it] = __ldg(rp + 2 * plane);
            }
        }
#pragma unroll
        for (int it = 0; it < kStageIt; ++it) {
            const int e = it * kThreads + threadIdx.x;
            const int row = e / kN2SW, col = e - row * kN2SW;
            float r = vr[it], g = vg[it], bl = vb[it];
            if (SEQ) {
                for (int k = 0; k < pos; ++k) {
                    fwd_px<true>(ssop[k], ssc[k], r, g, bl);
                    if (clip_each) { r = clip01(r); g = clip01(g); bl = clip01(bl); }
                }
            }
            const float cr = clip01(r), cg = clip01(g), cb = clip01(bl);
            float yy = (0.299f * cr + 0.587f * cg) + 0.114f * cb;
            if (variant == 0) { r = cr; g = cg; bl = cb; }                       // the filter averages what it clipped
            else if (variant >= 2) yy = (variant == 2) ? r : (variant == 3 ? g : bl);
            const bool in = e < kStageN;
            if (in) {
                float* pa = n2sm + row * kN2RS + col + 1;
                float* pb = n2sm + kN2Copy + row * kN2RS + col;
                pa[0] = yy;            pb[0] = yy;
                pa[kN2Plane] = r;      pb[kN2Plane] = r;
                pa[2 * kN2Plane] = g;  pb[2 * kN2Plane] = g;
                pa[3 * kN2Plane] = bl; pb[3 * kN2Plane] = bl;
            }
            const bool nz = in && ((r != 0.f) | (g != 0.f) | (bl != 0.f));
            const unsigned any = __ballot_sync(0xffffffffu, nz);
            const unsigned grp = __match_any_sync(0xffffffffu, in ? row : -1);
            if (in && (any & grp) && lane == __ffs(grp) - 1) atomicOr(&rownz[row], 1);
        }
    }
    __syncthreads();

    const int r0 = warp * kN2Rows;               // first output row of this thread, relative to y0
    // own luma pairs: staged rows r0+5 .. r0+10, staged columns 5+2l, 6+2l (copy A index 6+2l)
    f32x2 yo[kN2Rows + 4];
#pragma unroll
    for (int j = 0; j < kN2Rows + 4; ++j) yo[j] = lds2(n2sm + (r0 + 5 + j) * kN2RS + 6 + 2 * lane);

    f32x2 ac[3][kN2Rows], bc[3][kN2Rows], wsum2[kN2Rows], wd2[kN2Rows];
#pragma unroll
    for (int i = 0; i < kN2Rows; ++i) {
        wsum2[i] = pack2(0.f, 0.f);
        wd2[i] = pack2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 3; ++c) { ac[c][i] = pack2(0.f, 0.f); bc[c][i] = pack2(0.f, 0.f); }
    }

    bool live = false;
    for (int t = 0; t < kN2Rows + 2 * kNlmHalo; ++t) live |= (rownz[r0 + t] != 0);
    live = __any_sync(0xffffffffu, live);

    const f32x2 negk2 = pack2(negk, negk);
    const int lc = min(lane, 29);                // lanes 30, 31 only feed the box sums: keep their RGB reads in range
    // source offsets run +5 .. -5: terms are accumulated in the reference's order (denoise.py:106-109)
    for (int dx = live ? 5 : -6; dx >= -5; --dx) {
        const int odd = dx & 1;
        const float* cpy = n2sm + odd * kN2Copy;
        // pair starting at staged column s: copy A index s+1 (s odd), copy B index s (s even)
        const float* yp = cpy + r0 * kN2RS + (6 + 2 * lane + dx - odd);                      // luma rows r0 .. r0+15
        const float* cp = cpy + kN2Plane + (r0 + 2) * kN2RS + (8 + 2 * lc + dx - odd);       // RGB rows r0+2 .. r0+13
        f32x2 yw[kN2Rows + 14];
        f32x2 cw[3][kN2Rows + 10];
#pragma unroll
        for (int j = 0; j < kN2Rows + 4; ++j) yw[j + 10] = lds2(yp + (j + 10) * kN2RS);
#pragma unroll
        for (int i = 0; i < kN2Rows; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) cw[c][i + 10] = lds2(cp + c * kN2Plane + (i + 10) * kN2RS);
#pragma unroll
        for (int dy = 5; dy >= -5; --dy) {
            if (dy > -5) {   // the rows the next step slides in
                yw[dy + 4] = lds2(yp + (dy + 4) * kN2RS);
#pragma unroll
                for (int c = 0; c < 3; ++c) cw[c][dy + 4] = lds2(cp + c * kN2Plane + (dy + 4) * kN2RS);
            }
            f32x2 t[kN2Rows + 4];
#pragma unroll
            for (int j = 0; j < kN2Rows + 4; ++j) t[j] = sub2(yo[j], yw[j + dy + 5]);
            f32x2 m = mul2(t[1], t[1]);
            m = fma2(t[2], t[2], m);
            m = fma2(t[3], t[3], m);
            m = fma2(t[4], t[4], m);
            f32x2 v[kN2Rows];
            v[0] = fma2(t[0], t[0], m);
            v[1] = fma2(t[5], t[5], m);
#pragma unroll
            for (int i = 0; i < kN2Rows; ++i) {
                // windows of five columns starting at the lane's own even / odd column
                const float ve = lo2(v[i]), vo = hi2(v[i]);
                const float pr = ve + vo;
                const float s1 = __shfl_down_sync(0xffffffffu, pr, 1);
                const float s2 = __shfl_down_sync(0xffffffffu, pr, 2);
                const float e2 = __shfl_down_sync(0xffffffffu, ve, 2);
                const float ba = (pr + s1) + e2, bb = (vo + s1) + s2;
                // box >= 0: the reference's relu is a no-op
                const f32x2 dist = pack2(sqrt_approx(ba), sqrt_approx(bb));
                const f32x2 arg = mul2(dist, negk2);
                const f32x2 ww = pack2(ex2_approx(lo2(arg)), ex2_approx(hi2(arg)));
                wsum2[i] = add2(wsum2[i], ww);
                const int u = i + dy + 5;
                ac[0][i] = fma2(ww, cw[0][u], ac[0][i]);
                ac[1][i] = fma2(ww, cw[1][u], ac[1][i]);
                ac[2][i] = fma2(ww, cw[2][u], ac[2][i]);
                if (WITH_GRAD) {
                    const f32x2 wdi = mul2(ww, dist);
                    wd2[i] = add2(wd2[i], wdi);
                    bc[0][i] = fma2(wdi, cw[0][u], bc[0][i]);
                    bc[1][i] = fma2(wdi, cw[1][u], bc[1][i]);
                    bc[2][i] = fma2(wdi, cw[2][u], bc[2][i]);
                }
            }
        }
    }

    const int gx = x0 + 2 * lane;
    if (lane >= kN2TW / 2 || gx >= W) return;
    const float gcoef = (h > 0.f) ? 1.0f / (hh * hh) : 0.f;  // relu'(h) / hh^2
    const bool two = gx + 1 < W;
    const bool vec = two && vec8;                // 8-byte aligned pairs: even row length, aligned bases (host-checked)
#pragma unroll
    for (int i = 0; i < kN2Rows; ++i) {
        const int gy = y0 + r0 + i;
        if (gy >= H) continue;
        float ws[2] = {lo2(wsum2[i]), hi2(wsum2[i])};
        if (!live) ws[0] = ws[1] = 121.0f;       // all 121 weights are exactly 1
        const float wdv[2] = {lo2(wd2[i]), hi2(wd2[i])};
        const size_t px = (size_t)gy * W + gx;
        float yv[2][3], dv[2][3];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float iw = 1.0f / ws[k];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float a = k ? hi2(ac[c][i]) : lo2(ac[c][i]);
                float y = a * iw;
                if (variant == 0) {   // 0 * x: the (1 - mask) * img term of the reference's lerp (see nlm_kernel)
                    const float xin = (SEQ && pos > 0)
                                          ? n2sm[(1 + c) * kN2Plane + (r0 + i + kNlmHalo) * kN2RS + 8 + 2 * lane + k]
                                          : ((k == 0 || two) ? __ldg(src + (size_t)c * plane + px + k) : 0.f);
                    y = fmaf(0.f, xin, y);
                }
                yv[k][c] = clip01(y);
                if (WITH_GRAD) {
                    const float bcv = k ? hi2(bc[c][i]) : lo2(bc[c][i]);
                    dv[k][c] = pass01(y) * (bcv - y * wdv[k]) * iw * gcoef;
                }
            }
            if (SEQ) {
                for (int q = pos + 1; q < len; ++q) {
                    fwd_px<true>(ssop[q], ssc[q], yv[k][0], yv[k][1], yv[k][2]);
                    if (clip_each) { yv[k][0] = clip01(yv[k][0]); yv[k][1] = clip01(yv[k][1]); yv[k][2] = clip01(yv[k][2]); }
                }
            }
        }
        if (wsum_out) {
            float* wp = wsum_out + sb * plane + px;
            if (vec) *reinterpret_cast<float2*>(wp) = make_float2(ws[0], ws[1]);
            else { wp[0] = ws[0]; if (two) wp[1] = ws[1]; }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (!(variant < 2 || c == variant - 2)) continue;
            float* op = out + (size_t)b * 3 * plane + (size_t)c * plane + px;
            if (vec) *reinterpret_cast<float2*>(op) = make_float2(yv[0][c], yv[1][c]);
            else { op[0] = yv[0][c]; if (two) op[1] = yv[1][c]; }
            if (WITH_GRAD) {
                float* dp = dout_dh + sb * 3 * plane + (size_t)c * plane + px;
                if (vec) *reinterpret_cast<float2*>(dp) = make_float2(dv[0][c], dv[1][c]);
                else { dp[0] = dv[0][c]; if (two) dp[1] = dv[1][c]; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// NonLocalMeansParam (isp/denoise.py:122-157): the unfold / reflect-pad variant with one learnable h.
// Used nowhere in the reference, so this is a plain (not a fast) kernel: one output pixel per thread,
//   D(p, s) = sum_{o in window} dis(reflect(p + o), s),  dis(q, s) = (Y(q) - Y(reflect(q + s)))^2
// -- the reference pads the luma by reflection (:136), forms dis at the UNPADDED positions (:142), pads
// dis by reflection (:144) and box-sums it over the SEARCH window size (:145-146: the patch is as large
// as the search window) --, weights exp(-sqrt(D) / (relu(h) + EPS)), average of the reflect-padded rgb.
// `luma` is rgb_to_luminance(rgb) (the luma of the clipped image).  dout_dh: closed form as in nlm_kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int i, int n) {   // F.pad(mode='reflect'): -1 -> 1, n -> n - 2 (pad < n)
    i = i < 0 ? -i : i;
    return i >= n ? 2 * (n - 1) - i : i;
}

__global__ void __launch_bounds__(kThreads)
nlm_param_kernel(const float* __restrict__ rgb, const float* __restrict__ luma, float* __restrict__ out,
                 float* __restrict__ dout_dh, const float* __restrict__ hptr, int H, int W, int r) {
    pdl_prologue();
    const int b = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)H * W;
    const float* Y = luma + (size_t)b * plane;
    const float* X = rgb + (size_t)b * 3 * plane;
    const float h = hptr[0];
    const float hh = fmaxf(h, 0.f) + 1e-8f;
    float ws = 0.f, wd = 0.f, a[3] = {0.f, 0.f, 0.f}, bc[3] = {0.f, 0.f, 0.f};
    for (int sy = -r; sy <= r; ++sy)
        for (int sx = -r; sx <= r; ++sx) {
            float box = 0.f;
            for (int oy = -r; oy <= r; ++oy) {
                const int qy = reflect_idx(y + oy, H);
                const int ty = reflect_idx(qy + sy, H);
                for (int ox = -r; ox <= r; ++ox) {
                    const int qx = reflect_idx(x + ox, W);
                    const int tx = reflect_idx(qx + sx, W);
                    const float t = __ldg(Y + (size_t)qy * W + qx) - __ldg(Y + (size_t)ty * W + tx);
                    box = fmaf(t, t, box);
                }
            }
            const float d = sqrtf(fmaxf(box, 0.f));
            const float w = __expf(-d / hh);
            const size_t src = (size_t)reflect_idx(y + sy, H) * W + reflect_idx(x + sx, W);
            ws += w;
            wd = fmaf(w, d, wd);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = __ldg(X + (size_t)c * plane + src);
                a[c] = fmaf(w, v, a[c]);
                bc[c] = fmaf(w * d, v, bc[c]);
            }
        }
    const float iw = 1.0f / ws;
    const float gcoef = (h > 0.f) ? 1.0f / (hh * hh) : 0.f;
    const size_t px = (size_t)y * W + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = a[c] * iw;
        out[(size_t)b * 3 * plane + (size_t)c * plane + px] = clip01(v);
        if (dout_dh) dout_dh[(size_t)b * 3 * plane + (size_t)c * plane + px] = pass01(v) * (bc[c] - v * wd) * iw * gcoef;
    }
}

cudaError_t launch_nlm_param_fwd(const float* rgb, const float* luma, float* out, float* dout_dh, const float* h, int B,
                                 int H, int W, int window, cudaStream_t st) {
    dim3 grid((W + 31) / 32, (H + 7) / 8, B);
    launch_pdl(nlm_param_kernel, grid, kThreads, st, rgb, luma, out, dout_dh, h, H, W, window / 2);
    return cudaGetLastError();
}

// grad_h[b] = sum g * dout_dh : plain streaming dot product, chunked like the per-pixel kernels
__global__ void __launch_bounds__(kThreads)
nlm_dot_kernel(const float* __restrict__ gout, const float* __restrict__ stash, const int32_t* __restrict__ ops,
               long long n /* 3*H*W */, float* __restrict__ partial, BankMap bm, PooledGrad pool) {
    pdl_prologue();
    __shared__ float red[kWarps * AISP_ACC_STRIDE];
    const int b = bank_sample(bm, blockIdx.y);
    if (sample_op(ops, bm, b) != AISP_OP_NLM) return;
    const float* g = gout + (size_t)b * n;
    const float* s = stash + (size_t)(b / bm.F) * n;
    float acc[1] = {0.f};
    const long long lo = (long long)blockIdx.x * (3 * kPwChunkPx);
    const long long hi = min(lo + 3 * kPwChunkPx, n);
    const long long plane = n / 3;
    // the gradient of the pooled image at flat element i of the [3,H,W] sample
    auto pooled = [&](long long i) {
        const int ch = (i >= 2 * plane) ? 2 : (i >= plane ? 1 : 0);
        const int rem = (int)(i - ch * plane);
        return pooled_at(pool, b, ch, rem >> pool.ws, rem & ((1 << pool.ws) - 1));
    };
    if (((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(stash)) & 15u) == 0) {
        for (long long i = lo + 4 * threadIdx.x; i < hi; i += 4 * kThreads) {
            float4 a = ldg_stream4(g + i);
            const float4 c = ldg_stream4(s + i);
            if (pool.g) {   // (W % 4 == 0 here: the four elements share a row and a pooling block)
                const float u = pooled(i);
                a.x += u; a.y += u; a.z += u; a.w += u;
            }
            acc[0] += (a.x * c.x + a.y * c.y) + (a.z * c.z + a.w * c.w);
        }
    } else {
        for (long long i = lo + threadIdx.x; i < hi; i += kThreads)
            acc[0] = fmaf(pool.g ? g[i] + pooled(i) : g[i], s[i], acc[0]);
    }
    block_reduce_store<1>(acc, red, partial + ((size_t)b * gridDim.x + blockIdx.x) * AISP_ACC_STRIDE);
}

cudaError_t launch_finalize(const float* partial, int nrows, const float* params, const int32_t* ops, int family,
                            int B, float* grad_params, BankMap bm, cudaStream_t st);
int pointwise_rows(int H, int W);

// A/B switch for measurements: AISP_NLM_LAYOUT=1col selects the one-column-per-lane kernel (nlm_kernel)
static bool nlm_one_column_layout() {
    static const int v = [] {
        const char* e = getenv("AISP_NLM_LAYOUT");
        return (e && e[0] == '1') ? 1 : 0;
    }();
    return v != 0;
}

// one NLM forward over [B,3,H,W]: 28-wide tiles, a remainder of <= 12 columns goes to the half-warp layout
// (12 columns x 64 rows per CTA).  seq == true: per-sample sequences (see nlm_kernel<.., SEQ>).
static cudaError_t launch_nlm_any(const float* img, float* out, const float* params, const int32_t* ops, int B, int H,
                                  int W, float* dout_dh, float* wsum, BankMap bm, bool seq, const int32_t* seq_len,
                                  int S, int clip_each, int variant, cudaStream_t st) {
    if (!nlm_one_column_layout()) {
        // two-columns-per-lane layout: 60 x 16 tiles cover every width, no remainder launch
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(nlm2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kN2SmemBytes);
            cudaFuncSetAttribute(nlm2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kN2SmemBytes);
            cudaFuncSetAttribute(nlm2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kN2SmemBytes);
            attr_done = true;
        }
        const uintptr_t bases = reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout_dh) |
                                reinterpret_cast<uintptr_t>(wsum);
        const int vec8 = ((W & 1) == 0) && ((bases & 7u) == 0);
        dim3 grid((W + kN2TW - 1) / kN2TW, (H + kN2TH - 1) / kN2TH, B);
        if (seq)
            launch_pdl_smem(nlm2_kernel<false, true>, grid, kThreads, kN2SmemBytes, st, img, out, nullptr, nullptr, params,
                            ops, H, W, bm, seq_len, S, clip_each, variant, vec8);
        else if (dout_dh)
            launch_pdl_smem(nlm2_kernel<true, false>, grid, kThreads, kN2SmemBytes, st, img, out, dout_dh, wsum, params, ops,
                            H, W, bm, nullptr, 1, 0, variant, vec8);
        else
            launch_pdl_smem(nlm2_kernel<false, false>, grid, kThreads, kN2SmemBytes, st, img, out, nullptr, wsum, params, ops,
                            H, W, bm, nullptr, 1, 0, variant, vec8);
        return cudaGetLastError();
    }
    using G32 = NlmGeo<32>;
    using G16 = NlmGeo<16>;
    int n32 = W / G32::TW;
    const int rem = W - n32 * G32::TW;
    const bool half = rem > 0 && rem <= G16::TW && n32 > 0;
    if (rem > 0 && !half) ++n32;
    if (n32 > 0) {
        dim3 grid(n32, (H + G32::TH - 1) / G32::TH, B);
        if (seq)
            launch_pdl(nlm_kernel<false, 32, true>, grid, kThreads, st, img, out, nullptr, nullptr, params, ops, H, W, 0, bm,
                       seq_len, S, clip_each, variant);
        else if (dout_dh)
            launch_pdl(nlm_kernel<true, 32, false>, grid, kThreads, st, img, out, dout_dh, wsum, params, ops, H, W, 0, bm,
                       nullptr, 1, 0, variant);
        else
            launch_pdl(nlm_kernel<false, 32, false>, grid, kThreads, st, img, out, nullptr, wsum, params, ops, H, W, 0, bm,
                       nullptr, 1, 0, variant);
    }
    if (half) {
        dim3 grid(1, (H + G16::TH - 1) / G16::TH, B);
        const int x_off = n32 * G32::TW;
        if (seq)
            launch_pdl(nlm_kernel<false, 16, true>, grid, kThreads, st, img, out, nullptr, nullptr, params, ops, H, W, x_off,
                       bm, seq_len, S, clip_each, variant);
        else if (dout_dh)
            launch_pdl(nlm_kernel<true, 16, false>, grid, kThreads, st, img, out, dout_dh, wsum, params, ops, H, W, x_off, bm,
                       nullptr, 1, 0, variant);
        else
            launch_pdl(nlm_kernel<false, 16, false>, grid, kThreads, st, img, out, nullptr, wsum, params, ops, H, W, x_off,
                       bm, nullptr, 1, 0, variant);
    }
    return cudaGetLastError();
}

cudaError_t launch_nlm_fwd(const float* img, float* out, const float* params, const int32_t* ops, int B, int H, int W,
                           float* dout_dh, float* wsum, BankMap bm, cudaStream_t st) {
    return launch_nlm_any(img, out, params, ops, B, H, W, dout_dh, wsum, bm, false, nullptr, 1, 0, 0, st);
}

// the bare modules of isp/denoise.py (see `variant` at nlm_kernel): gray = NonLocalMeansGray, else NonLocalMeans
cudaError_t launch_nlm_module_fwd(const float* img, float* out, const float* params, const int32_t* ops, int B, int H,
                                  int W, float* dout_dh, int gray, cudaStream_t st) {
    if (gray) return launch_nlm_any(img, out, params, ops, B, H, W, dout_dh, nullptr, plain_batch(), false, nullptr, 1, 0, 1, st);
    for (int c = 0; c < 3; ++c) {
        cudaError_t e = launch_nlm_any(img, out, params, ops, B, H, W, dout_dh, nullptr, plain_batch(), false, nullptr, 1,
                                       0, 2 + c, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// per-sample sequences [prologue] -> NLM -> [epilogue] (no stashes: S > 1 has no closed-form d/dh)
cudaError_t launch_nlm_seq_fwd(const float* img, float* out, const float* params, const int32_t* ops,
                               const int32_t* seq_len, int B, int H, int W, int S, int clip_each, cudaStream_t st) {
    return launch_nlm_any(img, out, params, ops, B, H, W, nullptr, nullptr, plain_batch(), true, seq_len, S, clip_each, 0, st);
}

// ---------------------------------------------------------------------------------------------
// d L / d img of NLM.  Rare path (the reference's training never needs it: train.py:255), written
// for correctness and determinism, not speed: gather-only, no atomics.
//
// With U_c = gy_c / W, S = sum_c U_c y_c (gy = upstream gradient, W = sum of weights, y = output):
//   direct term      gxc_c(r) += sum_s w_s(r) * U_c(r+s)                (w_s(r-s) == w_-s(r))
//   through weights  a_s(p)   = sum_c U_c(p) x_c(p+s) - S(p)            d L / d w_s(p)
//                    delta_s(p) = -(w_s(p) / (2 hh d_s(p))) * (a_s(p) + a_-s(p+s)),   0 where box == 0
//                    gY(r)   += 2 (Y(r) - Y(r+s)) * box5x5[delta_s](r)
//   and gx = (gxc + luma * gY) * [0 <= x <= 1]   (the leading clip of DenoiseFilter.process).
// The (s, -s) pair shares one squared difference, which is why both directions fold into delta_s.
// CTA <-> (sample, 16x16 tile); per shift: phase A fills delta on tile+2, phase B box-sums it.
// ---------------------------------------------------------------------------------------------
constexpr int kGiT = 16, kGiA = kGiT + 14, kGiY = kGiT + 18, kGiD = kGiT + 4;

__global__ void __launch_bounds__(kThreads)
nlm_bwd_img_kernel(const float* __restrict__ img, const float* __restrict__ outp, const float* __restrict__ wsum,
                   const float* __restrict__ gout, const float* __restrict__ params, const int32_t* __restrict__ ops,
                   int H, int W, float* __restrict__ gimg) {
    pdl_prologue();
    __shared__ float sY[kGiY][kGiY + 1];
    __shared__ float sX[3][kGiA][kGiA + 1];
    __shared__ float sU[3][kGiA][kGiA + 1];
    __shared__ float sS[kGiA][kGiA + 1];
    __shared__ float sD[kGiD][kGiD + 1];
    __shared__ float sW[kGiD][kGiD + 1];
    const int b = blockIdx.z;
    if (ops[b] != AISP_OP_NLM) return;
    const int x0 = blockIdx.x * kGiT, y0 = blockIdx.y * kGiT;
    const size_t plane = (size_t)H * W;
    const float* src = img + (size_t)b * 3 * plane;
    const float* yo = outp + (size_t)b * 3 * plane;
    const float* go = gout + (size_t)b * 3 * plane;
    const float* ws = wsum + (size_t)b * plane;

    for (int e = threadIdx.x; e < kGiY * kGiY; e += kThreads) {
        const int row = e / kGiY, col = e - row * kGiY;
        const size_t off = (size_t)wrap(y0 - 9 + row, H) * W + wrap(x0 - 9 + col, W);
        const float r = clip01(__ldg(src + off)), g = clip01(__ldg(src + plane + off)),
                    bl = clip01(__ldg(src + 2 * plane + off));
        sY[row][col] = (0.299f * r + 0.587f * g) + 0.114f * bl;
    }
    for (int e = threadIdx.x; e < kGiA * kGiA; e += kThreads) {
        const int row = e / kGiA, col = e - row * kGiA;
        const size_t off = (size_t)wrap(y0 - 7 + row, H) * W + wrap(x0 - 7 + col, W);
        const float iw = 1.0f / __ldg(ws + off);
        float ssum = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float u = __ldg(go + c * plane + off) * iw;
            sX[c][row][col] = clip01(__ldg(src + c * plane + off));
            sU[c][row][col] = u;
            ssum = fmaf(u, __ldg(yo + c * plane + off), ssum);
        }
        sS[row][col] = ssum;
    }
    __syncthreads();

    const float h = params[(size_t)b * AISP_PSTRIDE];
    const float hh = fmaxf(h, 0.f) + 1e-8f;
    const float negk = -1.4426950408889634f / hh;
    const float half_inv_hh = 0.5f / hh;
    const int ty = threadIdx.x / kGiT, tx = threadIdx.x % kGiT;
    float gy_acc = 0.f, gc[3] = {0.f, 0.f, 0.f};

    for (int sx = -5; sx <= 5; ++sx) {
        for (int sy = -5; sy <= 5; ++sy) {
            for (int i = threadIdx.x; i < kGiD * kGiD; i += kThreads) {
                const int py = i / kGiD, px = i - py * kGiD;
                float box = 0.f;
#pragma unroll
                for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 5; ++kx) {
                        const float t = sY[py + ky + 5][px + kx + 5] - sY[py + ky + 5 + sy][px + kx + 5 + sx];
                        box = fmaf(t, t, box);
                    }
                const float d = sqrtf(box);
                const float w = exp2f(d * negk);
                const float kappa = (box > 0.f) ? w * half_inv_hh / d : 0.f;
                const int ay = py + 5, ax = px + 5;          // p in aux coordinates
                const int by = ay + sy, bx = ax + sx;        // p + s
                float afwd = -sS[ay][ax], arev = -sS[by][bx];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    afwd = fmaf(sU[c][ay][ax], sX[c][by][bx], afwd);
                    arev = fmaf(sU[c][by][bx], sX[c][ay][ax], arev);
                }
                sD[py][px] = -kappa * (afwd + arev);
                sW[py][px] = w;
            }
            __syncthreads();
            float gam = 0.f;
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) gam += sD[ty + ky][tx + kx];
            gy_acc = fmaf(2.0f * (sY[ty + 9][tx + 9] - sY[ty + 9 + sy][tx + 9 + sx]), gam, gy_acc);
            const float wr = sW[ty + 2][tx + 2];
#pragma unroll
            for (int c = 0; c < 3; ++c) gc[c] = fmaf(wr, sU[c][ty + 7 + sy][tx + 7 + sx], gc[c]);
            __syncthreads();
        }
    }
    const int gx = x0 + tx, gyy = y0 + ty;
    if (gx >= W || gyy >= H) return;
    const float lw[3] = {0.299f, 0.587f, 0.114f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const size_t o = (size_t)c * plane + (size_t)gyy * W + gx;
        gimg[(size_t)b * 3 * plane + o] = (gc[c] + lw[c] * gy_acc) * pass01(src[o]);
    }
}

cudaError_t launch_nlm_bwd_img(const float* img, const float* out, const float* wsum, const float* gout,
                               const float* params, const int32_t* ops, int B, int H, int W, float* gimg,
                               cudaStream_t st) {
    dim3 grid((W + kGiT - 1) / kGiT, (H + kGiT - 1) / kGiT, B);
    launch_pdl(nlm_bwd_img_kernel, grid, kThreads, st, img, out, wsum, gout, params, ops, H, W, gimg);
    return cudaGetLastError();
}

cudaError_t launch_nlm_bwd(const float* gout, const float* stash, const float* params_unused, const int32_t* ops,
                           int B, int H, int W, float* grad_params, float* partial, BankMap bm, PooledGrad pool,
                           cudaStream_t st) {
    (void)params_unused;
    const int rows = pointwise_rows(H, W);
    dim3 grid(rows, B);
    launch_pdl(nlm_dot_kernel, grid, kThreads, st, gout, stash, ops, 3LL * H * W, partial, bm, pool);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // finalize reads params only to derive constants; NLM needs none, so grad_params doubles as a
    // valid readable buffer of the right shape
    return launch_finalize(partial, rows, grad_params, ops, FAMILY_NLM, B, grad_params, bm, st);
}

}  // namespace aisp
