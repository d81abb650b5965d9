// Halo-tiled 3x3 sharpen (two variants) and 5x5 unsharp mask for sm_100a, forward + backward.
//
//   SHARPEN     y = clip(x*f + blur3(x)*(1-f))      isp/sharpen.py:105-142  (1-px border keeps x)
//   SHARPEN_V2  y = clip(x + (x - blur3(x))*f)      isp/sharpen.py:145-182
//   USM         y = clip(x + (x - G_sigma*x)*a)     isp/sharpen.py:84-102   (5x5, reflect padding)
//
// CTA <-> (sample, 128x16 tile).  The tile plus a 2-px halo of all three planes is staged in shared
// memory by ONE TMA bulk-tensor copy (cp.async.bulk.tensor.3d over a [B*3, H, W] tensor map, box
// 3 x 20 x 136 starting at column x0-4 -- the innermost TMA coordinate must be 16-byte aligned
// (measured: x0-2 raises an illegal-instruction fault) --, completion on an mbarrier): a single thread
// issues it, no warp spends instructions on addresses, and out-of-bounds halo elements arrive as zeros.  That is exactly right for the 3x3
// filters (frame pixels pass through and interior pixels never read outside the image) and for USM
// tiles whose halo stays inside the image; a USM tile whose halo leaves it (reflect padding) copies the
// few out-of-image columns / rows from their mirror positions inside the same tile after the copy has
// landed (usm_reflect_fixup); images whose rows are not 16-byte multiples take the cp.async path, which
// applies the reflect rule per element.
// Each thread then produces a 4x2 block per plane from registers as packed fp32 pairs (separable 5-tap
// passes for USM), so HBM sees ~24 B/px forward and ~24 B/px backward; halo re-reads by neighbouring CTAs
// hit L2.  Backward: the upstream-gradient tile rides on a second TMA copy into shared memory.
#include <cstring>

#include <cuda.h>  // CUtensorMap and enums only; cuTensorMapEncodeTiled is resolved through the runtime

#include "pointwise_math.cuh"   // fwd_px / stage_consts for the fused per-pixel prologue and epilogue (-fmad=false)

namespace aisp {

constexpr int kHalo = 2;
constexpr int kSmH = kShTileH + 2 * kHalo;           // 20 rows
constexpr int kCpW = kShTileW + 8;                   // 136 columns per staged row, tile column 0 at smem column 4:
constexpr int kColOff = 4;                           //   a TMA box must start on a 16-byte boundary of the row
constexpr int kTmaW = kCpW;                          //   (x0 - 4), which also keeps cp.async 16-byte aligned
constexpr int kSmFloats = 3 * kSmH * kCpW;
constexpr unsigned kTmaBytes = 3u * kSmH * kTmaW * sizeof(float);

__device__ __forceinline__ int reflect_clamp(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return min(max(i, 0), n - 1);
}

// ---- TMA + mbarrier primitives (PTX; SASS: UTMALDG / SYNCS) ----------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "AISP_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra AISP_MBAR_DONE;\n\t"
        "bra AISP_MBAR_WAIT;\n"
        "AISP_MBAR_DONE:\n\t}" ::"r"(a), "r"(parity) : "memory");
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

// L2 prefetch of a tile through the same tensor map (no shared memory, no barrier)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// ---- cp.async (LDGSTS) fallback staging --------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// a warp copies whole rows: one 16-byte async copy per lane for the 128 interior columns, lanes
// 0..3 fetch the four halo columns; the border rule (reflect, clamped) is applied on the way in.
__device__ __forceinline__ void stage_tile_cp(const float* __restrict__ img, float* sm, int H, int W, int x0, int y0,
                                              bool vec) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gx = x0 + 4 * lane;
    const bool own_vec = vec && (gx + 3 < W);
    const int hc = (lane < 2) ? lane - 2 : kShTileW - 2 + lane;   // halo column of lanes 0..3: -2,-1,128,129
    const int hx = reflect_clamp(x0 + hc, W);
    for (int rr = warp; rr < 3 * kSmH; rr += kWarps) {
        const int ch = rr / kSmH, row = rr - ch * kSmH;
        const float* src = img + ((size_t)ch * H + reflect_clamp(y0 - kHalo + row, H)) * W;
        float* dst = sm + (ch * kSmH + row) * kCpW + 4 + 4 * lane;
        if (own_vec) {
            cp_async16(dst, src + gx);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) cp_async4(dst + i, src + reflect_clamp(gx + i, W));
        }
        if (lane < 4) cp_async4(sm + (ch * kSmH + row) * kCpW + 4 + hc, src + hx);
    }
}

// Reflect padding on top of a TMA-staged tile (isp/sharpen.py:76-78, F.pad(mode='reflect') by 2): TMA zero-fills the
// halo elements that lie outside the image; the (at most two) columns and rows beyond each image edge are then copied
// from their mirror positions, which the same tile holds -- columns first, then whole rows (so corners come out right).
// Called by every thread of the CTA after the tile has landed; ends with a barrier.
__device__ __forceinline__ void usm_reflect_fixup(float* sm, int x0, int y0, int H, int W) {
    const int xs = x0 - kColOff, ys = y0 - kHalo;            // image coordinates of staged column 0 / row 0
    // columns x in {-2, -1, W, W+1}
    for (int e = threadIdx.x; e < 3 * kSmH * 4; e += kThreads) {
        const int which = e & 3, rowp = e >> 2;              // rowp = plane * kSmH + row
        const int x = (which < 2) ? which - 2 : W + which - 2;
        const int src = (x < 0) ? -x : 2 * (W - 1) - x;
        const int cx = x - xs, cs = src - xs;
        if (cx >= 0 && cx < kCpW && cs >= 0 && cs < kCpW && src >= 0 && src < W) sm[rowp * kCpW + cx] = sm[rowp * kCpW + cs];
    }
    __syncthreads();
    // rows y in {-2, -1, H, H+1}
    for (int e = threadIdx.x; e < 3 * 4 * kCpW; e += kThreads) {
        const int col = e % kCpW, t = e / kCpW;
        const int which = t & 3, pl = t >> 2;
        const int y = (which < 2) ? which - 2 : H + which - 2;
        const int src = (y < 0) ? -y : 2 * (H - 1) - y;
        const int ry = y - ys, rs = src - ys;
        if (ry >= 0 && ry < kSmH && rs >= 0 && rs < kSmH && src >= 0 && src < H)
            sm[(pl * kSmH + ry) * kCpW + col] = sm[(pl * kSmH + rs) * kCpW + col];
    }
    __syncthreads();
}

__device__ __forceinline__ void load_consts(const float* __restrict__ params, int b, int op, float* sc /*smem*/) {
    if (threadIdx.x == 0) {
        float raw[kConst];
        for (int k = 0; k < AISP_PSTRIDE; ++k) raw[k] = params[(size_t)b * AISP_PSTRIDE + k];
        for (int k = AISP_PSTRIDE; k < kConst; ++k) raw[k] = 0.f;
        float c[kConst];
        for (int k = 0; k < kConst; ++k) c[k] = 0.f;
        derive_consts(op, raw, c);
        for (int k = 0; k < kConst; ++k) sc[k] = c[k];
    }
}

// ---------------------------------------------------------------------------------------------
// Packed-pair stencil bodies (round 2).  A thread's four output columns are the ALIGNED pairs
// (c0, c1), (c2, c3) of its staged row, so every per-pixel quantity is a natural f32x2 and the blur,
// its sigma-derivative, the sharpen value and the gradient products run as FADD2 / FMUL2 / FFMA2 -- one
// issue slot per two pixels.  Taps at even offsets of a pair are pairs again; the +-1 taps are the
// mixed sums (c-1 + c1, c0 + c2), formed by two scalar adds whose results the register allocator
// places side by side.  USM runs the vertical 5-tap pass first (pure pair arithmetic on the loaded
// columns), then the horizontal one on two rows.
//   pl: plane base + by * kCpW + bx + kColOff (column c0 of the thread's first staged row, 16-byte aligned)
// ---------------------------------------------------------------------------------------------
template <bool WITH_D>
__device__ __forceinline__ void usm_block2(const float* pl, const float* sc, f32x2 (&xc)[2][2], f32x2 (&blur)[2][2],
                                           f32x2 (&dblur)[2][2]) {
    const f32x2 k0 = splat2(sc[0]), k1 = splat2(sc[1]), k2 = splat2(sc[2]);
    const f32x2 d0 = splat2(sc[5]), d1 = splat2(sc[6]), d2 = splat2(sc[7]);
    f32x2 VK[2][4], VD[2][4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {   // column pairs (c-2, c-1), (c0, c1), (c2, c3), (c4, c5)
        f32x2 V[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) V[r] = lds2(pl + r * kCpW - 2 + 2 * p);
        if (p == 1 || p == 2) { xc[0][p - 1] = V[2]; xc[1][p - 1] = V[3]; }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const f32x2 e2 = add2(V[r], V[r + 4]), e1 = add2(V[r + 1], V[r + 3]), e0 = V[r + 2];
            VK[r][p] = fma2(k0, e2, fma2(k1, e1, mul2(k2, e0)));
            if (WITH_D) VD[r][p] = fma2(d0, e2, fma2(d1, e1, mul2(d2, e0)));
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const f32x2 L = VK[r][q], C = VK[r][q + 1], R = VK[r][q + 2];
            const f32x2 e2 = add2(L, R);
            const f32x2 e1 = pack2(hi2(L) + hi2(C), lo2(C) + lo2(R));
            blur[r][q] = fma2(k0, e2, fma2(k1, e1, mul2(k2, C)));
            if (WITH_D) {   // d(kv (x) kh) = dkv (x) kh + kv (x) dkh
                const f32x2 hd = fma2(d0, e2, fma2(d1, e1, mul2(d2, C)));
                const f32x2 LD = VD[r][q], CD = VD[r][q + 1], RD = VD[r][q + 2];
                const f32x2 f2 = add2(LD, RD);
                const f32x2 f1 = pack2(hi2(LD) + hi2(CD), lo2(CD) + lo2(RD));
                dblur[r][q] = add2(hd, fma2(k0, f2, fma2(k1, f1, mul2(k2, CD))));
            }
        }
}

// frame pixels of a thread's 4x2 block as a bit mask (bit 4r + i: output (r, i) lies on the image frame)
__device__ __forceinline__ unsigned frame_bits(int gx0, int gy0, int H, int W) {
    unsigned m = 0;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gy = gy0 + r, gx = gx0 + i;
            if ((gy == 0) || (gy == H - 1) || (gx == 0) || (gx == W - 1)) m |= 1u << (4 * r + i);
        }
    return m;
}

// 3x3: vertical 3-row sums of the column pairs first (rows 1, 2 of the window are shared by both output
// rows), then the horizontal sum; ring = box9 - x, blur = ring / 13 + 5 x / 13 (isp/sharpen.py:105-142)
template <bool FRAME>
__device__ __forceinline__ void box3_block2(const float* pl, unsigned fbits, f32x2 (&xc)[2][2], f32x2 (&blur)[2][2]) {
    const f32x2 a = splat2(1.0f / 13.0f), bc = splat2(5.0f / 13.0f);
    f32x2 VS[2][4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {   // column pairs (c-2, c-1), (c0, c1), (c2, c3), (c4, c5)
        f32x2 V[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) V[r] = lds2(pl + (1 + r) * kCpW - 2 + 2 * p);
        if (p == 1 || p == 2) { xc[0][p - 1] = V[1]; xc[1][p - 1] = V[2]; }
        const f32x2 m = add2(V[1], V[2]);
        VS[0][p] = add2(V[0], m);
        VS[1][p] = add2(m, V[3]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const f32x2 L = VS[r][q], C = VS[r][q + 1], R = VS[r][q + 2];
            const f32x2 x = xc[r][q];
            const f32x2 box = add2(pack2(hi2(L) + hi2(C), lo2(C) + lo2(R)), C);
            f32x2 bl = fma2(a, sub2(box, x), mul2(bc, x));
            if (FRAME) {  // only threads that own a frame pixel pay for the pass-through select
                const unsigned two = (fbits >> (4 * r + 2 * q)) & 3u;
                if (two) bl = pack2((two & 1u) ? lo2(x) : lo2(bl), (two & 2u) ? hi2(x) : hi2(bl));
            }
            blur[r][q] = bl;
        }
}

// One thread's 4x2 block of all three planes.  FULL: the tile lies inside the image, rows are 16-byte
// multiples, the upstream gradient sits in shared memory and no pooled gradient is folded in.
// KIND: 0 = 3x3 away from the frame, 1 = 3x3 on a frame tile, 2 = 5x5 USM.
template <bool BWD, bool WRITE_GY, bool FULL, int KIND>
__device__ __forceinline__ void sharpen_tile(const float* tile, const float* gtile, const float* sc, int op, int H, int W,
                                             int gx0, int gy0, int b, size_t base, bool vec_ok, bool g_tma,
                                             const float* __restrict__ gout, float* __restrict__ out,
                                             const PooledGrad& pool, float (&acc)[2]) {
    constexpr bool usm = (KIND == 2);
    const float fs = usm ? sc[10] : sc[0];
    const f32x2 f = splat2(fs), omf = splat2(1.0f - fs), zero2 = splat2(0.f);
    const bool blend = (op == AISP_OP_SHARPEN);          // x*f + blur*(1-f); otherwise x + (x - blur)*f
    const unsigned fbits = (KIND == 1) ? frame_bits(gx0, gy0, H, W) : 0u;
    f32x2 acc0 = zero2, acc1 = zero2;
#pragma unroll 1
    for (int ch = 0; ch < 3; ++ch) {
        f32x2 xc[2][2], blur[2][2], dblur[2][2];
        const float* pl = tile + ch * kSmH * kCpW;
        if (usm) usm_block2<BWD>(pl, sc, xc, blur, dblur);
        else box3_block2<KIND == 1>(pl, fbits, xc, blur);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int gy = gy0 + r;
            const size_t off = base + ((size_t)ch * H + gy) * W + gx0;
            const bool row_ok = FULL || ((gy < H) && (gx0 < W));
            f32x2 y[2], xmb[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const f32x2 x = xc[r][q];
                xmb[q] = sub2(x, blur[r][q]);
                // products rounded, then added, as the reference's ATen expressions are
                const f32x2 v = (!usm && blend) ? add2_sep(mul2(x, f), mul2(blur[r][q], omf)) : add2_sep(x, mul2(xmb[q], f));
                y[q] = fma2(zero2, x, v);   // 0 * x: the (1 - mask) * img term of the reference's lerp (NaN iff x is inf / NaN)
            }
            if (!BWD) {
                if (!row_ok) continue;
                const float o0 = clip01(lo2(y[0])), o1 = clip01(hi2(y[0])), o2 = clip01(lo2(y[1])), o3 = clip01(hi2(y[1]));
                if (FULL || vec_ok) {
                    stg_stream4(out + off, make_float4(o0, o1, o2, o3));
                } else {
                    const float o[4] = {o0, o1, o2, o3};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gx0 + i < W) out[off + i] = o[i];
                }
            } else {
                float g[4];
                if (FULL || g_tma) {   // staged by TMA: zero beyond the image
                    const float4 t = *reinterpret_cast<const float4*>(gtile + (ch * kShTileH + r) * kShTileW);
                    g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) g[i] = (row_ok && gx0 + i < W) ? gout[off + i] : 0.f;
                }
                if (!FULL && pool.g && row_ok) {   // + the gradient of the pooled image
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gx0 + i < W) g[i] += pooled_at(pool, b, ch, gy, gx0 + i);
                }
                g[0] *= pass01(lo2(y[0])); g[1] *= pass01(hi2(y[0]));
                g[2] *= pass01(lo2(y[1])); g[3] *= pass01(hi2(y[1]));
                const f32x2 ga = pack2(g[0], g[1]), gb = pack2(g[2], g[3]);
                if (usm) {
                    acc0 = fma2(ga, dblur[r][0], acc0); acc0 = fma2(gb, dblur[r][1], acc0);
                    acc1 = fma2(ga, xmb[0], acc1); acc1 = fma2(gb, xmb[1], acc1);
                } else {
                    acc0 = fma2(ga, xmb[0], acc0); acc0 = fma2(gb, xmb[1], acc0);
                }
                if (WRITE_GY && row_ok) {
                    if (FULL || vec_ok) {
                        stg_stream4(out + off, make_float4(g[0], g[1], g[2], g[3]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (gx0 + i < W) out[off + i] = g[i];
                    }
                }
            }
        }
    }
    acc[0] = lo2(acc0) + hi2(acc0);
    acc[1] = lo2(acc1) + hi2(acc1);
}

// dynamic shared memory of sharpen_kernel: image tile + halo | upstream-gradient tile (backward) | constants | barrier.
// 57,296 bytes in the backward: FOUR CTAs per SM fit (4 x (57,296 + 1 KB reserved) <= 228 KB), which is what the
// one-shot CTA (load -> wait -> compute -> reduce) needs to keep enough tiles in flight; the reduction needs no shared
// memory (per-warp entries of the scratch row) and only the 16 constants the sharpen family reads are kept.
constexpr int kGoFloats = 3 * kShTileH * kShTileW;                        // 24 KB
constexpr unsigned kGoBytes = kGoFloats * sizeof(float);
constexpr int kShConst = 16;
constexpr size_t kShSmemFwd = (size_t)(kSmFloats + kShConst) * sizeof(float) + 16;
constexpr size_t kShSmemBwd = (size_t)(kSmFloats + kGoFloats + kShConst) * sizeof(float) + 16;
static_assert(kShSmemBwd <= 57344, "four backward CTAs per SM");
static_assert(kWarps * 4 == AISP_ACC_STRIDE, "one scratch row holds four floats per warp");

// constants of one sample into shared memory (thread 0; the caller synchronises)
__device__ __forceinline__ void load_consts16(const float* __restrict__ params, int b, int op, float* sc /*smem, kShConst*/) {
    if (threadIdx.x == 0) {
        float raw[kConst];
        for (int k = 0; k < AISP_PSTRIDE; ++k) raw[k] = params[(size_t)b * AISP_PSTRIDE + k];
        for (int k = AISP_PSTRIDE; k < kConst; ++k) raw[k] = 0.f;
        float c[kConst];
        for (int k = 0; k < kConst; ++k) c[k] = 0.f;
        derive_consts(op, raw, c);
        for (int k = 0; k < kShConst; ++k) sc[k] = c[k];
    }
}

template <bool BWD, bool WRITE_GY>
__global__ void __launch_bounds__(kThreads, 4)
sharpen_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap gmap, int tma_ok,
               const float* __restrict__ img, const float* __restrict__ gout, float* __restrict__ out,
               const float* __restrict__ params, const int32_t* __restrict__ ops, int H, int W, int vec,
               float* __restrict__ partial, BankMap bm, PooledGrad pool) {
    pdl_prologue();
    extern __shared__ __align__(128) float shsm[];
    float* sm = shsm;                                            // [3][kSmH][kCpW]
    float* gs = shsm + kSmFloats;                                // [3][kShTileH][kShTileW]  (backward, TMA path)
    float* sc = shsm + kSmFloats + (BWD ? kGoFloats : 0);        // [kShConst]
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(sc + kShConst);
    const int b = bank_sample(bm, blockIdx.z);   // filter-bank launches: see BankMap
    const int op = sample_op(ops, bm, b);
    if (!is_sharpen(op)) return;
    const int x0 = blockIdx.x * kShTileW, y0 = blockIdx.y * kShTileH;
    const size_t base = (size_t)b * 3 * H * W;
    const bool frame_tile = (x0 == 0) || (y0 == 0) || (x0 + kShTileW >= W) || (y0 + kShTileH >= H);  // CTA-uniform
    // zero-filled out-of-bounds halos are only wrong for reflect padding: a USM tile takes the
    // reflecting cp.async path as soon as its 2-px HALO leaves the image (H = y0 + 17 puts halo row
    // y0 + 17 == H outside although the tile itself ends one row short of the frame)
    const bool usm_halo_out = (x0 < kHalo) || (y0 < kHalo) || (x0 + kShTileW + kHalo > W) || (y0 + kShTileH + kHalo > H);
    const bool use_tma = tma_ok != 0;   // USM tiles whose halo leaves the image: zero fill now, mirrored below
    // the upstream gradient rides on TMA whenever the launch can (its zero fill makes rows / columns past the
    // image contribute nothing); tma_ok covers both maps
    const bool g_tma = BWD && tma_ok;
    if (tma_ok) {
        // two arrivals complete the phase: the TMA issue (with its byte count) and the constants (below), so that a
        // thread that has seen the phase flip may read tile, gradient tile AND constants -- no CTA barrier after the load
        if (threadIdx.x == 0) mbar_init(bar, 2);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (use_tma ? kTmaBytes : 0u) + (g_tma ? kGoBytes : 0u));
            if (use_tma) tma_load_3d(sm, &tmap, x0 - kColOff, y0 - kHalo, (b / bm.F) * 3, bar);
            if (g_tma) tma_load_3d(gs, &gmap, x0, y0, b * 3, bar);
        }
        // one wave ahead: the tiles that a CTA scheduled ~one machine-fill later will load go to L2 now
        if (threadIdx.x == 32) {
            const int ahead = 148 * 4;
            int lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x + ahead;
            const int px = lin % gridDim.x;
            lin /= gridDim.x;
            const int py = lin % gridDim.y, pz = lin / gridDim.y;
            if (pz < (int)gridDim.z) {
                const int pb = bank_sample(bm, pz);
                if (is_sharpen(sample_op(ops, bm, pb))) {
                    tma_prefetch_3d(&tmap, px * kShTileW - kColOff, py * kShTileH - kHalo, (pb / bm.F) * 3);
                    if (BWD) tma_prefetch_3d(&gmap, px * kShTileW, py * kShTileH, pb * 3);
                }
            }
        }
    }
    if (!use_tma) stage_tile_cp(img + (size_t)(b / bm.F) * 3 * H * W, sm, H, W, x0, y0, vec != 0);
    load_consts16(params, b, op, sc);
    if (tma_ok && threadIdx.x == 0) mbar_arrive(bar);   // release: the constants are written

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int bx = tx * 4, by = ty * 2;
    const int gx0 = x0 + bx, gy0 = y0 + by;
    const bool vec_ok = vec && (gx0 + 3 < W);  // vec: W % 4 == 0 and 16B-aligned global pointers
    if (tma_ok) {
        mbar_wait(bar, 0);
    } else {
        cp_async_wait_all();
        __syncthreads();
    }
    if (use_tma && op == AISP_OP_USM && usm_halo_out) {   // CTA-uniform
        __syncthreads();                                  // every thread has seen the tile before cells are rewritten
        usm_reflect_fixup(sm, x0, y0, H, W);
    }
    float acc[2] = {0.f, 0.f};
    // CTA-uniform specialisation: whole tile inside the image with 128-bit accesses (no per-row / per-column
    // tests), and the stencil kind (5x5 USM, 3x3 away from the frame, 3x3 on the frame)
    const bool full = vec && (x0 + kShTileW <= W) && (y0 + kShTileH <= H) && (!BWD || g_tma) && !pool.g;
    const int kind = (op == AISP_OP_USM) ? 2 : (frame_tile ? 1 : 0);
    const float* tile = sm + by * kCpW + bx + kColOff;
    const float* gtile = gs + by * kShTileW + bx;
#define AISP_SHARPEN_TILE(FULL, KIND)                                                                                    \
    sharpen_tile<BWD, WRITE_GY, FULL, KIND>(tile, gtile, sc, op, H, W, gx0, gy0, b, base, vec_ok, g_tma, gout, out, pool, acc)
    if (full) {
        if (kind == 2) AISP_SHARPEN_TILE(true, 2);
        else if (kind == 1) AISP_SHARPEN_TILE(true, 1);
        else AISP_SHARPEN_TILE(true, 0);
    } else {
        if (kind == 2) AISP_SHARPEN_TILE(false, 2);
        else if (kind == 1) AISP_SHARPEN_TILE(false, 1);
        else AISP_SHARPEN_TILE(false, 0);
    }
#undef AISP_SHARPEN_TILE
    if (BWD) {
        // no CTA barrier: every warp reduces its two sums by shuffles and owns four floats (two sums, two zeros) of the
        // tile's scratch row; finalize_kernel adds the eight warps' entries in fixed order (FAMILY_SHARPEN)
        const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
        const int ntiles = gridDim.x * gridDim.y;
        float a0 = acc[0], a1 = acc[1];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        }
        if (tx < 4)
            partial[((size_t)b * ntiles + tile_id) * AISP_ACC_STRIDE + ty * 4 + tx] = (tx == 0) ? a0 : ((tx == 1) ? a1 : 0.f);
    }
}


// ---------------------------------------------------------------------------------------------
// Sequence forward: [per-pixel prologue] -> 3x3 sharpen / USM -> [per-pixel epilogue] in ONE launch
// (the fixed chain of isp/filters.py:753-815, E -> G -> WB -> CCM -> Shr, and the replayed pipelines
// of yolov3/val_adaptiveisp.py:291-327 whose sequence holds one stencil step).
//   * the tile + halo lands in shared memory exactly as in sharpen_kernel (TMA, or the reflecting
//     cp.async path); the prologue steps are then applied IN PLACE to the staged tile (halo included:
//     per-pixel filters commute with the halo's coordinate map), 2720 staged pixels per 2048 outputs;
//   * the stencil produces the thread's 3 x 2 x 4 outputs in registers, the epilogue steps run on
//     those registers, and only the final values are stored: 24 B/px for the whole sequence;
//   * two jobs per launch: the low-resolution batch and its high-resolution twin (same parameters,
//     isp/filters.py:116-122, agent.py:155-157) are tiles of the same grid (blockIdx.x >= tiles of
//     job 0 -> job 1, with its own tensor map);
//   * EMIT: the 64x64 block means of job 0's OUTPUT (agent.py:97 / value.py:63 pool it next) leave
//     through the store path: per-thread sums -> shared memory -> one thread per block, fixed order.
// ---------------------------------------------------------------------------------------------
struct SeqJob {
    const float* img;
    float* out;
    int H, W;
    int tiles_x, tiles;   // tiles per row / per image (tiles == 0: job absent)
    int tma_ok, vec;
};

template <bool EMIT>
__global__ void __launch_bounds__(kThreads, 3)
sharpen_seq_fwd_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1, SeqJob j0,
                       SeqJob j1, const float* __restrict__ params, const int32_t* __restrict__ ops,
                       const int32_t* __restrict__ seq_len, int S, int clip_each, float* __restrict__ down, int bh,
                       int bw) {
    pdl_prologue();
    __shared__ __align__(128) float sm[kSmFloats];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ float raw[AISP_MAX_STEPS][kConst];
    __shared__ float sc[AISP_MAX_STEPS][kConst];
    __shared__ int sop[AISP_MAX_STEPS];
    __shared__ float bs[EMIT ? 3 * kThreads : 1];
    const int b = blockIdx.z;
    int len = seq_len ? min(max(seq_len[b], 0), S) : S;
    const int pos = find_stencil(ops + (size_t)b * S, len, &len);
    if (pos < 0 || !is_sharpen(ops[(size_t)b * S + pos])) return;   // another family owns this sample
    const bool second = (int)blockIdx.x >= j0.tiles;
    const SeqJob& J = second ? j1 : j0;
    const CUtensorMap* tmap = second ? &tmap1 : &tmap0;
    const int tile = second ? (int)blockIdx.x - j0.tiles : (int)blockIdx.x;
    const int H = J.H, W = J.W;
    const int x0 = (tile % J.tiles_x) * kShTileW, y0 = (tile / J.tiles_x) * kShTileH;
    const size_t base = (size_t)b * 3 * H * W;
    const int op = ops[(size_t)b * S + pos];
    const bool frame_tile = (x0 == 0) || (y0 == 0) || (x0 + kShTileW >= W) || (y0 + kShTileH >= H);
    const bool usm_halo_out = (x0 < kHalo) || (y0 < kHalo) || (x0 + kShTileW + kHalo > W) || (y0 + kShTileH + kHalo > H);
    const bool use_tma = J.tma_ok != 0;
    if (use_tma) {
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, kTmaBytes);
            tma_load_3d(sm, tmap, x0 - kColOff, y0 - kHalo, b * 3, &bar);
        }
    } else {
        stage_tile_cp(J.img + base, sm, H, W, x0, y0, J.vec != 0);
    }
    stage_consts(params, ops, b, S, len, raw, sc, sop, BankMap{1, 0, 0ull, 0ull});   // ends with a barrier
    if (use_tma) mbar_wait(&bar, 0);
    else cp_async_wait_all();
    __syncthreads();
    if (use_tma && op == AISP_OP_USM && usm_halo_out) usm_reflect_fixup(sm, x0, y0, H, W);   // CTA-uniform

    // prologue: steps 0 .. pos-1 on every staged pixel (tile + halo), in place
    if (pos > 0) {
        constexpr int kPlane = kSmH * kCpW;
        for (int e = threadIdx.x; e < kPlane; e += kThreads) {
            float r = sm[e], g = sm[kPlane + e], bl = sm[2 * kPlane + e];
            for (int k = 0; k < pos; ++k) {
                fwd_px<true>(sop[k], sc[k], r, g, bl);
                if (clip_each) { r = clip01(r); g = clip01(g); bl = clip01(bl); }
            }
            sm[e] = r; sm[kPlane + e] = g; sm[2 * kPlane + e] = bl;
        }
        __syncthreads();
    }

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int bx = tx * 4, by = ty * 2;
    const int gx0 = x0 + bx, gy0 = y0 + by;
    const bool vec_ok = J.vec && (gx0 + 3 < W);
    const float* scs = sc[pos];
    const float f = (op == AISP_OP_USM) ? scs[10] : scs[0];
    float y[3][2][4];
    {   // same packed-pair bodies (and therefore the same bits) as sharpen_kernel
        const unsigned fbits = (op != AISP_OP_USM && frame_tile) ? frame_bits(gx0, gy0, H, W) : 0u;
        const f32x2 f2v = splat2(f), omf = splat2(1.0f - f), zero2 = splat2(0.f);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            f32x2 xc[2][2], blur[2][2], dblur[2][2];
            const float* pl = sm + ch * kSmH * kCpW + by * kCpW + bx + kColOff;
            if (op == AISP_OP_USM) usm_block2<false>(pl, scs, xc, blur, dblur);
            else if (frame_tile) box3_block2<true>(pl, fbits, xc, blur);
            else box3_block2<false>(pl, fbits, xc, blur);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const f32x2 x = xc[r][q];
                    const f32x2 v = (op == AISP_OP_SHARPEN) ? add2_sep(mul2(x, f2v), mul2(blur[r][q], omf))
                                                            : add2_sep(x, mul2(sub2(x, blur[r][q]), f2v));
                    const f32x2 yy = fma2(zero2, x, v);
                    y[ch][r][2 * q] = clip01(lo2(yy));
                    y[ch][r][2 * q + 1] = clip01(hi2(yy));
                }
        }
    }
    // epilogue: steps pos+1 .. len-1 on the thread's own outputs
    for (int k = pos + 1; k < len; ++k) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                fwd_px<true>(sop[k], sc[k], y[0][r][i], y[1][r][i], y[2][r][i]);
                if (clip_each) { y[0][r][i] = clip01(y[0][r][i]); y[1][r][i] = clip01(y[1][r][i]); y[2][r][i] = clip01(y[2][r][i]); }
            }
    }
    float psum[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int gy = gy0 + r;
            if (gy >= H || gx0 >= W) continue;
            const size_t off = base + ((size_t)ch * H + gy) * W + gx0;
            if (vec_ok) {
                stg_stream4(J.out + off, make_float4(y[ch][r][0], y[ch][r][1], y[ch][r][2], y[ch][r][3]));
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (gx0 + i < W) J.out[off + i] = y[ch][r][i];
            }
            if (EMIT) psum[ch] += (y[ch][r][0] + y[ch][r][1]) + (y[ch][r][2] + y[ch][r][3]);
        }
    if (EMIT) {
        // (the host only asks for EMIT when W % bw == 0, H % bh == 0, bw % 4 == 0, bh % 2 == 0 and both
        //  divide the tile: a thread's 4 x 2 outputs then lie in one pooling block, blocks in one tile)
        if (second) return;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) bs[ch * kThreads + threadIdx.x] = psum[ch];
        __syncthreads();
        const int nbx = kShTileW / bw, nby = kShTileH / bh;      // pooling blocks per tile
        const int tpb_x = bw / 4, tpb_y = bh / 2;                 // threads per block in x / y
        const int ow = W / bw, oh = H / bh;
        for (int e = threadIdx.x; e < 3 * nbx * nby; e += kThreads) {
            const int ch = e / (nbx * nby), rem = e - ch * (nbx * nby);
            const int pby = rem / nbx, pbx = rem - pby * nbx;
            const int oy = y0 / bh + pby, ox = x0 / bw + pbx;
            if (oy >= oh || ox >= ow) continue;
            float s = 0.f;
            for (int yy = 0; yy < tpb_y; ++yy)
                for (int xx = 0; xx < tpb_x; ++xx)
                    s += bs[ch * kThreads + (pby * tpb_y + yy) * 32 + pbx * tpb_x + xx];
            down[(((size_t)b * 3 + ch) * oh + oy) * ow + ox] = s / (float)(bh * bw);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// transposed stencil: grad_img from the masked upstream gradient gy (written by sharpen_kernel<true, true>)
// Same geometry as the forward: CTA <-> (sample, 128x16 tile); tile + 2-px halo of the three gy planes is
// staged by ONE TMA copy -- its zero fill outside the image is exactly the zero extension a transposed
// correlation needs -- and a thread produces a 4x2 block per plane from registers, 128-bit stores.
//   * 5x5 USM is k (x) k.  Reflect padding is separable too (rows are padded, then columns), so its adjoint
//     folds per axis: output index q collects the padded positions q' that reflect onto it,
//         w_q(rho) = k(q - rho) + sum_{q' in mirrors(q)} k(q' - rho),   mirrors: 1 -> -1, 2 -> -2, n-2 -> n, n-3 -> n+1,
//     i.e. outputs 1, 2, n-3, n-2 of an axis get one or two extra taps (k1 gy(0) + k0 gy(1); k0 gy(0);
//     k0 gy(n-1); k0 gy(n-2) + k1 gy(n-1)) on top of the plain separable pass -- a few predicated FMAs in the
//     threads next to the frame; tiles away from the frame never test for them.
//   * the 3x3 kernel is (ones3x3 + 4 delta) / 13; frame pixels pass x through (isp/sharpen.py:105-142), so
//     their gy does not flow through the blur (masked as it is read) and reaches grad_img directly.
// ---------------------------------------------------------------------------------------------
// staging without TMA (rows that are not 16-byte multiples): zero fill outside the image
__device__ __forceinline__ void stage_tile_zero(const float* __restrict__ src, float* sm, int H, int W, int x0, int y0) {
    constexpr int kPlane = kSmH * kCpW;
    for (int e = threadIdx.x; e < 3 * kPlane; e += kThreads) {
        const int ch = e / kPlane, rem = e - ch * kPlane;
        const int row = rem / kCpW, col = rem - row * kCpW;
        const int y = y0 - kHalo + row, x = x0 - kColOff + col;
        sm[e] = (y >= 0 && y < H && x >= 0 && x < W) ? src[((size_t)ch * H + y) * W + x] : 0.f;
    }
}

// extra taps of the reflect fold for output index q of an axis of length n; t[j] is the value at q - 2 + j
__device__ __forceinline__ float fold_extra(int q, int n, float k0, float k1, float t0, float t1, float t2, float t3,
                                            float t4) {
    float e = 0.f;
    if (q == 1) e = fmaf(k1, t1, fmaf(k0, t2, e));          // q' = -1 reads gy(0), gy(1)
    if (q == 2) e = fmaf(k0, t0, e);                        // q' = -2 reads gy(0)
    if (q == n - 2) e = fmaf(k0, t2, fmaf(k1, t3, e));      // q' = n reads gy(n-2), gy(n-1)
    if (q == n - 3) e = fmaf(k0, t4, e);                    // q' = n+1 reads gy(n-1)
    return e;
}

template <bool FRAME>
__device__ __forceinline__ void adjoint_block(const float* sm, const float* sc, int op, int H, int W, int gx0, int gy0,
                                              int bx, int by, float* __restrict__ dst /* sample base of grad_img */,
                                              bool vec_ok) {
    const bool usm = (op == AISP_OP_USM);
    float alpha, beta;
    if (op == AISP_OP_SHARPEN) { alpha = sc[0]; beta = 1.0f - sc[0]; }
    else if (op == AISP_OP_SHARPEN_V2) { alpha = 1.0f + sc[0]; beta = -sc[0]; }
    else { alpha = 1.0f + sc[10]; beta = -sc[10]; }
    const float k0 = sc[0], k1 = sc[1], k2 = sc[2];
    // per-thread: does the 4x2 block hold an output that collects mirror taps (USM) ...
    const bool near_x = FRAME && usm && ((gx0 <= 2) || (gx0 + 3 >= W - 3));
    const bool near_y = FRAME && usm && ((gy0 <= 2) || (gy0 + 1 >= H - 3));
    if (!FRAME) {
        // away from the frame the transposed stencil IS the forward blur (both kernels are symmetric) applied to gy:
        // the packed-pair bodies of sharpen_kernel, then grad = alpha * gy + beta * blur(gy)
        const f32x2 a2 = splat2(alpha), b2 = splat2(beta);
#pragma unroll 1
        for (int ch = 0; ch < 3; ++ch) {
            f32x2 gc[2][2], bl[2][2], unused[2][2];
            const float* pc = sm + ch * kSmH * kCpW + by * kCpW + bx + kColOff;
            if (usm) usm_block2<false>(pc, sc, gc, bl, unused);
            else box3_block2<false>(pc, 0u, gc, bl);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int gy = gy0 + r;
                if (gy >= H) continue;
                const f32x2 o0 = add2_sep(mul2(a2, gc[r][0]), mul2(b2, bl[r][0]));
                const f32x2 o1 = add2_sep(mul2(a2, gc[r][1]), mul2(b2, bl[r][1]));
                float* q = dst + ((size_t)ch * H + gy) * W + gx0;
                if (vec_ok) {
                    stg_stream4(q, make_float4(lo2(o0), hi2(o0), lo2(o1), hi2(o1)));
                } else {
                    const float o[4] = {lo2(o0), hi2(o0), lo2(o1), hi2(o1)};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gx0 + i < W) q[i] = o[i];
                }
            }
        }
        return;
    }
#pragma unroll 1   // rolled: a third of the code (the unrolled kernel stalled on instruction fetch)
    for (int ch = 0; ch < 3; ++ch) {
        const float* pl = sm + ch * kSmH * kCpW + by * kCpW + bx + kColOff - 2;   // 8-byte aligned
        float hrow[6][4], g0[2][4], ctr[2][4];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const float2 t = *reinterpret_cast<const float2*>(pl + r * kCpW + i);
                v[i] = t.x; v[i + 1] = t.y;
            }
            if (r >= 2 && r < 4) {
#pragma unroll
                for (int i = 0; i < 4; ++i) g0[r - 2][i] = v[i + 2];
            }
            if (usm) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    hrow[r][i] = fmaf(k0, v[i] + v[i + 4], fmaf(k1, v[i + 1] + v[i + 3], k2 * v[i + 2]));
                if (near_x) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        hrow[r][i] += fold_extra(gx0 + i, W, k0, k1, v[i], v[i + 1], v[i + 2], v[i + 3], v[i + 4]);
                }
            } else {
                if (FRAME) {   // ... or a frame pixel, whose gy does not flow through the blur (3x3)
                    const int y = gy0 - 2 + r;
                    const bool rowf = (y == 0) || (y == H - 1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int x = gx0 - 2 + i;
                        v[i] = (rowf || (x == 0) || (x == W - 1)) ? 0.f : v[i];
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) hrow[r][i] = (v[i + 1] + v[i + 2]) + v[i + 3];
                if (r >= 2 && r < 4) {   // centre tap: the (masked) value itself
#pragma unroll
                    for (int i = 0; i < 4; ++i) ctr[r - 2][i] = v[i + 2];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int gy = gy0 + r;
            if (gy >= H) continue;
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float s;
                if (usm) {
                    s = fmaf(k0, hrow[r][i] + hrow[r + 4][i], fmaf(k1, hrow[r + 1][i] + hrow[r + 3][i], k2 * hrow[r + 2][i]));
                    if (near_y) s += fold_extra(gy, H, k0, k1, hrow[r][i], hrow[r + 1][i], hrow[r + 2][i], hrow[r + 3][i], hrow[r + 4][i]);
                } else {
                    const float box = (hrow[r + 1][i] + hrow[r + 2][i]) + hrow[r + 3][i];
                    s = fmaf(4.0f / 13.0f, ctr[r][i], box * (1.0f / 13.0f));
                }
                o[i] = alpha * g0[r][i] + beta * s;
                if (FRAME && !usm) {
                    const int x = gx0 + i;
                    if ((x == 0) || (gy == 0) || (x == W - 1) || (gy == H - 1)) o[i] += beta * g0[r][i];
                }
            }
            float* q = dst + ((size_t)ch * H + gy) * W + gx0;
            if (vec_ok) {
                stg_stream4(q, make_float4(o[0], o[1], o[2], o[3]));
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (gx0 + i < W) q[i] = o[i];
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads, 3)
sharpen_adjoint_kernel(const __grid_constant__ CUtensorMap tmap, int tma_ok, const float* __restrict__ gy,
                       float* __restrict__ gimg, const float* __restrict__ params, const int32_t* __restrict__ ops, int H,
                       int W, int vec) {
    pdl_prologue();
    __shared__ __align__(128) float sm[kSmFloats];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ float sc[kConst];
    const int b = blockIdx.z;
    const int op = ops[b];
    if (!is_sharpen(op)) return;
    const int x0 = blockIdx.x * kShTileW, y0 = blockIdx.y * kShTileH;
    const size_t base = (size_t)b * 3 * H * W;
    if (tma_ok) {
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, kTmaBytes);
            tma_load_3d(sm, &tmap, x0 - kColOff, y0 - kHalo, b * 3, &bar);
        }
    } else {
        stage_tile_zero(gy + base, sm, H, W, x0, y0);
    }
    load_consts(params, b, op, sc);
    if (tma_ok) mbar_wait(&bar, 0);
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int bx = tx * 4, by = ty * 2;
    const int gx0 = x0 + bx, gy0 = y0 + by;
    if (gx0 >= W || gy0 >= H) return;
    const bool vec_ok = vec && (gx0 + 3 < W);
    // CTA-uniform: does the tile hold an output within 3 pixels of the image frame?  (USM outputs 1, 2, n-3, n-2 of an
    // axis collect mirror taps; 3x3 outputs next to the frame read masked frame pixels)
    const bool frame = (x0 <= kHalo) || (y0 <= kHalo) || (x0 + kShTileW + kHalo >= W) || (y0 + kShTileH + kHalo >= H);
    if (frame) adjoint_block<true>(sm, sc, op, H, W, gx0, gy0, bx, by, gimg + base, vec_ok);
    else adjoint_block<false>(sm, sc, op, H, W, gx0, gy0, bx, by, gimg + base, vec_ok);
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sharpen_rows(int H, int W) {
    return ((W + kShTileW - 1) / kShTileW) * ((H + kShTileH - 1) / kShTileH);
}

cudaError_t launch_finalize(const float* partial, int nrows, const float* params, const int32_t* ops, int family,
                            int B, float* grad_params, BankMap bm, cudaStream_t st);

// [B*3, H, W] fp32 tensor map with a 3 x 20 x 136 box.  Returns false when TMA cannot describe the
// image (rows not a multiple of 16 bytes, unaligned base, too many planes) -> cp.async path.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static bool make_tile_map(CUtensorMap* map, const float* img, int B, int H, int W, int box_w = kTmaW, int box_h = kSmH) {
    memset(map, 0, sizeof(*map));
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || (W & 3) != 0 || !al16(img) || W < 4) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * 3};
    const cuuint64_t gstr[2] = {(cuuint64_t)W * sizeof(float), (cuuint64_t)W * H * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 3};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(img), gdim, gstr, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool BWD, bool WRITE_GY>
static void sharpen_attrs() {
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !done[dev]) {
        cudaFuncSetAttribute(sharpen_kernel<BWD, WRITE_GY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(BWD ? kShSmemBwd : kShSmemFwd));
        done[dev] = true;
    }
}

cudaError_t launch_sharpen_fwd(const float* img, float* out, const float* params, const int32_t* ops, int B, int H,
                               int W, BankMap bm, cudaStream_t st) {
    dim3 grid((W + kShTileW - 1) / kShTileW, (H + kShTileH - 1) / kShTileH, B);
    const int vec = ((W & 3) == 0) && al16(img) && al16(out);
    CUtensorMap map;
    const int tma_ok = make_tile_map(&map, img, bm.n ? B / bm.n : B, H, W) ? 1 : 0;
    sharpen_attrs<false, false>();
    return launch_pdl_smem(sharpen_kernel<false, false>, grid, kThreads, kShSmemFwd, st, map, map, tma_ok, img, nullptr, out,
                           params, ops, H, W, vec, nullptr, bm, no_pooled_grad());
}

// Sequence forward over one or two jobs (job 1 = the high-resolution twin, hr_img == nullptr: absent).
// down != nullptr: block means of job 0's output, when the pooling blocks tile the 128 x 16 CTA tile
// (the caller checks with sharpen_can_emit and otherwise runs the stand-alone block-mean pass).
bool sharpen_can_emit(int H, int W, int oh, int ow) {
    if (oh <= 0 || ow <= 0 || H % oh || W % ow) return false;
    const int bh = H / oh, bw = W / ow;
    return (bw % 4 == 0) && (bh % 2 == 0) && (kShTileW % bw == 0) && (kShTileH % bh == 0);
}

cudaError_t launch_sharpen_seq_fwd(const float* img, float* out, const float* params, const int32_t* ops,
                                   const int32_t* seq_len, int B, int H, int W, int S, int clip_each,
                                   const float* hr_img, float* hr_out, int hr_H, int hr_W, float* down, int oh, int ow,
                                   cudaStream_t st) {
    CUtensorMap map0, map1;
    SeqJob j0{}, j1{};
    j0.img = img; j0.out = out; j0.H = H; j0.W = W;
    j0.tiles_x = (W + kShTileW - 1) / kShTileW;
    j0.tiles = j0.tiles_x * ((H + kShTileH - 1) / kShTileH);
    j0.vec = ((W & 3) == 0) && al16(img) && al16(out);
    j0.tma_ok = make_tile_map(&map0, img, B, H, W) ? 1 : 0;
    memset(&map1, 0, sizeof(map1));
    if (hr_img) {
        j1.img = hr_img; j1.out = hr_out; j1.H = hr_H; j1.W = hr_W;
        j1.tiles_x = (hr_W + kShTileW - 1) / kShTileW;
        j1.tiles = j1.tiles_x * ((hr_H + kShTileH - 1) / kShTileH);
        j1.vec = ((hr_W & 3) == 0) && al16(hr_img) && al16(hr_out);
        j1.tma_ok = make_tile_map(&map1, hr_img, B, hr_H, hr_W) ? 1 : 0;
    }
    dim3 grid((unsigned)(j0.tiles + j1.tiles), 1, (unsigned)B);
    if (down)
        launch_pdl(sharpen_seq_fwd_kernel<true>, grid, kThreads, st, map0, map1, j0, j1, params, ops, seq_len, S, clip_each,
                   down, H / oh, W / ow);
    else
        launch_pdl(sharpen_seq_fwd_kernel<false>, grid, kThreads, st, map0, map1, j0, j1, params, ops, seq_len, S,
                   clip_each, nullptr, 1, 1);
    return cudaGetLastError();
}

cudaError_t launch_sharpen_bwd(const float* img, const float* gout, const float* params, const int32_t* ops, int B,
                               int H, int W, float* grad_params, float* grad_img, float* gy_scratch, float* partial,
                               BankMap bm, PooledGrad pool, cudaStream_t st) {
    dim3 grid((W + kShTileW - 1) / kShTileW, (H + kShTileH - 1) / kShTileH, B);
    const int vec = ((W & 3) == 0) && al16(img) && al16(gout) && (!grad_img || al16(gy_scratch));
    CUtensorMap map, gmap;
    // the upstream gradient is [virtual samples, 3, H, W]: B samples here, images x F filters in a bank launch
    const int gplanes = bm.n ? (B / bm.n) * bm.F : B;
    const int tma_ok = (make_tile_map(&map, img, bm.n ? B / bm.n : B, H, W) &&
                        make_tile_map(&gmap, gout, gplanes, H, W, kShTileW, kShTileH)) ? 1 : 0;
    if (grad_img) {
        sharpen_attrs<true, true>();
        launch_pdl_smem(sharpen_kernel<true, true>, grid, kThreads, kShSmemBwd, st, map, gmap, tma_ok, img, gout, gy_scratch,
                        params, ops, H, W, vec, partial, bm, pool);
    } else {
        sharpen_attrs<true, false>();
        launch_pdl_smem(sharpen_kernel<true, false>, grid, kThreads, kShSmemBwd, st, map, gmap, tma_ok, img, gout, nullptr,
                        params, ops, H, W, vec, partial, bm, pool);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = launch_finalize(partial, sharpen_rows(H, W), params, ops, FAMILY_SHARPEN, B, grad_params, bm, st);
    if (e != cudaSuccess) return e;
    if (grad_img) {
        CUtensorMap gmap;
        const int gtma = make_tile_map(&gmap, gy_scratch, B, H, W) ? 1 : 0;
        const int gvec = ((W & 3) == 0) && al16(grad_img);
        launch_pdl(sharpen_adjoint_kernel, grid, kThreads, st, gmap, gtma, gy_scratch, grad_img, params, ops, H, W, gvec);
        e = cudaGetLastError();
    }
    return e;
}

}  // namespace aisp
