// Halo-tiled 3x3 sharpen (two variants) and 5x5 unsharp mask for sm_100a, forward + backward.
//
//   SHARPEN     y = clip(x*f + blur3(x)*(1-f))      isp/sharpen.py:105-142  (1-px border keeps x)
//   SHARPEN_V2  y = clip(x + (x - blur3(x))*f)      isp/sharpen.py:145-182
//   USM         y = clip(x + (x - G_sigma*x)*a)     isp/sharpen.py:84-102   (5x5, reflect padding)
//
// CTA <-> (sample, 128x16 tile).  The tile plus a 2-px halo of all three planes is staged in shared
// memory by ONE TMA bulk-tensor copy (cp.async.bulk.tensor.3d over a [B*3, H, W] tensor map, box
// 3 x 20 x 136 starting at column x0-4 -- the innermost TMA coordinate must be 16-byte aligned
// (measured: x0-2 raises an illegal-instruction fault) --, completion on an mbarrier): a single thread issues it, no warp spends instructions on
// addresses, and out-of-bounds halo elements arrive as zeros.  That is exactly right for the 3x3
// filters (frame pixels pass through and interior pixels never read outside the image) and for USM
// tiles that do not touch the image frame; USM frame tiles (reflect padding) and images whose rows
// are not 16-byte multiples take the cp.async path, which applies the reflect rule per element.
// Each thread then produces a 4x2 block per plane from registers (separable 5-tap passes for USM),
// so HBM sees ~24 B/px forward and ~24 B/px backward; halo re-reads by neighbouring CTAs hit L2.
#include <cstring>
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

// blur (and optionally d blur / d sigma) of a 4x2 block for one plane.
//   sm: plane base, SW: row stride, COFF: smem column of tile column 0; (bx, by): block origin in the tile.
template <bool USM, bool WITH_D, bool FRAME, int SW, int COFF>
__device__ __forceinline__ void block_blur(const float* sm, int bx, int by, const float* sc, int gx0, int gy0, int H,
                                           int W, float (&xc)[2][4], float (&blur)[2][4], float (&dblur)[2][4]) {
    if (USM) {
        float hk[6][4], hd[6][4];
        const float k0 = sc[0], k1 = sc[1], k2 = sc[2];
        const float d0 = sc[5], d1 = sc[6], d2 = sc[7];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = sm[(by + r) * SW + bx + COFF - 2 + i];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float e2 = v[i] + v[i + 4], e1 = v[i + 1] + v[i + 3], e0 = v[i + 2];
                hk[r][i] = fmaf(k0, e2, fmaf(k1, e1, k2 * e0));
                if (WITH_D) hd[r][i] = fmaf(d0, e2, fmaf(d1, e1, d2 * e0));
                if (r >= 2 && r < 4) xc[r - 2][i] = e0;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float e2 = hk[r][i] + hk[r + 4][i], e1 = hk[r + 1][i] + hk[r + 3][i], e0 = hk[r + 2][i];
                blur[r][i] = fmaf(k0, e2, fmaf(k1, e1, k2 * e0));
                if (WITH_D) {
                    const float f2 = hd[r][i] + hd[r + 4][i], f1 = hd[r + 1][i] + hd[r + 3][i], f0 = hd[r + 2][i];
                    dblur[r][i] = fmaf(d0, e2, fmaf(d1, e1, d2 * e0)) + fmaf(k0, f2, fmaf(k1, f1, k2 * f0));
                }
            }
    } else {
        const float a = 1.0f / 13.0f, bc = 5.0f / 13.0f;
        float hs[4][4], ctr[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float v[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) v[i] = sm[(by + 1 + r) * SW + bx + COFF - 1 + i];
#pragma unroll
            for (int i = 0; i < 4; ++i) { hs[r][i] = (v[i] + v[i + 1]) + v[i + 2]; ctr[r][i] = v[i + 1]; }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float x = ctr[r + 1][i];
                const float ring = (hs[r][i] + hs[r + 2][i]) + (hs[r + 1][i] - x);
                xc[r][i] = x;
                blur[r][i] = fmaf(a, ring, bc * x);
                if (FRAME) {  // only tiles that touch the image frame pay for the pass-through test
                    const int gx = gx0 + i, gy = gy0 + r;
                    if ((gx == 0) || (gy == 0) || (gx == W - 1) || (gy == H - 1)) blur[r][i] = x;
                }
                if (WITH_D) dblur[r][i] = 0.f;
            }
    }
}

__device__ __forceinline__ float sharpen_value(int op, float x, float blur, float f) {
    if (op == AISP_OP_SHARPEN) return x * f + blur * (1.0f - f);
    return x + (x - blur) * f;  // SHARPEN_V2 and USM
}

template <bool BWD, bool WRITE_GY>
__global__ void __launch_bounds__(kThreads, BWD ? 3 : 4)
sharpen_kernel(const __grid_constant__ CUtensorMap tmap, int tma_ok, const float* __restrict__ img,
               const float* __restrict__ gout, float* __restrict__ out, const float* __restrict__ params,
               const int32_t* __restrict__ ops, int H, int W, int vec, float* __restrict__ partial, BankMap bm,
               PooledGrad pool) {
    pdl_prologue();
    __shared__ __align__(128) float sm[kSmFloats];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ float sc[kConst];
    __shared__ float red[kWarps * AISP_ACC_STRIDE];
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
    const bool use_tma = tma_ok && !(op == AISP_OP_USM && usm_halo_out);
    if (use_tma) {
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, kTmaBytes);
            tma_load_3d(sm, &tmap, x0 - kColOff, y0 - kHalo, (b / bm.F) * 3, &bar);
        }
        // one wave ahead: the tile that a CTA scheduled ~one machine-fill later will load goes to L2 now
        if (threadIdx.x == 32) {
            const int ahead = 148 * (BWD ? 3 : 4);
            int lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x + ahead;
            const int px = lin % gridDim.x;
            lin /= gridDim.x;
            const int py = lin % gridDim.y, pz = lin / gridDim.y;
            if (pz < (int)gridDim.z) {
                const int pb = bank_sample(bm, pz);
                if (is_sharpen(sample_op(ops, bm, pb)))
                    tma_prefetch_3d(&tmap, px * kShTileW - kColOff, py * kShTileH - kHalo, (pb / bm.F) * 3);
            }
        }
    } else {
        stage_tile_cp(img + (size_t)(b / bm.F) * 3 * H * W, sm, H, W, x0, y0, vec != 0);
    }
    load_consts(params, b, op, sc);

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int bx = tx * 4, by = ty * 2;
    const int gx0 = x0 + bx, gy0 = y0 + by;
    const bool vec_ok = vec && (gx0 + 3 < W);  // vec: W % 4 == 0 and 16B-aligned global pointers
    float acc[2] = {0.f, 0.f};
    // upstream gradient of this thread's 3 x 2 x 4 outputs: requested before the tile has landed
    float gpre[3][2][4];
    if (BWD) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int gy = gy0 + r;
                const size_t off = base + ((size_t)ch * H + gy) * W + gx0;
                if (gy < H && vec_ok) {
                    const float4 t = ldg_stream4(gout + off);
                    gpre[ch][r][0] = t.x; gpre[ch][r][1] = t.y; gpre[ch][r][2] = t.z; gpre[ch][r][3] = t.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        gpre[ch][r][i] = (gy < H && gx0 + i < W) ? gout[off + i] : 0.f;
                }
                if (pool.g && gy < H) {   // + the gradient of the pooled image
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gx0 + i < W) gpre[ch][r][i] += pooled_at(pool, b, ch, gy, gx0 + i);
                }
            }
    }
    if (use_tma) {
        mbar_wait(&bar, 0);
    } else {
        cp_async_wait_all();
    }
    __syncthreads();
    const float f = (op == AISP_OP_USM) ? sc[10] : sc[0];

#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float xc[2][4], blur[2][4], dblur[2][4];
        const float* pl = sm + ch * kSmH * kCpW;
        if (op == AISP_OP_USM)
            block_blur<true, BWD, false, kCpW, kColOff>(pl, bx, by, sc, gx0, gy0, H, W, xc, blur, dblur);
        else if (frame_tile)
            block_blur<false, BWD, true, kCpW, kColOff>(pl, bx, by, sc, gx0, gy0, H, W, xc, blur, dblur);
        else
            block_blur<false, BWD, false, kCpW, kColOff>(pl, bx, by, sc, gx0, gy0, H, W, xc, blur, dblur);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int gy = gy0 + r;
            if (gy >= H || gx0 >= W) continue;
            const size_t off = base + ((size_t)ch * H + gy) * W + gx0;
            float y[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)   // 0 * x: the (1 - mask) * img term of the reference's lerp (NaN iff x is inf / NaN)
                y[i] = fmaf(0.f, xc[r][i], sharpen_value(op, xc[r][i], blur[r][i], f));
            if (!BWD) {
                if (vec_ok) {
                    stg_stream4(out + off, make_float4(clip01(y[0]), clip01(y[1]), clip01(y[2]), clip01(y[3])));
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gx0 + i < W) out[off + i] = clip01(y[i]);
                }
            } else {
                float g[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) g[i] = gpre[ch][r][i];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    g[i] *= pass01(y[i]);
                    if (gx0 + i < W) {
                        if (op == AISP_OP_USM) {
                            acc[0] = fmaf(g[i], dblur[r][i], acc[0]);
                            acc[1] = fmaf(g[i], xc[r][i] - blur[r][i], acc[1]);
                        } else {
                            acc[0] = fmaf(g[i], xc[r][i] - blur[r][i], acc[0]);
                        }
                    }
                }
                if (WRITE_GY) {
                    if (vec_ok) {
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
    if (BWD) {
        const int tile = blockIdx.y * gridDim.x + blockIdx.x;
        const int ntiles = gridDim.x * gridDim.y;
        block_reduce_store<2>(acc, red, partial + ((size_t)b * ntiles + tile) * AISP_ACC_STRIDE);
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
    const bool use_tma = J.tma_ok && !(op == AISP_OP_USM && usm_halo_out);
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
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float xc[2][4], blur[2][4], dblur[2][4];
        const float* pl = sm + ch * kSmH * kCpW;
        if (op == AISP_OP_USM)
            block_blur<true, false, false, kCpW, kColOff>(pl, bx, by, scs, gx0, gy0, H, W, xc, blur, dblur);
        else if (frame_tile)
            block_blur<false, false, true, kCpW, kColOff>(pl, bx, by, scs, gx0, gy0, H, W, xc, blur, dblur);
        else
            block_blur<false, false, false, kCpW, kColOff>(pl, bx, by, scs, gx0, gy0, H, W, xc, blur, dblur);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                y[ch][r][i] = clip01(fmaf(0.f, xc[r][i], sharpen_value(op, xc[r][i], blur[r][i], f)));
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
// transposed stencil: grad_img from the masked upstream gradient gy (staged by the kernel above)
// CTA <-> (sample, 128x8 tile); a thread owns a 1x4 strip of all three planes and evaluates the transposed
// stencil separably from registers (5 rows x 8 staged values per plane): the 5x5 Gaussian is k (x) k, the
// 3x3 kernel is (ones3x3 + 4 delta) / 13.  Reflect-padding mirrors (USM, pixels within 2 of the frame) are
// folded in per pixel from global memory -- a few pixels per row.
// ---------------------------------------------------------------------------------------------
constexpr int kAdjW = 128, kAdjH = 8, kAdjSW = kAdjW + 8;   // staged row: tile column 0 at smem column 4

__device__ __forceinline__ float gy_valid(const float* __restrict__ gy, int op, int H, int W, int y, int x) {
    // gradient of output pixel (y,x) that flows through its blur term
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.f;
    if (op != AISP_OP_USM && (y == 0 || x == 0 || y == H - 1 || x == W - 1)) return 0.f;  // border: blur == x
    return gy[(size_t)y * W + x];
}

__device__ inline float gpad_at(const float* __restrict__ gy, int op, const float (*w)[5], int H, int W, int y,
                                int x) {
    // sum_k K(k) * gy(y - k): transposed correlation evaluated at a (possibly padded) position
    float s = 0.f;
    for (int dy = -2; dy <= 2; ++dy)
        for (int dx = -2; dx <= 2; ++dx) s = fmaf(w[dy + 2][dx + 2], gy_valid(gy, op, H, W, y - dy, x - dx), s);
    return s;
}

__device__ __forceinline__ int mirrors(int q, int n, int* m) {
    // padded indices q' in [-2, n+1] whose reflect source is q (besides q itself)
    int c = 0;
    if (q >= 1 && q <= 2) m[c++] = -q;
    const int hi = 2 * (n - 1) - q;
    if (hi >= n && hi <= n + 1) m[c++] = hi;
    return c;
}

__global__ void __launch_bounds__(kThreads)
sharpen_adjoint_kernel(const float* __restrict__ gy, float* __restrict__ gimg, const float* __restrict__ params,
                       const int32_t* __restrict__ ops, int H, int W) {
    pdl_prologue();
    __shared__ float sm[3][kAdjH + 4][kAdjSW];
    __shared__ float sc[kConst];
    __shared__ float wk[5][5];
    const int b = blockIdx.z;
    const int op = ops[b];
    if (!is_sharpen(op)) return;
    const int x0 = blockIdx.x * kAdjW, y0 = blockIdx.y * kAdjH;
    const size_t base = (size_t)b * 3 * H * W;
    load_consts(params, b, op, sc);
    __syncthreads();
    if (threadIdx.x < 25) {   // dense 5x5 weights: only the mirror terms of frame pixels use them
        const int i = threadIdx.x / 5, j = threadIdx.x % 5;
        float v;
        if (op == AISP_OP_USM) v = sc[i] * sc[j];
        else v = (i == 0 || i == 4 || j == 0 || j == 4) ? 0.f : ((i == 2 && j == 2) ? 5.0f / 13.0f : 1.0f / 13.0f);
        wk[i][j] = v;
    }
    constexpr int kCols = kAdjW + 4;   // staged columns x0-2 .. x0+kAdjW+1 live at smem columns 2 .. kAdjW+5
    for (int e = threadIdx.x; e < 3 * (kAdjH + 4) * kCols; e += kThreads) {
        const int ch = e / ((kAdjH + 4) * kCols);
        const int rem = e - ch * ((kAdjH + 4) * kCols);
        const int row = rem / kCols, col = rem - row * kCols;
        sm[ch][row][col + 2] = gy_valid(gy + base + (size_t)ch * H * W, op, H, W, y0 - 2 + row, x0 - 2 + col);
    }
    __syncthreads();
    const int lx = (threadIdx.x & 31) * 4, ly = threadIdx.x >> 5;
    const int xs = x0 + lx, y = y0 + ly;
    if (xs >= W || y >= H) return;
    float alpha, beta;
    if (op == AISP_OP_SHARPEN) { alpha = sc[0]; beta = 1.0f - sc[0]; }
    else if (op == AISP_OP_SHARPEN_V2) { alpha = 1.0f + sc[0]; beta = -sc[0]; }
    else { alpha = 1.0f + sc[10]; beta = -sc[10]; }
    const bool usm = (op == AISP_OP_USM);
    const float k0 = sc[0], k1 = sc[1], k2 = sc[2];
    int my[2];
    const int nmy = usm ? mirrors(y, H, my) : 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float* gch = gy + base + (size_t)ch * H * W;
        // rows y-2 .. y+2, columns xs-2 .. xs+5 of the masked gradient
        float v[5][8];
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[r][i] = sm[ch][ly + r][lx + 2 + i];
        float s[4];
        if (usm) {
            float hrow[5][4];
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    hrow[r][i] = fmaf(k0, v[r][i] + v[r][i + 4], fmaf(k1, v[r][i + 1] + v[r][i + 3], k2 * v[r][i + 2]));
#pragma unroll
            for (int i = 0; i < 4; ++i)
                s[i] = fmaf(k0, hrow[0][i] + hrow[4][i], fmaf(k1, hrow[1][i] + hrow[3][i], k2 * hrow[2][i]));
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float box = 0.f;
#pragma unroll
                for (int r = 1; r < 4; ++r) box += (v[r][i + 1] + v[r][i + 2]) + v[r][i + 3];
                s[i] = fmaf(4.0f / 13.0f, v[2][i + 2], box * (1.0f / 13.0f));
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = xs + i;
            if (x >= W) continue;
            float si = s[i];
            if (usm) {   // reflect padding folds the halo of the padded plane back onto its mirror pixels
                int mx[2];
                const int nmx = mirrors(x, W, mx);
                for (int a = 0; a < nmy; ++a) si += gpad_at(gch, op, wk, H, W, my[a], x);
                for (int c2 = 0; c2 < nmx; ++c2) si += gpad_at(gch, op, wk, H, W, y, mx[c2]);
                for (int a = 0; a < nmy; ++a)
                    for (int c2 = 0; c2 < nmx; ++c2) si += gpad_at(gch, op, wk, H, W, my[a], mx[c2]);
            }
            const float g0 = gch[(size_t)y * W + x];
            float out = alpha * g0 + beta * si;
            const bool border = (x == 0) || (y == 0) || (x == W - 1) || (y == H - 1);
            if (!usm && border) out += beta * g0;
            gimg[base + ((size_t)ch * H + y) * W + x] = out;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
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

static bool make_tile_map(CUtensorMap* map, const float* img, int B, int H, int W) {
    memset(map, 0, sizeof(*map));
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || (W & 3) != 0 || !al16(img) || W < 4) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * 3};
    const cuuint64_t gstr[2] = {(cuuint64_t)W * sizeof(float), (cuuint64_t)W * H * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)kTmaW, (cuuint32_t)kSmH, 3};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(img), gdim, gstr, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch_sharpen_fwd(const float* img, float* out, const float* params, const int32_t* ops, int B, int H,
                               int W, BankMap bm, cudaStream_t st) {
    dim3 grid((W + kShTileW - 1) / kShTileW, (H + kShTileH - 1) / kShTileH, B);
    const int vec = ((W & 3) == 0) && al16(img) && al16(out);
    CUtensorMap map;
    const int tma_ok = make_tile_map(&map, img, bm.n ? B / bm.n : B, H, W) ? 1 : 0;
    launch_pdl(sharpen_kernel<false, false>, grid, kThreads, st, map, tma_ok, img, nullptr, out, params, ops, H, W, vec,
               nullptr, bm, no_pooled_grad());
    return cudaGetLastError();
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
    CUtensorMap map;
    const int tma_ok = make_tile_map(&map, img, bm.n ? B / bm.n : B, H, W) ? 1 : 0;
    if (grad_img)
        launch_pdl(sharpen_kernel<true, true>, grid, kThreads, st, map, tma_ok, img, gout, gy_scratch, params, ops, H, W, vec,
                   partial, bm, pool);
    else
        launch_pdl(sharpen_kernel<true, false>, grid, kThreads, st, map, tma_ok, img, gout, nullptr, params, ops, H, W, vec,
                   partial, bm, pool);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = launch_finalize(partial, sharpen_rows(H, W), params, ops, FAMILY_SHARPEN, B, grad_params, bm, st);
    if (e != cudaSuccess) return e;
    if (grad_img) {
        dim3 g2((W + kAdjW - 1) / kAdjW, (H + kAdjH - 1) / kAdjH, B);
        launch_pdl(sharpen_adjoint_kernel, g2, kThreads, st, gy_scratch, grad_img, params, ops, H, W);
        e = cudaGetLastError();
    }
    return e;
}

}  // namespace aisp
