// Fused per-pixel ISP pass for sm_100a: exposure, gamma, white balance, CCM, tone / colour curves,
// contrast, saturation+, desaturation -- forward over a per-sample op sequence in one pass over HBM,
// and the single-step backward (parameter gradients always, image gradient on request).
//
// Layout: CTA <-> (sample b = blockIdx.y, chunk of kPwChunkPx pixels = blockIdx.x).  The op id is
// uniform per CTA, so the `switch` never diverges.  Each thread streams 16 pixels as 3 planes x
// 4 x 128-bit loads (12 LDG.128 in flight), computes in registers, and streams the result back.
// HBM-bound: 24 B/px forward, 24 B/px backward (36 with grad_img).  No shared-memory staging is
// needed for the pixels (no reuse); shared memory only holds the per-step derived constants.
#include "pointwise_math.cuh"

namespace aisp {

// =============================================================================================
// kernels
// =============================================================================================
// Extras of a "sequence launch set" (aisp_sequence_fwd); all zero for the plain entry points.
struct PwExt {
    const float* img1;   // high-resolution twin (isp/filters.py:116-122, agent.py:155-157): the same per-sample
    float* out1;         //   sequence and parameters applied to a second image set in the SAME launch --
    int N1;              //   chunks >= chunks0 of the grid belong to it
    int chunks0;
    float* down;         // block means of job 0's OUTPUT [B,3,H/bh,W/bw] leave through the store path (agent.py:97 /
    int W, bh, bw;       //   value.py:63 pool the retouched image next); nullptr: not requested
    int owned;           // 1: samples whose sequence holds a stencil step belong to a stencil kernel of the set
};

template <int VEC, bool EXT>
__global__ void __launch_bounds__(kThreads, 4)
pw_fwd_kernel(const float* __restrict__ img, float* __restrict__ out, const float* __restrict__ params,
              const int32_t* __restrict__ ops, const int32_t* __restrict__ seq_len, int N, int S, int flags,
              BankMap bm, PwExt ext) {
    pdl_prologue();
    static_assert(AISP_MAX_STEPS <= kWarps, "one warp per step stages the constants");
    __shared__ float raw[AISP_MAX_STEPS][kConst];
    __shared__ float sc[AISP_MAX_STEPS][kConst];
    __shared__ int sop[AISP_MAX_STEPS];
    __shared__ float bs[EXT ? 3 * (kPwChunkPx / 4) : 1];   // per-thread 4-pixel sums of the chunk (EXT + down only)
    const int b = bank_sample(bm, blockIdx.y);
    const int clip_each = flags & AISP_SEQ_CLIP;
    const bool strict = (flags & AISP_SEQ_STRICT) != 0;
    int chunk = blockIdx.x;
    bool second = false;
    if (EXT && chunk >= ext.chunks0) {   // CTA-uniform: this CTA works on the high-resolution twin
        second = true;
        chunk -= ext.chunks0;
        N = ext.N1;
        img = ext.img1;
        out = ext.out1;
    }
    const bool emit = EXT && VEC == 4 && ext.down != nullptr && !second;
    int len = seq_len ? min(max(seq_len[b], 0), S) : S;
    const int op0 = sample_op(ops, bm, b, S);
    // fills this CTA's chunk of the sample's output with a constant and leaves
    auto fill_chunk = [&](float v) {
        float* q = out + (size_t)b * 3 * (size_t)N;
        const int c0 = chunk * kPwChunkPx;
        for (int pl = 0; pl < 3; ++pl)
            for (int i = c0 + threadIdx.x; i < min(c0 + kPwChunkPx, N); i += kThreads) q[(size_t)pl * N + i] = v;
    };
    if (len > 0 && !is_pointwise(op0)) {
        // another family owns this sample -- except AISP_OP_NONE: the all-zero one-hot row of
        // agent.py:18-23,154 (pdf_sample returned -1), whose gathered image is exactly zero --
        // and except in strict mode, where the caller owns the whole sequence: NaN, never garbage
        if (op0 == AISP_OP_NONE) {
            fill_chunk(0.f);
            if (emit) {   // ... and so are its block means
                const int ow = ext.W / ext.bw, rows = kPwChunkPx / ext.W, oh = (N / ext.W) / ext.bh;
                for (int e = threadIdx.x; e < 3 * (rows / ext.bh) * ow; e += kThreads) {
                    const int ch = e / ((rows / ext.bh) * ow), rem = e - ch * ((rows / ext.bh) * ow);
                    const int oy = (chunk * rows) / ext.bh + rem / ow, ox = rem % ow;
                    if (oy < oh) ext.down[(((size_t)b * 3 + ch) * oh + oy) * ow + ox] = 0.f;
                }
            }
        } else if (strict) {
            fill_chunk(__int_as_float(0x7fc00000));
        }
        return;
    }
    if (EXT && ext.owned) {
        int end;
        if (find_stencil(ops + (size_t)b * S, len, &end) >= 0) return;   // a stencil kernel of the set owns this sample
        len = end;
    }
    stage_consts(params, ops, b, S, len, raw, sc, sop, bm);
    // a stencil op inside a sequence terminates it (documented in the header); strict mode poisons
    for (int k = 0; k < len; ++k)
        if (!is_pointwise(sop[k])) {
            if (strict) { fill_chunk(__int_as_float(0x7fc00000)); return; }
            len = k;
            break;
        }

    constexpr int GROUPS = kPwChunkPx / (kThreads * VEC);
    constexpr int G = (VEC == 4) ? 2 : 8;  // 8 px per thread per round, 4 CTAs/SM: TLP hides the load->compute->store phases
    constexpr int NPX = G * VEC;
    // filter-bank launches (BankMap): outputs and parameters are indexed by the virtual sample b,
    // the image by b / F; the slots of one image are neighbours in the grid and share it through L2
    const size_t base = (size_t)b * 3 * (size_t)N;
    const float* pr = img + (size_t)(b / bm.F) * 3 * (size_t)N;
    float* qr = out + base;
    const int chunk0 = chunk * kPwChunkPx;
    if (VEC == 4) prefetch_chunk_l2(pr, N, chunk0);

    for (int g0 = 0; g0 < GROUPS; g0 += G) {
        float R[NPX], Gc[NPX], Bc[NPX];
        Pack<VEC> t;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int i = chunk0 + ((g0 + j) * kThreads + threadIdx.x) * VEC;
            if (i < N) {
                t.load(pr + i);
#pragma unroll
                for (int v = 0; v < VEC; ++v) R[j * VEC + v] = t.v[v];
                t.load(pr + N + i);
#pragma unroll
                for (int v = 0; v < VEC; ++v) Gc[j * VEC + v] = t.v[v];
                t.load(pr + 2 * (size_t)N + i);
#pragma unroll
                for (int v = 0; v < VEC; ++v) Bc[j * VEC + v] = t.v[v];
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { R[j * VEC + v] = 0.f; Gc[j * VEC + v] = 0.f; Bc[j * VEC + v] = 0.f; }
            }
        }
        for (int k = 0; k < len; ++k) {
            fwd_step<NPX>(sop[k], sc[k], R, Gc, Bc);
            if (clip_each) {
#pragma unroll
                for (int i = 0; i < NPX; ++i) { R[i] = clip01(R[i]); Gc[i] = clip01(Gc[i]); Bc[i] = clip01(Bc[i]); }
            }
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int i = chunk0 + ((g0 + j) * kThreads + threadIdx.x) * VEC;
            if (i < N) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) t.v[v] = R[j * VEC + v];
                t.store(qr + i);
#pragma unroll
                for (int v = 0; v < VEC; ++v) t.v[v] = Gc[j * VEC + v];
                t.store(qr + N + i);
#pragma unroll
                for (int v = 0; v < VEC; ++v) t.v[v] = Bc[j * VEC + v];
                t.store(qr + 2 * (size_t)N + i);
            }
            if (EXT && VEC == 4 && emit) {
                const int q = (g0 + j) * kThreads + threadIdx.x;
                bs[q] = (R[j * 4] + R[j * 4 + 1]) + (R[j * 4 + 2] + R[j * 4 + 3]);
                bs[kPwChunkPx / 4 + q] = (Gc[j * 4] + Gc[j * 4 + 1]) + (Gc[j * 4 + 2] + Gc[j * 4 + 3]);
                bs[2 * (kPwChunkPx / 4) + q] = (Bc[j * 4] + Bc[j * 4 + 1]) + (Bc[j * 4 + 2] + Bc[j * 4 + 3]);
            }
        }
    }
    if (EXT && emit) {
        // (the host asks for this only when the chunk is a whole number of pooling-block rows: W divides
        //  4096, bh divides 4096 / W, bw % 4 == 0 -- see pointwise_can_emit)
        __syncthreads();
        const int W4 = ext.W / 4, rows = kPwChunkPx / ext.W;
        const int ow = ext.W / ext.bw, oh = (N / ext.W) / ext.bh, nby = rows / ext.bh, tpb = ext.bw / 4;
        const float inv = 1.0f / (float)(ext.bh * ext.bw);
        for (int e = threadIdx.x; e < 3 * nby * ow; e += kThreads) {
            const int ch = e / (nby * ow), rem = e - ch * (nby * ow);
            const int pby = rem / ow, pbx = rem - pby * ow;
            const int oy = (chunk * rows) / ext.bh + pby;
            if (oy >= oh) continue;
            float s = 0.f;
            for (int yy = 0; yy < ext.bh; ++yy)
                for (int xx = 0; xx < tpb; ++xx) s += bs[ch * (kPwChunkPx / 4) + (pby * ext.bh + yy) * W4 + pbx * tpb + xx];
            ext.down[(((size_t)b * 3 + ch) * oh + oy) * ow + pbx] = s * inv;
        }
    }
}

// Filter-bank forward: CTA <-> (image, chunk).  The chunk is read ONCE into registers and every
// per-pixel slot of the bank is applied to it in turn (op switch uniform per CTA), each result going
// to its own plane set of the [B,F,3,H,W] stack -- per pixel 12 B read + 12 B written per slot,
// and no reliance on L2 for the re-reads.  Arithmetic per slot is fwd_step, as in pw_fwd_kernel.
template <int VEC>
__global__ void __launch_bounds__(kThreads, 4)
pw_bank_fwd_kernel(const float* __restrict__ img, float* __restrict__ out, const float* __restrict__ params, int N,
                   int clip, BankMap bm) {
    pdl_prologue();
    __shared__ float raw[kMaxBankFilters][kConst];
    __shared__ float sc[kMaxBankFilters][kConst];
    __shared__ int sop[kMaxBankFilters];
    __shared__ int svs[kMaxBankFilters];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = bm.n;
    if (VEC == 4 && threadIdx.x < 3 && (blockIdx.x + 1) * kPwChunkPx <= N) {   // whole chunk -> L2 before the first round
        const float* nx = img + ((size_t)blockIdx.y * 3 + threadIdx.x) * (size_t)N + (size_t)blockIdx.x * kPwChunkPx;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"((unsigned)(kPwChunkPx * sizeof(float))) : "memory");
    }
    for (int j = warp; j < n; j += kWarps) {   // one warp per slot stages its row, lane 0 derives
        const int v = bank_sample(bm, blockIdx.y * n + j);
        raw[j][lane] = (lane < AISP_PSTRIDE) ? params[(size_t)v * AISP_PSTRIDE + lane] : 0.f;
        sc[j][lane] = 0.f;
        __syncwarp();
        if (lane == 0) {
            const int op = bank_op(bm, v);
            sop[j] = op;
            svs[j] = v;
            derive_consts(op, raw[j], sc[j]);
        }
    }
    __syncthreads();

    constexpr int GROUPS = kPwChunkPx / (kThreads * VEC);
    constexpr int NPX = VEC;   // one 128-bit load per plane per round: the slot loop supplies the work per load
    const float* pr = img + (size_t)blockIdx.y * 3 * (size_t)N;
    const int chunk0 = blockIdx.x * kPwChunkPx;
    for (int g0 = 0; g0 < GROUPS; ++g0) {
        const int i = chunk0 + (g0 * kThreads + threadIdx.x) * VEC;
        if (i >= N) break;
        Pack<VEC> xr, xg, xb;
        xr.load(pr + i);
        xg.load(pr + N + i);
        xb.load(pr + 2 * (size_t)N + i);
        for (int j = 0; j < n; ++j) {
            float R[NPX], Gc[NPX], Bc[NPX];
#pragma unroll
            for (int v = 0; v < VEC; ++v) { R[v] = xr.v[v]; Gc[v] = xg.v[v]; Bc[v] = xb.v[v]; }
            fwd_step<NPX>(sop[j], sc[j], R, Gc, Bc);
            Pack<VEC> t;
            float* q = out + (size_t)svs[j] * 3 * (size_t)N + i;
#pragma unroll
            for (int v = 0; v < VEC; ++v) t.v[v] = clip ? clip01(R[v]) : R[v];
            t.store(q);
#pragma unroll
            for (int v = 0; v < VEC; ++v) t.v[v] = clip ? clip01(Gc[v]) : Gc[v];
            t.store(q + N);
#pragma unroll
            for (int v = 0; v < VEC; ++v) t.v[v] = clip ? clip01(Bc[v]) : Bc[v];
            t.store(q + 2 * (size_t)N);
        }
    }
}

template <int OP, int VEC, bool GIMG, bool CLIP>
__device__ __forceinline__ void pw_bwd_body(const float* __restrict__ pr, const float* __restrict__ pg,
                                            float* __restrict__ gi, const float* c, int N, int chunk,
                                            float* red, float* dst, const PooledGrad& pool, int b) {
    constexpr int NACC = PwBwd<OP>::NACC;
    constexpr int GROUPS = kPwChunkPx / (kThreads * VEC);
    constexpr int G = (VEC == 4) ? 1 : 4;  // 4 px per thread per round (6 LDG.128 in flight), 3 CTAs/SM
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
    const int chunk0 = chunk * kPwChunkPx;
    for (int g0 = 0; g0 < GROUPS; g0 += G) {
        Pack<VEC> xr[G], xg[G], xb[G], dr[G], dg[G], db[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int i = chunk0 + ((g0 + j) * kThreads + threadIdx.x) * VEC;
            if (i < N) {
                xr[j].load(pr + i); xg[j].load(pr + N + i); xb[j].load(pr + 2 * (size_t)N + i);
                dr[j].load(pg + i); dg[j].load(pg + N + i); db[j].load(pg + 2 * (size_t)N + i);
                if (pool.g) {   // + the gradient of the pooled image (the VEC pixels share one pooling block)
                    const int y = i >> pool.ws, x = i & ((1 << pool.ws) - 1);
                    const float ar = pooled_at(pool, b, 0, y, x), ag = pooled_at(pool, b, 1, y, x),
                                ab = pooled_at(pool, b, 2, y, x);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) { dr[j].v[v] += ar; dg[j].v[v] += ag; db[j].v[v] += ab; }
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    xr[j].v[v] = xg[j].v[v] = xb[j].v[v] = 0.f;
                    dr[j].v[v] = dg[j].v[v] = db[j].v[v] = 0.f;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                PwBwd<OP>::template px<GIMG, CLIP>(c, xr[j].v[v], xg[j].v[v], xb[j].v[v], dr[j].v[v], dg[j].v[v],
                                                   db[j].v[v], acc);
            if (GIMG) {
                const int i = chunk0 + ((g0 + j) * kThreads + threadIdx.x) * VEC;
                if (i < N) {
                    dr[j].store(gi + i); dg[j].store(gi + N + i); db[j].store(gi + 2 * (size_t)N + i);
                }
            }
        }
    }
    block_reduce_store<NACC>(acc, red, dst);
}

// COLOR == false: every per-pixel filter but the 24-knot ColorFilter, CTA <-> (sample, chunk).
// COLOR == true : ColorFilter samples only -- its 27 partial sums need ~100 registers, which would force
//   spills (a 300-byte stack) onto every other filter's code in a shared 64-register kernel; it gets its
//   own instantiation, launched with ONE CTA per sample that walks the sample's chunks (ColorFilter is
//   outside cfg.filters, so that launch normally consists of B CTAs that exit at once).
template <int VEC, bool GIMG, bool COLOR>
__global__ void __launch_bounds__(kThreads, COLOR ? 1 : 4)
pw_bwd_kernel(const float* __restrict__ img, const float* __restrict__ gout, const float* __restrict__ params,
              const int32_t* __restrict__ ops, int N, int clip, int nchunks, float* __restrict__ gimg,
              float* __restrict__ partial, BankMap bm, PooledGrad pool) {
    pdl_prologue();
    __shared__ float raw[1][kConst];
    __shared__ float sc[1][kConst];
    __shared__ int sop[1];
    __shared__ float red[kWarps * AISP_ACC_STRIDE];
    const int b = bank_sample(bm, blockIdx.y);
    const int op = sample_op(ops, bm, b);
    if (COLOR ? (op != AISP_OP_COLOR) : (!is_pointwise(op) || op == AISP_OP_COLOR)) {
        if (!COLOR && GIMG && op == AISP_OP_NONE) {  // zero image -> zero gradient
            float* q = gimg + (size_t)b * 3 * (size_t)N;
            for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
                const int c0 = chunk * kPwChunkPx;
                for (int pl = 0; pl < 3; ++pl)
                    for (int i = c0 + threadIdx.x; i < min(c0 + kPwChunkPx, N); i += kThreads) q[(size_t)pl * N + i] = 0.f;
            }
        }
        return;
    }
    stage_consts(params, ops, b, 1, 1, raw, sc, sop, bm);
    const size_t base = (size_t)b * 3 * (size_t)N;
    const float* pr = img + (size_t)(b / bm.F) * 3 * (size_t)N;
    const float* pg = gout + base;
    float* gi = GIMG ? gimg + base : nullptr;
    const float* c = sc[0];
    // (no chunk prefetch here: with only four dependent-free rounds per CTA the demand loads are already
    //  all in flight; measured +7 us per launch with it)
    for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        float* dst = partial + ((size_t)b * nchunks + chunk) * AISP_ACC_STRIDE;
        if (chunk != (int)blockIdx.x) __syncthreads();   // `red` is reused chunk after chunk
        // the clip flag is CTA-uniform: one branch here instead of a predicate on every pixel
#define AISP_CASE(OPC)                                                                     \
    case OPC:                                                                              \
        if (clip) pw_bwd_body<OPC, VEC, GIMG, true>(pr, pg, gi, c, N, chunk, red, dst, pool, b);    \
        else pw_bwd_body<OPC, VEC, GIMG, false>(pr, pg, gi, c, N, chunk, red, dst, pool, b);        \
        break;
        if (COLOR) {
            switch (op) {
                AISP_CASE(AISP_OP_COLOR)
            default: break;
            }
        } else {
            switch (op) {
                AISP_CASE(AISP_OP_EXPOSURE)
                AISP_CASE(AISP_OP_GAMMA)
                AISP_CASE(AISP_OP_WB)
                AISP_CASE(AISP_OP_CCM)
                AISP_CASE(AISP_OP_TONE)
                AISP_CASE(AISP_OP_CONTRAST)
                AISP_CASE(AISP_OP_WNB)
                AISP_CASE(AISP_OP_SATPLUS)
            default: break;
            }
        }
#undef AISP_CASE
    }
}

// Filter-bank backward (parameter gradients): CTA <-> (image, chunk).  Each thread parks its own
// pixels of the chunk in shared memory once (private slots: no barrier, conflict-free 128-bit
// accesses) and sweeps the bank's per-pixel slots over them; only the upstream gradient of each
// slot is streamed from HBM.  Per-thread pixel assignment, accumulation order and the block
// reduction are those of pw_bwd_body, so the partial sums are bit-identical to F separate launches.
template <int OP, int VEC, bool CLIP>
__device__ __forceinline__ void pw_bank_bwd_body(const float* __restrict__ sx, const float* __restrict__ pg,
                                                 const float* c, int N, float* red, float* dst) {
    constexpr int NACC = PwBwd<OP>::NACC;
    constexpr int GROUPS = kPwChunkPx / (kThreads * VEC);
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
    const int chunk0 = blockIdx.x * kPwChunkPx;
    for (int g0 = 0; g0 < GROUPS; ++g0) {
        const int i = chunk0 + (g0 * kThreads + threadIdx.x) * VEC;
        Pack<VEC> dr, dg, db;
        float xr[VEC], xg[VEC], xb[VEC];
        if (i < N) {
            dr.load(pg + i); dg.load(pg + N + i); db.load(pg + 2 * (size_t)N + i);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) dr.v[v] = dg.v[v] = db.v[v] = 0.f;
        }
        const float* px = sx + ((size_t)(g0 * 3) * kThreads + threadIdx.x) * VEC;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            xr[v] = px[v];
            xg[v] = px[kThreads * VEC + v];
            xb[v] = px[2 * kThreads * VEC + v];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            PwBwd<OP>::template px<false, CLIP>(c, xr[v], xg[v], xb[v], dr.v[v], dg.v[v], db.v[v], acc);
    }
    block_reduce_store<NACC>(acc, red, dst);
}

// (COLOR: the bank's ColorFilter slots run in their own instantiation, see pw_bwd_kernel)
template <int VEC, bool COLOR>
__global__ void __launch_bounds__(kThreads, COLOR ? 1 : 4)
pw_bank_bwd_kernel(const float* __restrict__ img, const float* __restrict__ gout, const float* __restrict__ params,
                   int N, int clip, float* __restrict__ partial, BankMap bm) {
    pdl_prologue();
    extern __shared__ float4 sx4[];     // [GROUPS][3][kThreads] packs of VEC floats = 3 * kPwChunkPx floats
    float* sx = reinterpret_cast<float*>(sx4);
    __shared__ float raw[kMaxBankFilters][kConst];
    __shared__ float sc[kMaxBankFilters][kConst];
    __shared__ int sop[kMaxBankFilters];
    __shared__ int svs[kMaxBankFilters];
    __shared__ float red[2][kWarps * AISP_ACC_STRIDE];   // alternating per slot: one barrier per slot suffices
    constexpr int GROUPS = kPwChunkPx / (kThreads * VEC);
    const int n = bm.n;
    if (VEC == 4 && threadIdx.x < 3 && (blockIdx.x + 1) * kPwChunkPx <= N) {   // first slot's gradient chunk -> L2
        const float* nx = gout + (size_t)bank_sample(bm, blockIdx.y * n) * 3 * (size_t)N + (size_t)threadIdx.x * N +
                          (size_t)blockIdx.x * kPwChunkPx;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"((unsigned)(kPwChunkPx * sizeof(float))) : "memory");
    }
    {   // every slot's constants up front, one warp per slot (as in pw_bank_fwd_kernel)
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int j = warp; j < n; j += kWarps) {
            const int v = bank_sample(bm, blockIdx.y * n + j);
            raw[j][lane] = (lane < AISP_PSTRIDE) ? params[(size_t)v * AISP_PSTRIDE + lane] : 0.f;
            sc[j][lane] = 0.f;
            __syncwarp();
            if (lane == 0) {
                const int op = bank_op(bm, v);
                sop[j] = op;
                svs[j] = v;
                derive_consts(op, raw[j], sc[j]);
            }
        }
    }
    const float* pr = img + (size_t)blockIdx.y * 3 * (size_t)N;
    const int chunk0 = blockIdx.x * kPwChunkPx;
#pragma unroll
    for (int g0 = 0; g0 < GROUPS; ++g0) {
        const int i = chunk0 + (g0 * kThreads + threadIdx.x) * VEC;
        Pack<VEC> t[3];
        if (i < N) {
            t[0].load(pr + i); t[1].load(pr + N + i); t[2].load(pr + 2 * (size_t)N + i);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) t[0].v[v] = t[1].v[v] = t[2].v[v] = 0.f;
        }
        float* px = sx + ((size_t)(g0 * 3) * kThreads + threadIdx.x) * VEC;
#pragma unroll
        for (int pl = 0; pl < 3; ++pl)
#pragma unroll
            for (int v = 0; v < VEC; ++v) px[pl * kThreads * VEC + v] = t[pl].v[v];
    }
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        const int b = svs[j];
        const float* pg = gout + (size_t)b * 3 * (size_t)N;
        // pull the next slot's upstream-gradient chunk into L2 while this slot computes (bulk prefetch:
        // no registers, no shared memory; the demand loads of the next slot then see L2 latency)
        if (VEC == 4 && j + 1 < n && threadIdx.x < 3 && chunk0 + kPwChunkPx <= N) {
            const float* nx = gout + (size_t)svs[j + 1] * 3 * (size_t)N + (size_t)threadIdx.x * N + chunk0;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"((unsigned)(kPwChunkPx * sizeof(float))) : "memory");
        }
        float* dst = partial + ((size_t)b * gridDim.x + blockIdx.x) * AISP_ACC_STRIDE;
        const float* c = sc[j];
        float* redj = red[j & 1];
#define AISP_CASE(OPC)                                                          \
    case OPC:                                                                   \
        if (clip) pw_bank_bwd_body<OPC, VEC, true>(sx, pg, c, N, redj, dst);    \
        else pw_bank_bwd_body<OPC, VEC, false>(sx, pg, c, N, redj, dst);        \
        break;
        if (COLOR) {
            switch (sop[j]) {
                AISP_CASE(AISP_OP_COLOR)
            default: break;
            }
        } else {
            switch (sop[j]) {
                AISP_CASE(AISP_OP_EXPOSURE)
                AISP_CASE(AISP_OP_GAMMA)
                AISP_CASE(AISP_OP_WB)
                AISP_CASE(AISP_OP_CCM)
                AISP_CASE(AISP_OP_TONE)
                AISP_CASE(AISP_OP_CONTRAST)
                AISP_CASE(AISP_OP_WNB)
                AISP_CASE(AISP_OP_SATPLUS)
            default: break;
            }
        }
#undef AISP_CASE
    }
}

__global__ void __launch_bounds__(kThreads)
finalize_kernel(const float* __restrict__ partial, int nrows, const float* __restrict__ params,
                const int32_t* __restrict__ ops, int family, float* __restrict__ grad_params, BankMap bm) {
    pdl_prologue();
    __shared__ double part[kWarps][AISP_ACC_STRIDE];
    __shared__ double tot[AISP_ACC_STRIDE];
    __shared__ float raw[kConst];
    __shared__ float c[kConst];
    const int b = bank_sample(bm, blockIdx.x);
    const int op = sample_op(ops, bm, b);
    if (!in_family(op, family)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* rows = partial + (size_t)b * nrows * AISP_ACC_STRIDE;
    double s = 0.0;
    for (int r = warp; r < nrows; r += kWarps) s += (double)rows[(size_t)r * AISP_ACC_STRIDE + lane];
    part[warp][lane] = s;
    if (warp == 0) {
        raw[lane] = (lane < AISP_PSTRIDE) ? params[(size_t)b * AISP_PSTRIDE + lane] : 0.f;
        c[lane] = 0.f;
    }
    __syncthreads();
    if (warp == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += part[w][lane];
        tot[lane] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (family == FAMILY_SHARPEN) {   // a sharpen scratch row holds (sum0, sum1, 0, 0) per warp: see sharpen_kernel
            double f0 = 0.0, f1 = 0.0;
            for (int w = 0; w < kWarps; ++w) { f0 += tot[4 * w]; f1 += tot[4 * w + 1]; }
            tot[0] = f0;
            tot[1] = f1;
        }
        derive_consts(op, raw, c);
        float gp[AISP_PSTRIDE];
        finalize_grads(op, tot, c, raw, gp);
        for (int k = 0; k < AISP_PSTRIDE; ++k) grad_params[(size_t)b * AISP_PSTRIDE + k] = gp[k];
    }
}

// =============================================================================================
// host-side launchers (called from capi.cu)
// =============================================================================================
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

cudaError_t launch_pointwise_fwd(const float* img, float* out, const float* params, const int32_t* ops,
                                 const int32_t* seq_len, int B, int H, int W, int S, int clip_each, BankMap bm,
                                 cudaStream_t st) {
    const long long N = (long long)H * W;
    dim3 grid((unsigned)((N + kPwChunkPx - 1) / kPwChunkPx), (unsigned)B);
    const bool vec = (N % 4 == 0) && aligned16(img) && aligned16(out);
    const PwExt none{};
    if (vec)
        launch_pdl(pw_fwd_kernel<4, false>, grid, kThreads, st, img, out, params, ops, seq_len, (int)N, S, clip_each, bm, none);
    else
        launch_pdl(pw_fwd_kernel<1, false>, grid, kThreads, st, img, out, params, ops, seq_len, (int)N, S, clip_each, bm, none);
    return cudaGetLastError();
}

// Can the per-pixel kernel emit the [oh, ow] block means of an H x W image from its store path?  A CTA's
// 4096-pixel chunk must be a whole number of pooling-block rows and a thread's 4 pixels lie in one block.
bool pointwise_can_emit(int H, int W, int oh, int ow) {
    if (oh <= 0 || ow <= 0 || H % oh || W % ow || (W & 3)) return false;
    const int bh = H / oh, bw = W / ow;
    return (bw % 4 == 0) && (kPwChunkPx % W == 0) && ((kPwChunkPx / W) % bh == 0);
}

// Per-pixel member of a sequence launch set: samples whose sequence holds a stencil step are left to the
// stencil kernels; optional high-resolution twin in the same launch; optional block means of the output.
cudaError_t launch_pointwise_seq_fwd(const float* img, float* out, const float* params, const int32_t* ops,
                                     const int32_t* seq_len, int B, int H, int W, int S, int flags, const float* hr_img,
                                     float* hr_out, int hr_H, int hr_W, float* down, int oh, int ow, cudaStream_t st) {
    const long long N = (long long)H * W, N1 = hr_img ? (long long)hr_H * hr_W : 0;
    PwExt ext{};
    ext.owned = 1;
    ext.chunks0 = (int)((N + kPwChunkPx - 1) / kPwChunkPx);
    const bool vec0 = (N % 4 == 0) && aligned16(img) && aligned16(out);
    const bool vec1 = !hr_img || ((N1 % 4 == 0) && aligned16(hr_img) && aligned16(hr_out));
    // the twin shares the launch when both image sets take the same (vector / scalar) instantiation
    const bool together = hr_img && (vec0 == vec1);
    if (down) { ext.down = down; ext.W = W; ext.bh = H / oh; ext.bw = W / ow; }
    int chunks = ext.chunks0;
    if (together) {
        ext.img1 = hr_img; ext.out1 = hr_out; ext.N1 = (int)N1;
        chunks += (int)((N1 + kPwChunkPx - 1) / kPwChunkPx);
    }
    dim3 grid((unsigned)chunks, (unsigned)B);
    if (vec0)
        launch_pdl(pw_fwd_kernel<4, true>, grid, kThreads, st, img, out, params, ops, seq_len, (int)N, S, flags, plain_batch(), ext);
    else
        launch_pdl(pw_fwd_kernel<1, true>, grid, kThreads, st, img, out, params, ops, seq_len, (int)N, S, flags, plain_batch(), ext);
    if (hr_img && !together) {
        PwExt e2{};
        e2.owned = 1;
        e2.chunks0 = (int)((N1 + kPwChunkPx - 1) / kPwChunkPx);
        dim3 g2((unsigned)e2.chunks0, (unsigned)B);
        if (vec1)
            launch_pdl(pw_fwd_kernel<4, true>, g2, kThreads, st, hr_img, hr_out, params, ops, seq_len, (int)N1, S, flags, plain_batch(), e2);
        else
            launch_pdl(pw_fwd_kernel<1, true>, g2, kThreads, st, hr_img, hr_out, params, ops, seq_len, (int)N1, S, flags, plain_batch(), e2);
    }
    return cudaGetLastError();
}

// bank forward over the per-pixel slots of `bm` (bm.n >= 1): grid (chunks, images)
cudaError_t launch_pointwise_bank_fwd(const float* img, float* out, const float* params, int B, int H, int W, int clip,
                                      BankMap bm, cudaStream_t st) {
    const long long N = (long long)H * W;
    dim3 grid((unsigned)((N + kPwChunkPx - 1) / kPwChunkPx), (unsigned)B);
    if ((N % 4 == 0) && aligned16(img) && aligned16(out))
        launch_pdl(pw_bank_fwd_kernel<4>, grid, kThreads, st, img, out, params, (int)N, clip, bm);
    else
        launch_pdl(pw_bank_fwd_kernel<1>, grid, kThreads, st, img, out, params, (int)N, clip, bm);
    return cudaGetLastError();
}

int pointwise_rows(int H, int W) { return (int)(((long long)H * W + kPwChunkPx - 1) / kPwChunkPx); }

cudaError_t launch_finalize(const float* partial, int nrows, const float* params, const int32_t* ops, int family,
                            int B, float* grad_params, BankMap bm, cudaStream_t st) {
    launch_pdl(finalize_kernel, B, kThreads, st, partial, nrows, params, ops, family, grad_params, bm);
    return cudaGetLastError();
}

cudaError_t launch_pointwise_bwd(const float* img, const float* gout, const float* params, const int32_t* ops,
                                 int B, int H, int W, int clip, float* grad_params, float* grad_img,
                                 float* partial, BankMap bm, PooledGrad pool, cudaStream_t st) {
    const long long N = (long long)H * W;
    const int rows = pointwise_rows(H, W);
    // main launch: CTA <-> (sample, chunk); ColorFilter launch: one CTA per sample (see pw_bwd_kernel)
    const dim3 grid((unsigned)rows, (unsigned)B), grid_color(1, (unsigned)B);
    const bool vec = (N % 4 == 0) && aligned16(img) && aligned16(gout) && (!grad_img || aligned16(grad_img));
#define AISP_LAUNCH(VEC, GIMG)                                                                                        \
    do {                                                                                                              \
        launch_pdl(pw_bwd_kernel<VEC, GIMG, false>, grid, kThreads, st, img, gout, params, ops, (int)N, clip, rows,  \
                   grad_img, partial, bm, pool);                                                                      \
        launch_pdl(pw_bwd_kernel<VEC, GIMG, true>, grid_color, kThreads, st, img, gout, params, ops, (int)N, clip,   \
                   rows, grad_img, partial, bm, pool);                                                                \
    } while (0)
    if (vec) {
        if (grad_img) AISP_LAUNCH(4, true);
        else AISP_LAUNCH(4, false);
    } else {
        if (grad_img) AISP_LAUNCH(1, true);
        else AISP_LAUNCH(1, false);
    }
#undef AISP_LAUNCH
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_finalize(partial, rows, params, ops, FAMILY_POINTWISE, B, grad_params, bm, st);
}

// bank backward over the per-pixel slots of `bm` (bm.n >= 1): grid (chunks, images) + finalize.
// `color`: the slots are ColorFilter slots (their own instantiation, see pw_bwd_kernel).
template <int VEC, bool COLOR>
static void bank_bwd_attrs(size_t smem) {
    cudaFuncSetAttribute(pw_bank_bwd_kernel<VEC, COLOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(pw_bank_bwd_kernel<VEC, COLOR>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
}

cudaError_t launch_pointwise_bank_bwd(const float* img, const float* gout, const float* params, int B, int H, int W,
                                      int clip, float* grad_params, float* partial, BankMap bm, bool color,
                                      cudaStream_t st) {
    const long long N = (long long)H * W;
    const int rows = pointwise_rows(H, W);
    dim3 grid((unsigned)rows, (unsigned)B);
    constexpr size_t smem = 3 * kPwChunkPx * sizeof(float);
    static bool attr_set_on[64] = {};   // per device: function attributes belong to the device's context
    int devi = 0;
    cudaGetDevice(&devi);
    bool& attr_set = attr_set_on[devi & 63];
    if (!attr_set) {   // > 48 KB of dynamic shared memory is opt-in (per function, idempotent)
        bank_bwd_attrs<4, false>(smem);
        bank_bwd_attrs<1, false>(smem);
        bank_bwd_attrs<4, true>(smem);
        bank_bwd_attrs<1, true>(smem);
        attr_set = true;
    }
    const bool vec = (N % 4 == 0) && aligned16(img) && aligned16(gout);
    if (vec && !color)
        launch_pdl_smem(pw_bank_bwd_kernel<4, false>, grid, kThreads, smem, st, img, gout, params, (int)N, clip, partial, bm);
    else if (vec)
        launch_pdl_smem(pw_bank_bwd_kernel<4, true>, grid, kThreads, smem, st, img, gout, params, (int)N, clip, partial, bm);
    else if (!color)
        launch_pdl_smem(pw_bank_bwd_kernel<1, false>, grid, kThreads, smem, st, img, gout, params, (int)N, clip, partial, bm);
    else
        launch_pdl_smem(pw_bank_bwd_kernel<1, true>, grid, kThreads, smem, st, img, gout, params, (int)N, clip, partial, bm);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_finalize(partial, rows, params, nullptr, FAMILY_POINTWISE, B * bm.n, grad_params, bm, st);
}

}  // namespace aisp
