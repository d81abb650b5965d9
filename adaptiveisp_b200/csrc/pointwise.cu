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
template <int VEC>
__global__ void __launch_bounds__(kThreads, 4)
pw_fwd_kernel(const float* __restrict__ img, float* __restrict__ out, const float* __restrict__ params,
              const int32_t* __restrict__ ops, const int32_t* __restrict__ seq_len, int N, int S, int flags,
              BankMap bm) {
    pdl_prologue();
    static_assert(AISP_MAX_STEPS <= kWarps, "one warp per step stages the constants");
    __shared__ float raw[AISP_MAX_STEPS][kConst];
    __shared__ float sc[AISP_MAX_STEPS][kConst];
    __shared__ int sop[AISP_MAX_STEPS];
    const int b = bank_sample(bm, blockIdx.y);
    const int clip_each = flags & AISP_SEQ_CLIP;
    const bool strict = (flags & AISP_SEQ_STRICT) != 0;
    int len = seq_len ? min(max(seq_len[b], 0), S) : S;
    const int op0 = sample_op(ops, bm, b, S);
    // fills this CTA's chunk of the sample's output with a constant and leaves
    auto fill_chunk = [&](float v) {
        float* q = out + (size_t)b * 3 * (size_t)N;
        const int c0 = blockIdx.x * kPwChunkPx;
        for (int pl = 0; pl < 3; ++pl)
            for (int i = c0 + threadIdx.x; i < min(c0 + kPwChunkPx, N); i += kThreads) q[(size_t)pl * N + i] = v;
    };
    if (len > 0 && !is_pointwise(op0)) {
        // another family owns this sample -- except AISP_OP_NONE: the all-zero one-hot row of
        // agent.py:18-23,154 (pdf_sample returned -1), whose gathered image is exactly zero --
        // and except in strict mode, where the caller owns the whole sequence: NaN, never garbage
        if (op0 == AISP_OP_NONE) fill_chunk(0.f);
        else if (strict) fill_chunk(__int_as_float(0x7fc00000));
        return;
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
    const int chunk0 = blockIdx.x * kPwChunkPx;
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

template <int OP, int VEC, bool GIMG>
__device__ __forceinline__ void pw_bwd_body(const float* __restrict__ pr, const float* __restrict__ pg,
                                            float* __restrict__ gi, const float* c, int N, int clip,
                                            float* red, float* dst) {
    constexpr int NACC = PwBwd<OP>::NACC;
    constexpr int GROUPS = kPwChunkPx / (kThreads * VEC);
    constexpr int G = (VEC == 4) ? 1 : 4;  // 4 px per thread per round (6 LDG.128 in flight), 3 CTAs/SM
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
    const int chunk0 = blockIdx.x * kPwChunkPx;
    for (int g0 = 0; g0 < GROUPS; g0 += G) {
        Pack<VEC> xr[G], xg[G], xb[G], dr[G], dg[G], db[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int i = chunk0 + ((g0 + j) * kThreads + threadIdx.x) * VEC;
            if (i < N) {
                xr[j].load(pr + i); xg[j].load(pr + N + i); xb[j].load(pr + 2 * (size_t)N + i);
                dr[j].load(pg + i); dg[j].load(pg + N + i); db[j].load(pg + 2 * (size_t)N + i);
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
                PwBwd<OP>::template px<GIMG>(c, xr[j].v[v], xg[j].v[v], xb[j].v[v], dr[j].v[v], dg[j].v[v],
                                             db[j].v[v], clip, acc);
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

template <int VEC, bool GIMG>
__global__ void __launch_bounds__(kThreads, 4)
pw_bwd_kernel(const float* __restrict__ img, const float* __restrict__ gout, const float* __restrict__ params,
              const int32_t* __restrict__ ops, int N, int clip, float* __restrict__ gimg,
              float* __restrict__ partial, BankMap bm) {
    pdl_prologue();
    __shared__ float raw[1][kConst];
    __shared__ float sc[1][kConst];
    __shared__ int sop[1];
    __shared__ float red[kWarps * AISP_ACC_STRIDE];
    const int b = bank_sample(bm, blockIdx.y);
    const int op = sample_op(ops, bm, b);
    if (!is_pointwise(op)) {
        if (GIMG && op == AISP_OP_NONE) {  // zero image -> zero gradient
            float* q = gimg + (size_t)b * 3 * (size_t)N;
            const int c0 = blockIdx.x * kPwChunkPx;
            for (int pl = 0; pl < 3; ++pl)
                for (int i = c0 + threadIdx.x; i < min(c0 + kPwChunkPx, N); i += kThreads) q[(size_t)pl * N + i] = 0.f;
        }
        return;
    }
    stage_consts(params, ops, b, 1, 1, raw, sc, sop, bm);
    const size_t base = (size_t)b * 3 * (size_t)N;
    const float* pr = img + (size_t)(b / bm.F) * 3 * (size_t)N;
    const float* pg = gout + base;
    float* gi = GIMG ? gimg + base : nullptr;
    float* dst = partial + ((size_t)b * gridDim.x + blockIdx.x) * AISP_ACC_STRIDE;
    const float* c = sc[0];
    // (no chunk prefetch here: with only four dependent-free rounds per CTA the demand loads are already
    //  all in flight; measured +7 us per launch with it)
    switch (op) {
#define AISP_CASE(OPC) \
    case OPC: pw_bwd_body<OPC, VEC, GIMG>(pr, pg, gi, c, N, clip, red, dst); break;
        AISP_CASE(AISP_OP_EXPOSURE)
        AISP_CASE(AISP_OP_GAMMA)
        AISP_CASE(AISP_OP_WB)
        AISP_CASE(AISP_OP_CCM)
        AISP_CASE(AISP_OP_TONE)
        AISP_CASE(AISP_OP_COLOR)
        AISP_CASE(AISP_OP_CONTRAST)
        AISP_CASE(AISP_OP_WNB)
        AISP_CASE(AISP_OP_SATPLUS)
#undef AISP_CASE
    default: break;
    }
}

// Filter-bank backward (parameter gradients): CTA <-> (image, chunk).  Each thread parks its own
// pixels of the chunk in shared memory once (private slots: no barrier, conflict-free 128-bit
// accesses) and sweeps the bank's per-pixel slots over them; only the upstream gradient of each
// slot is streamed from HBM.  Per-thread pixel assignment, accumulation order and the block
// reduction are those of pw_bwd_body, so the partial sums are bit-identical to F separate launches.
template <int OP, int VEC>
__device__ __forceinline__ void pw_bank_bwd_body(const float* __restrict__ sx, const float* __restrict__ pg,
                                                 const float* c, int N, int clip, float* red, float* dst) {
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
            PwBwd<OP>::template px<false>(c, xr[v], xg[v], xb[v], dr.v[v], dg.v[v], db.v[v], clip, acc);
    }
    block_reduce_store<NACC>(acc, red, dst);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads, 4)
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
        switch (sop[j]) {
#define AISP_CASE(OPC) \
    case OPC: pw_bank_bwd_body<OPC, VEC>(sx, pg, c, N, clip, redj, dst); break;
            AISP_CASE(AISP_OP_EXPOSURE)
            AISP_CASE(AISP_OP_GAMMA)
            AISP_CASE(AISP_OP_WB)
            AISP_CASE(AISP_OP_CCM)
            AISP_CASE(AISP_OP_TONE)
            AISP_CASE(AISP_OP_COLOR)
            AISP_CASE(AISP_OP_CONTRAST)
            AISP_CASE(AISP_OP_WNB)
            AISP_CASE(AISP_OP_SATPLUS)
#undef AISP_CASE
        default: break;
        }
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
    if (vec)
        launch_pdl(pw_fwd_kernel<4>, grid, kThreads, st, img, out, params, ops, seq_len, (int)N, S, clip_each, bm);
    else
        launch_pdl(pw_fwd_kernel<1>, grid, kThreads, st, img, out, params, ops, seq_len, (int)N, S, clip_each, bm);
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
                                 float* partial, BankMap bm, cudaStream_t st) {
    const long long N = (long long)H * W;
    const int rows = pointwise_rows(H, W);
    dim3 grid((unsigned)rows, (unsigned)B);
    const bool vec = (N % 4 == 0) && aligned16(img) && aligned16(gout) && (!grad_img || aligned16(grad_img));
    if (vec) {
        if (grad_img)
            launch_pdl(pw_bwd_kernel<4, true>, grid, kThreads, st, img, gout, params, ops, (int)N, clip, grad_img, partial, bm);
        else
            launch_pdl(pw_bwd_kernel<4, false>, grid, kThreads, st, img, gout, params, ops, (int)N, clip, nullptr, partial, bm);
    } else {
        if (grad_img)
            launch_pdl(pw_bwd_kernel<1, true>, grid, kThreads, st, img, gout, params, ops, (int)N, clip, grad_img, partial, bm);
        else
            launch_pdl(pw_bwd_kernel<1, false>, grid, kThreads, st, img, gout, params, ops, (int)N, clip, nullptr, partial, bm);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_finalize(partial, rows, params, ops, FAMILY_POINTWISE, B, grad_params, bm, st);
}

// bank backward over the per-pixel slots of `bm` (bm.n >= 1): grid (chunks, images) + finalize
cudaError_t launch_pointwise_bank_bwd(const float* img, const float* gout, const float* params, int B, int H, int W,
                                      int clip, float* grad_params, float* partial, BankMap bm, cudaStream_t st) {
    const long long N = (long long)H * W;
    const int rows = pointwise_rows(H, W);
    dim3 grid((unsigned)rows, (unsigned)B);
    constexpr size_t smem = 3 * kPwChunkPx * sizeof(float);
    static bool attr_set_on[64] = {};   // per device: function attributes belong to the device's context
    int devi = 0;
    cudaGetDevice(&devi);
    bool& attr_set = attr_set_on[devi & 63];
    if (!attr_set) {   // > 48 KB of dynamic shared memory is opt-in (per function, idempotent)
        cudaFuncSetAttribute(pw_bank_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(pw_bank_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(pw_bank_bwd_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(pw_bank_bwd_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    if ((N % 4 == 0) && aligned16(img) && aligned16(gout))
        launch_pdl_smem(pw_bank_bwd_kernel<4>, grid, kThreads, smem, st, img, gout, params, (int)N, clip, partial, bm);
    else
        launch_pdl_smem(pw_bank_bwd_kernel<1>, grid, kThreads, smem, st, img, gout, params, (int)N, clip, partial, bm);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_finalize(partial, rows, params, nullptr, FAMILY_POINTWISE, B * bm.n, grad_params, bm, st);
}

}  // namespace aisp
