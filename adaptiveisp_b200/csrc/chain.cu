// Fused multi-step backward for sm_100a: a per-sample SEQUENCE of up to kChainMax per-pixel filters is
// differentiated in ONE pass over HBM (image 12 + upstream gradient 12 [+ image gradient 12] bytes per
// pixel for the whole sequence -- the traffic of a single-step backward).
//
// CTA <-> (sample, chunk of rounds x 1024 pixels), 256 threads, 4 pixels per thread and round.
//   * The image and upstream-gradient vectors of round r+1 are copied global -> shared with cp.async
//     (LDGSTS, 16 B per copy, L1 bypassed) into a two-slot ring while round r computes.  Every thread
//     copies exactly the six vectors it will read itself, so the ring needs no barrier at all -- only
//     cp.async.wait_group -- and no register is spent on data in flight.
//   * Forward sweep: the chain is recomputed stage by stage; the input of every stage after the first
//     is parked in shared memory (thread-private slots, conflict-free 128-bit accesses) instead of in
//     registers: the register file then only holds one stage's pixels, the running gradient and the
//     per-stage partial sums.
//   * Reverse sweep: each stage turns the running gradient into its parameter-gradient partial sums
//     (the single-step bodies PwBwd<OP>, bit-identical arithmetic) and into the gradient w.r.t. its
//     input.  The op `switch` is CTA-uniform.
//   * One scratch row per (sample, chunk, stage); only the accumulators the stage's op really has are
//     reduced (1 for E/G/Ct/S+/BW, 3 for W, 9 for CCM/T, 27 for C).
// Sequences that contain the 24-knot ColorFilter (27 partial sums per stage) are served by a second
// instantiation with 27-wide accumulators launched with ONE CTA per sample (it walks the sample's
// chunks); the main instantiation skips those samples and vice versa, so ColorFilter costs the common
// sequences nothing.
#include "pointwise_math.cuh"

namespace aisp {

constexpr int kChainMax = AISP_MAX_CHAIN_BWD;   // longest sequence differentiated in one pass
constexpr int kChainPx = 4;                     // pixels per thread per round
constexpr int kRoundPx = kThreads * kChainPx;   // 1024 pixels per CTA round

__host__ __device__ __forceinline__ int op_nacc(int op) {
    switch (op) {
    case AISP_OP_WB: return 3;
    case AISP_OP_CCM:
    case AISP_OP_TONE: return 9;
    case AISP_OP_COLOR: return 27;
    default: return 1;
    }
}

template <int NPX, bool GX, bool WITH_COLOR, bool CLIP>
__device__ __forceinline__ void bwd_step(int op, const float* __restrict__ c, const float (&R)[NPX],
                                         const float (&G)[NPX], const float (&B)[NPX], float (&gr)[NPX],
                                         float (&gg)[NPX], float (&gb)[NPX], float* acc) {
    switch (op) {
#define AISP_CHAIN_CASE(OPC)                                                                         \
    case OPC: {                                                                                      \
        _Pragma("unroll") for (int i = 0; i < NPX; ++i)                                              \
            PwBwd<OPC>::template px<GX, CLIP>(c, R[i], G[i], B[i], gr[i], gg[i], gb[i], acc);        \
        break;                                                                                       \
    }
        AISP_CHAIN_CASE(AISP_OP_EXPOSURE)
        AISP_CHAIN_CASE(AISP_OP_GAMMA)
        AISP_CHAIN_CASE(AISP_OP_WB)
        AISP_CHAIN_CASE(AISP_OP_CCM)
        AISP_CHAIN_CASE(AISP_OP_TONE)
        AISP_CHAIN_CASE(AISP_OP_CONTRAST)
        AISP_CHAIN_CASE(AISP_OP_WNB)
        AISP_CHAIN_CASE(AISP_OP_SATPLUS)
    case AISP_OP_COLOR:
        if (WITH_COLOR) {
#pragma unroll
            for (int i = 0; i < NPX; ++i)
                PwBwd<AISP_OP_COLOR>::template px<GX, CLIP>(c, R[i], G[i], B[i], gr[i], gg[i], gb[i], acc);
        }
        break;
#undef AISP_CHAIN_CASE
    default: break;
    }
}

// first `n` of NMAX per-thread partial sums -> one scratch row (all AISP_ACC_STRIDE lanes written)
template <int NMAX>
__device__ __forceinline__ void block_reduce_store_n(const float (&acc)[NMAX], int n, float* red, float* dst) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
        if (k < n) {   // CTA-uniform
            float v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp * AISP_ACC_STRIDE + k] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x < AISP_ACC_STRIDE) {
        float s = 0.f;
        if ((int)threadIdx.x < n) {
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += red[w * AISP_ACC_STRIDE + threadIdx.x];
        }
        dst[threadIdx.x] = s;
    }
}

// What a CTA does with sample b: returns the effective sequence length (>= 1) when this instantiation
// owns the sample, 0 when it has nothing to compute.  `fill` tells the caller what to put into the
// image gradient in that case: 0 = leave untouched (another kernel owns the sample), 1 = zeros
// (AISP_OP_NONE: zero image, zero gradient), 2 = NaN (strict mode: a non-per-pixel op in the sequence),
// 3 = a copy of the upstream gradient (empty sequence: the forward was the identity).
__device__ __forceinline__ int classify_sequence(const int32_t* __restrict__ ops, const int32_t* __restrict__ seq_len,
                                                 int b, int S, bool strict, bool with_color, int* fill) {
    int len = seq_len ? min(max(seq_len[b], 0), S) : S;
    *fill = 0;
    if (len == 0) { *fill = 3; return 0; }     // empty sequence: identity, the gradient passes through
    const int32_t* o = ops + (size_t)b * S;
    if (o[0] == AISP_OP_NONE) { *fill = 1; return 0; }
    bool color = false;
    for (int k = 0; k < len; ++k) {
        const int op = o[k];
        if (!is_pointwise(op)) {
            if (strict) { *fill = 2; return 0; }
            len = k;                           // a stencil op ends the sequence (select-apply semantics)
            break;
        }
        color |= (op == AISP_OP_COLOR);
    }
    if (len == 0) return 0;                    // led by a stencil op: another family owns the sample
    if (color != with_color) return 0;         // the other instantiation owns the sample
    return len;
}

// ---------------------------------------------------------------------------------------------
// Compile-time sequences.  A sequence that is known when the library is built gets its own kernel: the op
// dispatch, the shared-memory parking and the padded accumulator rows of the generic kernel disappear
// (every stage input lives in registers, a stage owns exactly the partial sums its op has), and the
// compiler schedules forward recomputation and reverse sweep as one straight-line block.  Registered:
// the per-pixel prefix of the reference's fixed pipeline (BASELINE configs[0], isp/filters.py:753-815):
// exposure -> gamma -> white balance -> CCM.  A sample whose sequence is exactly a registered one is served
// by that kernel and skipped by the generic one; both are always launched (the ops live on the device).
// ---------------------------------------------------------------------------------------------
template <int... OPS>
struct OpSeq {
    static constexpr int N = sizeof...(OPS);
    __host__ __device__ static constexpr int at(int k) {
        constexpr int v[N] = {OPS...};
        return v[k];
    }
    static __device__ __forceinline__ bool matches(const int32_t* __restrict__ o, int len) {
        if (len != N) return false;
        bool ok = true;
#pragma unroll
        for (int k = 0; k < N; ++k) ok &= (o[k] == at(k));
        return ok;
    }
};
using SeqEGWC = OpSeq<AISP_OP_EXPOSURE, AISP_OP_GAMMA, AISP_OP_WB, AISP_OP_CCM>;

// true when a compile-time instantiation owns this (already classified, per-pixel only) sequence
__device__ __forceinline__ bool owned_by_fixed(const int32_t* __restrict__ o, int len) {
    return SeqEGWC::matches(o, len);
}

// clip of a recomputed stage output on the FMA pipe: sat(y) + 0 * y is clip(y) for finite y and NaN for a
// non-finite one (what the reference's saved activation holds: its lerp term 0 * x poisons the pixel)
__device__ __forceinline__ float clip_next(float y) { return fmaf(0.f, y, __saturatef(y)); }

// reverse sweep from stage K down to stage 0 (compile-time recursion: the op is a template argument)
// clamp mask of a stage output as an all-ones / all-zeros word, taken in the forward sweep and ANDed onto the
// gradient in the reverse sweep: a value in a register, not a predicate that has to survive the sweep
// (inline PTX: written in C the compiler turns the word back into a predicate and then packs 18 of them into
// a register bit by bit)
__device__ __forceinline__ unsigned clip_pass_bits(float y) {
    unsigned m;
    asm("set.eq.u32.f32 %0, %1, %2;" : "=r"(m) : "f"(__saturatef(y)), "f"(y));   // FSET: 0xffffffff / 0, false for NaN
    return m;
}
__device__ __forceinline__ float and_bits(float g, unsigned m) {
    float r;
    asm("and.b32 %0, %1, %2;" : "=f"(r) : "f"(g), "r"(m));
    return r;
}

template <typename SEQ, int K, bool GIMG, bool CLIP, int kFixedPx>
__device__ __forceinline__ void fixed_reverse(const float (*sc)[kConst], const float (&X)[SEQ::N][3][kFixedPx],
                                              const unsigned (&M)[SEQ::N][3][kFixedPx], float (&gr)[kFixedPx],
                                              float (&gg)[kFixedPx], float (&gb)[kFixedPx], float (&acc)[SEQ::N][9]) {
    // the last stage's output only exists here: its body masks; earlier stages use the forward sweep's masks
    constexpr int MODE = !CLIP ? 0 : (K == SEQ::N - 1 ? 1 : 2);
#pragma unroll
    for (int v = 0; v < kFixedPx; ++v) {
        if (MODE == 2) { gr[v] = and_bits(gr[v], M[K][0][v]); gg[v] = and_bits(gg[v], M[K][1][v]); gb[v] = and_bits(gb[v], M[K][2][v]); }
        PwBwd<SEQ::at(K)>::template px<(K > 0) || GIMG, MODE>(sc[K], X[K][0][v], X[K][1][v], X[K][2][v], gr[v], gg[v], gb[v],
                                                              acc[K]);
    }
    if constexpr (K > 0) fixed_reverse<SEQ, K - 1, GIMG, CLIP, kFixedPx>(sc, X, M, gr, gg, gb, acc);
}

template <int VEC, bool GIMG, bool CLIP, typename SEQ, int MINB, int kFixedPx>
__global__ void __launch_bounds__(kThreads, MINB)
pw_chain_fixed_bwd_kernel(const float* __restrict__ img, const float* __restrict__ gout, const float* __restrict__ params,
                          const int32_t* __restrict__ ops, const int32_t* __restrict__ seq_len, int N, int S, int flags,
                          int rounds, int nchunks, float* __restrict__ gimg, float* __restrict__ partial) {
    pdl_prologue();
    extern __shared__ float4 dyn4[];
    float4* ring = dyn4;                               // [2 slots][6 planes: x r,g,b then g r,g,b][kThreads]
    constexpr int L = SEQ::N;
    __shared__ float raw[L][kConst];
    __shared__ float sc[L][kConst];
    __shared__ int sop[L];
    __shared__ float red[kWarps * AISP_ACC_STRIDE];
    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    int fill;
    const int len = classify_sequence(ops, seq_len, b, S, (flags & 2) != 0, false, &fill);
    if (len == 0 || !SEQ::matches(ops + (size_t)b * S, len)) return;   // the generic kernel owns (or fills) the sample
    stage_consts(params, ops, b, S, len, raw, sc, sop, BankMap{1, 0, 0ull, 0ull});
    const size_t base = (size_t)b * 3 * (size_t)N;
    const int chunk_px = rounds * kRoundPx;
    const float* pr = img + base;
    const float* pg = gout + base;
    float* gi = GIMG ? gimg + base : nullptr;

    for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int chunk0 = chunk * chunk_px;
        float acc[L][9];
#pragma unroll
        for (int k = 0; k < L; ++k)
#pragma unroll
            for (int j = 0; j < 9; ++j) acc[k][j] = 0.f;

        // this thread's six vectors of round r -> ring slot (r & 1).  VEC == 4: one running element offset per
        // thread, clamped so that the address of a (zero-filled) copy past the end stays inside the image
        auto issue = [&](int r) {
            const int i = chunk0 + (r * kThreads + tid) * kChainPx;
            float4* slot = ring + (size_t)(r & 1) * 6 * kThreads + tid;
            if (r < rounds) {
                if (VEC == 4) {
                    const bool ok = i < N;
                    const float* p0 = pr + (ok ? i : 0);
                    const float* g0 = pg + (ok ? i : 0);
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl) {
                        cp_async16_zfill(slot + pl * kThreads, p0 + (size_t)pl * N, ok);
                        cp_async16_zfill(slot + (3 + pl) * kThreads, g0 + (size_t)pl * N, ok);
                    }
                } else {
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl) {
#pragma unroll
                        for (int v = 0; v < kChainPx; ++v) {
                            const bool ok = i + v < N;
                            cp_async4_zfill(reinterpret_cast<float*>(slot + pl * kThreads) + v,
                                            ok ? pr + (size_t)pl * N + i + v : pr, ok);
                            cp_async4_zfill(reinterpret_cast<float*>(slot + (3 + pl) * kThreads) + v,
                                            ok ? pg + (size_t)pl * N + i + v : pg, ok);
                        }
                    }
                }
            }
            cp_async_commit();
        };

        issue(0);
        for (int r = 0; r < rounds; ++r) {
            issue(r + 1);
            cp_async_wait<1>();
            const int i = chunk0 + (r * kThreads + tid) * kChainPx;
            if (i >= N) continue;
            const float* slot = reinterpret_cast<const float*>(ring + (size_t)(r & 1) * 6 * kThreads + tid);
            // the four pixels of the ring vector go through forward + reverse sweep two at a time: half the
            // live stage inputs, and the loop body (not unrolled) is half the code
#pragma unroll 1
            for (int hv = 0; hv < kChainPx / kFixedPx; ++hv) {
                const int i2 = i + hv * kFixedPx;
                float X[L][3][kFixedPx];   // X[k] = input of stage k (X[0] straight from the ring)
                unsigned M[L][3][kFixedPx];   // clamp mask of stage k's output (stages 0 .. L-2)
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                    if (kFixedPx == 2) {
                        const float2 a = *reinterpret_cast<const float2*>(slot + 4 * pl * kThreads + 2 * hv);
                        X[0][pl][0] = a.x; X[0][pl][kFixedPx - 1] = a.y;
                    } else {
                        X[0][pl][0] = slot[4 * pl * kThreads + hv];
                    }
                }
                if (VEC == 1) {   // ragged tail: pixels past the end hold a harmless value (and get a zero gradient below)
#pragma unroll
                    for (int v = 0; v < kFixedPx; ++v)
                        if (i2 + v >= N) { X[0][0][v] = 0.5f; X[0][1][v] = 0.5f; X[0][2][v] = 0.5f; }
                }
#pragma unroll
                for (int k = 0; k + 1 < L; ++k) {
#pragma unroll
                    for (int v = 0; v < kFixedPx; ++v) {
                        float r_ = X[k][0][v], g_ = X[k][1][v], b_ = X[k][2][v];
                        fwd_px<false>(SEQ::at(k), sc[k], r_, g_, b_);
                        X[k + 1][0][v] = CLIP ? clip_next(r_) : r_;
                        X[k + 1][1][v] = CLIP ? clip_next(g_) : g_;
                        X[k + 1][2][v] = CLIP ? clip_next(b_) : b_;
                        if (CLIP) { M[k][0][v] = clip_pass_bits(r_); M[k][1][v] = clip_pass_bits(g_); M[k][2][v] = clip_pass_bits(b_); }
                    }
                }
                float gr[kFixedPx], gg[kFixedPx], gb[kFixedPx];
                if (kFixedPx == 2) {
                    const float2 a = *reinterpret_cast<const float2*>(slot + 4 * 3 * kThreads + 2 * hv);
                    const float2 c = *reinterpret_cast<const float2*>(slot + 4 * 4 * kThreads + 2 * hv);
                    const float2 d = *reinterpret_cast<const float2*>(slot + 4 * 5 * kThreads + 2 * hv);
                    gr[0] = a.x; gr[kFixedPx - 1] = a.y;
                    gg[0] = c.x; gg[kFixedPx - 1] = c.y;
                    gb[0] = d.x; gb[kFixedPx - 1] = d.y;
                } else {
                    gr[0] = slot[4 * 3 * kThreads + hv];
                    gg[0] = slot[4 * 4 * kThreads + hv];
                    gb[0] = slot[4 * 5 * kThreads + hv];
                }
                if (VEC == 1) {
#pragma unroll
                    for (int v = 0; v < kFixedPx; ++v)
                        if (i2 + v >= N) { gr[v] = 0.f; gg[v] = 0.f; gb[v] = 0.f; }
                }
                fixed_reverse<SEQ, L - 1, GIMG, CLIP, kFixedPx>(sc, X, M, gr, gg, gb, acc);
                if (GIMG) {
                    if (VEC == 4 && kFixedPx == 2) {
                        *reinterpret_cast<float2*>(gi + i2) = make_float2(gr[0], gr[kFixedPx - 1]);
                        *reinterpret_cast<float2*>(gi + N + i2) = make_float2(gg[0], gg[kFixedPx - 1]);
                        *reinterpret_cast<float2*>(gi + 2 * (size_t)N + i2) = make_float2(gb[0], gb[kFixedPx - 1]);
                    } else {
#pragma unroll
                        for (int v = 0; v < kFixedPx; ++v)
                            if (i2 + v < N) { gi[i2 + v] = gr[v]; gi[(size_t)N + i2 + v] = gg[v]; gi[2 * (size_t)N + i2 + v] = gb[v]; }
                    }
                }
            }
        }
        cp_async_wait<0>();
#pragma unroll
        for (int k = 0; k < L; ++k) {
            __syncthreads();   // `red` is reused stage after stage (and chunk after chunk)
            block_reduce_store_n<9>(acc[k], op_nacc(SEQ::at(k)), red,
                                    partial + (((size_t)b * nchunks + chunk) * S + k) * AISP_ACC_STRIDE);
        }
    }
}

template <int VEC, bool GIMG, int SMAX, int NACC, bool CLIP>
__global__ void __launch_bounds__(kThreads, NACC == 9 ? 2 : 1)
pw_chain_bwd_kernel(const float* __restrict__ img, const float* __restrict__ gout, const float* __restrict__ params,
                    const int32_t* __restrict__ ops, const int32_t* __restrict__ seq_len, int N, int S, int flags,
                    int rounds, int nchunks, float* __restrict__ gimg, float* __restrict__ partial) {
    pdl_prologue();
    extern __shared__ float4 dyn4[];
    float4* ring = dyn4;                               // [2 slots][6 planes: x r,g,b then g r,g,b][kThreads]
    float4* park = dyn4 + 2 * 6 * kThreads;            // [SMAX - 1 stages][3 planes][kThreads]
    __shared__ float raw[SMAX][kConst];
    __shared__ float sc[SMAX][kConst];
    __shared__ int sop[SMAX];
    __shared__ float red[kWarps * AISP_ACC_STRIDE];
    static_assert(SMAX <= kWarps, "one warp per step stages the constants");
    constexpr bool WITH_COLOR = (NACC == 27);
    const bool strict = (flags & 2) != 0;
    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    int fill;
    const int len = classify_sequence(ops, seq_len, b, S, strict, WITH_COLOR, &fill);
    const size_t base = (size_t)b * 3 * (size_t)N;
    const int chunk_px = rounds * kRoundPx;
    if (len != 0 && !WITH_COLOR && owned_by_fixed(ops + (size_t)b * S, len)) return;   // a compile-time instantiation's
    if (len == 0) {
        // the filling is done by the main instantiation only (the ColorFilter one would repeat it)
        if (GIMG && fill != 0 && !WITH_COLOR) {
            const float v = (fill == 2) ? __int_as_float(0x7fc00000) : 0.f;
            for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
                const int c0 = chunk * chunk_px, c1 = min(c0 + chunk_px, N);
                for (int pl = 0; pl < 3; ++pl)
                    for (int i = c0 + tid; i < c1; i += kThreads) {
                        const size_t o = base + (size_t)pl * N + i;
                        gimg[o] = (fill == 3) ? gout[o] : v;
                    }
            }
        }
        return;
    }
    stage_consts(params, ops, b, S, len, raw, sc, sop, BankMap{1, 0, 0ull, 0ull});

    const float* pr = img + base;
    const float* pg = gout + base;
    float* gi = GIMG ? gimg + base : nullptr;

    for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int chunk0 = chunk * chunk_px;
        float acc[SMAX][NACC];
#pragma unroll
        for (int k = 0; k < SMAX; ++k)
#pragma unroll
            for (int j = 0; j < NACC; ++j) acc[k][j] = 0.f;

        // this thread's six vectors of round r -> ring slot (r & 1)
        auto issue = [&](int r) {
            const int i = chunk0 + (r * kThreads + tid) * kChainPx;
            float4* slot = ring + (size_t)(r & 1) * 6 * kThreads + tid;
            if (r < rounds) {
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                    if (VEC == 4) {
                        const bool ok = i < N;
                        cp_async16_zfill(slot + pl * kThreads, ok ? pr + (size_t)pl * N + i : pr, ok);
                        cp_async16_zfill(slot + (3 + pl) * kThreads, ok ? pg + (size_t)pl * N + i : pg, ok);
                    } else {
#pragma unroll
                        for (int v = 0; v < kChainPx; ++v) {
                            const bool ok = i + v < N;
                            cp_async4_zfill(reinterpret_cast<float*>(slot + pl * kThreads) + v,
                                            ok ? pr + (size_t)pl * N + i + v : pr, ok);
                            cp_async4_zfill(reinterpret_cast<float*>(slot + (3 + pl) * kThreads) + v,
                                            ok ? pg + (size_t)pl * N + i + v : pg, ok);
                        }
                    }
                }
            }
            cp_async_commit();   // an empty group past the last round keeps the wait count uniform
        };

        issue(0);
        for (int r = 0; r < rounds; ++r) {
            issue(r + 1);
            cp_async_wait<1>();   // everything but the newest group has landed: round r is in the ring
            const int i = chunk0 + (r * kThreads + tid) * kChainPx;
            if (i >= N) continue;   // (the copies of later rounds are zero-fills: nothing to wait for)
            const float4* slot = ring + (size_t)(r & 1) * 6 * kThreads + tid;
            float R[kChainPx], G[kChainPx], B[kChainPx];
            {
                const float4 a = slot[0], c = slot[kThreads], d = slot[2 * kThreads];
                R[0] = a.x; R[1] = a.y; R[2] = a.z; R[3] = a.w;
                G[0] = c.x; G[1] = c.y; G[2] = c.z; G[3] = c.w;
                B[0] = d.x; B[1] = d.y; B[2] = d.z; B[3] = d.w;
            }
            if (VEC == 1) {   // ragged tail: pixels past the end repeat pixel 0 (and get a zero gradient below),
#pragma unroll           // so that they add 0 * (what a real pixel adds) to the partial sums
                for (int v = 1; v < kChainPx; ++v)
                    if (i + v >= N) { R[v] = R[0]; G[v] = G[0]; B[v] = B[0]; }
            }
            // forward sweep: park the input of stages 1 .. len-2; stage len-1's input stays in registers
#pragma unroll
            for (int k = 0; k < SMAX - 1; ++k) {
                if (k + 1 < len) {
                    if (k > 0) {
                        float4* p = park + (size_t)(k - 1) * 3 * kThreads + tid;
                        p[0] = make_float4(R[0], R[1], R[2], R[3]);
                        p[kThreads] = make_float4(G[0], G[1], G[2], G[3]);
                        p[2 * kThreads] = make_float4(B[0], B[1], B[2], B[3]);
                    }
                    fwd_step<kChainPx, false>(sop[k], sc[k], R, G, B);
                    if (CLIP) {
#pragma unroll
                        for (int v = 0; v < kChainPx; ++v) { R[v] = clip01(R[v]); G[v] = clip01(G[v]); B[v] = clip01(B[v]); }
                    }
                }
            }
            float gr[kChainPx], gg[kChainPx], gb[kChainPx];
            {
                const float4 a = slot[3 * kThreads], c = slot[4 * kThreads], d = slot[5 * kThreads];
                gr[0] = a.x; gr[1] = a.y; gr[2] = a.z; gr[3] = a.w;
                gg[0] = c.x; gg[1] = c.y; gg[2] = c.z; gg[3] = c.w;
                gb[0] = d.x; gb[1] = d.y; gb[2] = d.z; gb[3] = d.w;
            }
            if (VEC == 1) {   // ragged tail: pixels past the end carry no gradient
#pragma unroll
                for (int v = 0; v < kChainPx; ++v)
                    if (i + v >= N) { gr[v] = 0.f; gg[v] = 0.f; gb[v] = 0.f; }
            }
            // reverse sweep
#pragma unroll
            for (int k = SMAX - 1; k >= 0; --k) {
                if (k < len) {
                    if (k + 1 < len) {   // reload this stage's input (stage 0: still in the ring)
                        const float4* p = (k == 0) ? slot : park + (size_t)(k - 1) * 3 * kThreads + tid;
                        const float4 a = p[0], c = p[kThreads], d = p[2 * kThreads];
                        R[0] = a.x; R[1] = a.y; R[2] = a.z; R[3] = a.w;
                        G[0] = c.x; G[1] = c.y; G[2] = c.z; G[3] = c.w;
                        B[0] = d.x; B[1] = d.y; B[2] = d.z; B[3] = d.w;
                    }
                    if (k == 0 && !GIMG)
                        bwd_step<kChainPx, false, WITH_COLOR, CLIP>(sop[k], sc[k], R, G, B, gr, gg, gb, acc[k]);
                    else
                        bwd_step<kChainPx, true, WITH_COLOR, CLIP>(sop[k], sc[k], R, G, B, gr, gg, gb, acc[k]);
                }
            }
            if (GIMG) {
                if (VEC == 4) {
                    stg_stream4(gi + i, make_float4(gr[0], gr[1], gr[2], gr[3]));
                    stg_stream4(gi + N + i, make_float4(gg[0], gg[1], gg[2], gg[3]));
                    stg_stream4(gi + 2 * (size_t)N + i, make_float4(gb[0], gb[1], gb[2], gb[3]));
                } else {
#pragma unroll
                    for (int v = 0; v < kChainPx; ++v)
                        if (i + v < N) { gi[i + v] = gr[v]; gi[(size_t)N + i + v] = gg[v]; gi[2 * (size_t)N + i + v] = gb[v]; }
                }
            }
        }
        cp_async_wait<0>();
        // one scratch row per (sample, chunk, stage)
#pragma unroll
        for (int k = 0; k < SMAX; ++k) {
            if (k < len) {
                __syncthreads();   // `red` is reused stage after stage (and chunk after chunk)
                block_reduce_store_n<NACC>(acc[k], op_nacc(sop[k]), red,
                                           partial + (((size_t)b * nchunks + chunk) * S + k) * AISP_ACC_STRIDE);
            }
        }
    }
}

// grid = (B, S): sums the rows of (sample b, stage k) in fp64 and applies the stage's chain rule
__global__ void __launch_bounds__(kThreads)
chain_finalize_kernel(const float* __restrict__ partial, int nchunks, int S, int flags, const float* __restrict__ params,
                      const int32_t* __restrict__ ops, const int32_t* __restrict__ seq_len,
                      float* __restrict__ grad_params) {
    pdl_prologue();
    __shared__ double part[kWarps][AISP_ACC_STRIDE];
    __shared__ double tot[AISP_ACC_STRIDE];
    __shared__ float raw[kConst];
    __shared__ float c[kConst];
    const int b = blockIdx.x, k = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* gp_row = grad_params + ((size_t)b * S + k) * AISP_PSTRIDE;
    int fill;
    // the sequence is live for exactly one of the two instantiations: ask both
    int len = classify_sequence(ops, seq_len, b, S, (flags & 2) != 0, false, &fill);
    if (len == 0 && fill == 0) len = classify_sequence(ops, seq_len, b, S, (flags & 2) != 0, true, &fill);
    if (k >= len) {
        // not computed: exact zeros (idle steps, AISP_OP_NONE, samples of another family), NaN in strict
        // mode for a sequence this entry point cannot differentiate
        if (threadIdx.x < AISP_PSTRIDE) gp_row[threadIdx.x] = (fill == 2) ? __int_as_float(0x7fc00000) : 0.f;
        return;
    }
    const int op = ops[(size_t)b * S + k];
    double s = 0.0;
    for (int r = warp; r < nchunks; r += kWarps)
        s += (double)partial[(((size_t)b * nchunks + r) * S + k) * AISP_ACC_STRIDE + lane];
    part[warp][lane] = s;
    if (warp == 0) {
        raw[lane] = (lane < AISP_PSTRIDE) ? params[((size_t)b * S + k) * AISP_PSTRIDE + lane] : 0.f;
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
        for (int j = 0; j < AISP_PSTRIDE; ++j) gp_row[j] = gp[j];
    }
}

static inline bool aligned16c(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int VEC, bool GIMG, int SMAX, int NACC, bool CLIP>
static cudaError_t launch_one(dim3 grid, cudaStream_t st, const float* img, const float* gout, const float* params,
                              const int32_t* ops, const int32_t* seq_len, int N, int S, int flags, int rounds,
                              int nchunks, float* gimg, float* partial) {
    constexpr size_t smem = (size_t)(2 * 6 + (SMAX - 1) * 3) * kThreads * sizeof(float4);
    auto kern = pw_chain_bwd_kernel<VEC, GIMG, SMAX, NACC, CLIP>;
    static bool attr_set_on[64] = {};   // per device: function attributes belong to the device's context
    int devi = 0;
    cudaGetDevice(&devi);
    bool& attr_set = attr_set_on[devi & 63];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    launch_pdl_smem(kern, grid, kThreads, smem, st, img, gout, params, ops, seq_len, N, S, flags, rounds, nchunks, gimg,
                    partial);
    return cudaGetLastError();
}

template <int VEC, bool GIMG, bool CLIP, typename SEQ, int MINB, int FPX>
static cudaError_t launch_fixed(dim3 grid, cudaStream_t st, const float* img, const float* gout, const float* params,
                                const int32_t* ops, const int32_t* seq_len, int N, int S, int flags, int rounds,
                                int nchunks, float* gimg, float* partial) {
    constexpr size_t smem = (size_t)(2 * 6) * kThreads * sizeof(float4);
    auto kern = pw_chain_fixed_bwd_kernel<VEC, GIMG, CLIP, SEQ, MINB, FPX>;
    static bool attr_set_on[64] = {};
    int devi = 0;
    cudaGetDevice(&devi);
    bool& attr_set = attr_set_on[devi & 63];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    launch_pdl_smem(kern, grid, kThreads, smem, st, img, gout, params, ops, seq_len, N, S, flags, rounds, nchunks, gimg,
                    partial);
    return cudaGetLastError();
}

// rows of scratch per sample and stage for an H x W image (the launcher below uses the same rule)
static inline int chain_rounds(long long N, int B) {
    // 8 rounds per CTA (8192 px) amortise the per-chunk reduction; small problems keep 4 so that
    // there are enough CTAs to fill 148 SMs x 2
    const long long ctas8 = (long long)B * ((N + 8 * kRoundPx - 1) / (8 * kRoundPx));
    return ctas8 >= 2 * 296 ? 8 : 4;
}
int chain_rows(int B, int H, int W) {
    const long long N = (long long)H * W;
    const int rounds = chain_rounds(N, B);
    return (int)((N + (long long)rounds * kRoundPx - 1) / ((long long)rounds * kRoundPx));
}

cudaError_t launch_pointwise_chain_bwd(const float* img, const float* gout, const float* params, const int32_t* ops,
                                       const int32_t* seq_len, int B, int H, int W, int S, int flags,
                                       float* grad_params, float* grad_img, float* partial, cudaStream_t st) {
    const long long N = (long long)H * W;
    const int rounds = chain_rounds(N, B);
    const int nchunks = chain_rows(B, H, W);
    const dim3 grid((unsigned)nchunks, (unsigned)B), grid_color(1, (unsigned)B);
    const bool vec = (N % 4 == 0) && aligned16c(img) && aligned16c(gout) && (!grad_img || aligned16c(grad_img));
    cudaError_t e;
#define AISP_LAUNCH(VEC, GIMG, SMAX, NACC, GRID)                                                                   \
    ((flags & AISP_SEQ_CLIP)                                                                                       \
         ? launch_one<VEC, GIMG, SMAX, NACC, true>(GRID, st, img, gout, params, ops, seq_len, (int)N, S, flags, rounds, \
                                                   nchunks, grad_img, partial)                                   \
         : launch_one<VEC, GIMG, SMAX, NACC, false>(GRID, st, img, gout, params, ops, seq_len, (int)N, S, flags, rounds, \
                                                    nchunks, grad_img, partial))
    if (vec) {
        if (S <= 4) e = grad_img ? AISP_LAUNCH(4, true, 4, 9, grid) : AISP_LAUNCH(4, false, 4, 9, grid);
        else        e = grad_img ? AISP_LAUNCH(4, true, kChainMax, 9, grid) : AISP_LAUNCH(4, false, kChainMax, 9, grid);
        if (e != cudaSuccess) return e;
        e = grad_img ? AISP_LAUNCH(4, true, kChainMax, 27, grid_color) : AISP_LAUNCH(4, false, kChainMax, 27, grid_color);
    } else {
        e = grad_img ? AISP_LAUNCH(1, true, kChainMax, 9, grid) : AISP_LAUNCH(1, false, kChainMax, 9, grid);
        if (e != cudaSuccess) return e;
        e = grad_img ? AISP_LAUNCH(1, true, kChainMax, 27, grid_color) : AISP_LAUNCH(1, false, kChainMax, 27, grid_color);
    }
#undef AISP_LAUNCH
    if (e != cudaSuccess) return e;
    // compile-time sequences (samples the generic kernel skipped)
    if (S >= SeqEGWC::N) {
        // two pixels per pass at 80 registers (3 CTAs / SM): measured best of {1, 2} pixels x {3, 4} CTAs / SM
#define AISP_LAUNCH_FIXED(VEC, GIMG, CLIP)                                                                          \
    launch_fixed<VEC, GIMG, CLIP, SeqEGWC, 3, 2>(grid, st, img, gout, params, ops, seq_len, (int)N, S, flags, rounds, \
                                                 nchunks, grad_img, partial)
        const bool clip = (flags & AISP_SEQ_CLIP) != 0;
        if (vec) {
            if (grad_img) e = clip ? AISP_LAUNCH_FIXED(4, true, true) : AISP_LAUNCH_FIXED(4, true, false);
            else          e = clip ? AISP_LAUNCH_FIXED(4, false, true) : AISP_LAUNCH_FIXED(4, false, false);
        } else {
            if (grad_img) e = clip ? AISP_LAUNCH_FIXED(1, true, true) : AISP_LAUNCH_FIXED(1, true, false);
            else          e = clip ? AISP_LAUNCH_FIXED(1, false, true) : AISP_LAUNCH_FIXED(1, false, false);
        }
#undef AISP_LAUNCH_FIXED
        if (e != cudaSuccess) return e;
    }
    launch_pdl(chain_finalize_kernel, dim3(B, S), kThreads, st, partial, nchunks, S, flags, params, ops, seq_len,
               grad_params);
    return cudaGetLastError();
}

int chain_bwd_max_steps() { return kChainMax; }

}  // namespace aisp
