// Fused per-pixel ISP pass for sm_100a: exposure, gamma, white balance, CCM, tone / colour curves,
// contrast, saturation+, desaturation -- forward over a per-sample op sequence in one pass over HBM,
// and the single-step backward (parameter gradients always, image gradient on request).
//
// Layout: CTA <-> (sample b = blockIdx.y, chunk of kPwChunkPx pixels = blockIdx.x).  The op id is
// uniform per CTA, so the `switch` never diverges.  Each thread streams 16 pixels as 3 planes x
// 4 x 128-bit loads (12 LDG.128 in flight), computes in registers, and streams the result back.
// HBM-bound: 24 B/px forward, 24 B/px backward (36 with grad_img).  No shared-memory staging is
// needed for the pixels (no reuse); shared memory only holds the per-step derived constants.
#include "aisp_common.cuh"

namespace aisp {

// =============================================================================================
// per-pixel forward math (in place).  c = derived constants of this step.
// =============================================================================================
// NOTE on rounding: this file is compiled with -fmad=false and the forward expressions below keep
// the reference's operation order (each product rounded, then added, as ATen does on the CPU).
// Saturated pixels (x == 1.0) land on y == 1.0 +- 1 ulp for the curve / CCM / desaturation /
// saturation filters, and the clamp backward of Filter.forward passes the gradient iff y <= 1: only
// bit-identical forward arithmetic reproduces the reference's gradient mask on those pixels.
// Explicit fmaf() is used only in gradient accumulators, where order is free.
__device__ __forceinline__ float lum_isp(float r, float g, float b) {  // isp/filters.py:12-14
    return (0.27f * r + 0.67f * g) + 0.06f * b;
}

__device__ __forceinline__ float curve8(float x, const float* c, int stride) {
    // 8 * sum_k clip(x - k/8, 0, 1/8) * p_k, k ascending (isp/filters.py:342-344).
    // clip(x - k/8, 0, 1/8) == sat(8x - k) / 8 exactly (power-of-two scaling commutes with every
    // rounding), so each knot is one FFMA.SAT + FMUL + FADD on the FMA pipe instead of
    // FADD + 2 FMNMX on the half-rate ALU pipe; the factor 8 is folded into the scale (c[8]/8).
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float u = __saturatef(fmaf(x, 8.0f, -(float)k));
        // fused multiply-add: bit-identical to ATen's separate multiply and add wherever the clamp
        // mask could flip -- saturated (x >= 1) and dark (x <= 0) pixels have u_k in {0, 1}, whose
        // products are exact -- and within 1 ulp elsewhere
        acc = fmaf(u, c[k * stride], acc);
    }
    return acc;
}

// correctly rounded a / 6 without the IEEE-division slow path (Markstein: q' = RN(q + r*y) with
// y = RN(1/6), q faithful); keeps floor(6*h) on the reference's side of every sextant boundary
__device__ __forceinline__ float div6(float a) {
    const float y = 0.16666667163372039794921875f;
    const float q = a * y;
    const float r = fmaf(-6.0f, q, a);
    return fmaf(r, y, q);
}

// HSV round trip of SaturationPlusFilter (isp/filters.py:445-560) for one pixel.
// Inputs r,g,b are already clipped to [0,1].  Outputs the "full colour" pixel and (for the
// backward) the intermediates needed by the reverse sweep.
struct HsvState {
    float mx, mn, d, num, sat, m, u, s2, s, vv, f;
    int branch;  // 0: R is max, 1: G, 2: B, -1: achromatic (hue forced to 0, no gradient)
    int sextant;
    bool satzero;
};

__device__ __forceinline__ void satplus_full(float r, float g, float b, float& fr, float& fg, float& fb,
                                             HsvState& st) {
    const float mx = fmaxf(r, fmaxf(g, b));
    const float mn = fminf(r, fminf(g, b));
    const float d = (mx - mn) + 1e-8f;
    float hue = 0.f, num = 0.f;
    int branch = -1;
    // ordered overwrites: B first, then G, then R -> R wins ties (isp/filters.py:456-464)
    float base = 0.f;
    if (b == mx) { num = r - g; base = 4.0f; branch = 2; }
    if (g == mx) { num = b - r; base = 2.0f; branch = 1; }
    if (r == mx) { num = g - b; base = 0.0f; branch = 0; }
    // one correctly rounded division for the winning branch (|q| <= 1): hue must land on the
    // reference's side of integer values, which decide the sextant and the gradient routing
    const float q = __fdiv_rn(num, d);
    hue = (branch == 0) ? ((q < 0.f) ? q + 6.0f : q)   // python-style q % 6
                        : base + q;
    if (mn == mx) { hue = 0.f; branch = -1; }
    hue = div6(hue);
    // saturation only feeds continuous expressions: fast reciprocal is enough
    float sat = __fdividef(mx - mn, mx + 1e-8f);
    const bool satzero = (mx == 0.f);
    if (satzero) sat = 0.f;
    // enhanced saturation (isp/filters.py:552)
    const float u = 0.5f - mx;
    const float m = 0.5f - fabsf(u);
    const float s2 = sat + (1.f - sat) * m * 0.8f;
    // hsv2rgb (isp/filters.py:481-533)
    const float h = (hue >= 1.0f) ? hue - 1.0f : hue;  // h % 1 for h in [0,1]
    const float s = clip01(s2);
    const float vv = clip01(mx);
    const float h6 = h * 6.0f;
    const float hi = floorf(h6);
    const float f = h6 - hi;
    const float pp = vv * (1.f - s);
    const float qq = vv * (1.f - (f * s));
    const float tt = vv * (1.f - ((1.f - f) * s));
    const int sx = (int)hi;
    switch (sx) {
    case 0: fr = vv; fg = tt; fb = pp; break;
    case 1: fr = qq; fg = vv; fb = pp; break;
    case 2: fr = pp; fg = vv; fb = tt; break;
    case 3: fr = pp; fg = qq; fb = vv; break;
    case 4: fr = tt; fg = pp; fb = vv; break;
    case 5: fr = vv; fg = pp; fb = qq; break;
    default: fr = 0.f; fg = 0.f; fb = 0.f; break;
    }
    st.mx = mx; st.mn = mn; st.d = d; st.num = num; st.sat = sat; st.m = m; st.u = u; st.s2 = s2;
    st.s = s; st.vv = vv; st.f = f; st.branch = branch; st.sextant = sx; st.satzero = satzero;
}

// Forward-only form of the same round trip: identical hue / saturation arithmetic, but hsv -> rgb
// uses the branch-free k-form  c_n = v - v*s*sat(min(k, 4 - k)),  k = (n + 6h) mod 6,  n = 5, 3, 1
// (algebraically the sextant table of isp/filters.py:505-527) -- 6 ALU-pipe ops instead of the
// floor / float->int / 6-way select, which made this filter ALU-bound.  Results agree with the
// table form to ~1 ulp; the backward keeps the table form because it needs the sextant for routing.
__device__ __forceinline__ void satplus_forward(float r, float g, float b, float& fr, float& fg, float& fb) {
    const float mx = fmaxf(r, fmaxf(g, b));
    const float mn = fminf(r, fminf(g, b));
    const float d = (mx - mn) + 1e-8f;
    float num, base;
    if (r == mx) { num = g - b; base = 0.0f; }
    else if (g == mx) { num = b - r; base = 2.0f; }
    else { num = r - g; base = 4.0f; }
    const float q = __fdiv_rn(num, d);
    float hue = base + q;
    hue = (hue < 0.f) ? hue + 6.0f : hue;          // only the R branch can go negative: python-style % 6
    if (mn == mx) hue = 0.f;
    float h6 = div6(hue);                           // same rounding as the reference's hue / 6 ...
    h6 = ((h6 >= 1.0f) ? h6 - 1.0f : h6) * 6.0f;    // ... then (h % 1) * 6
    float sat = __fdividef(mx - mn, mx + 1e-8f);
    if (mx == 0.f) sat = 0.f;
    const float m = 0.5f - fabsf(0.5f - mx);
    const float s = clip01(sat + (1.f - sat) * m * 0.8f);
    const float nvs = -(mx * s);                    // v is already in [0,1]
    float k;
    k = h6 + 5.0f; k = (k >= 6.0f) ? k - 6.0f : k; fr = fmaf(nvs, __saturatef(fminf(k, 4.0f - k)), mx);
    k = h6 + 3.0f; k = (k >= 6.0f) ? k - 6.0f : k; fg = fmaf(nvs, __saturatef(fminf(k, 4.0f - k)), mx);
    k = h6 + 1.0f; k = (k >= 6.0f) ? k - 6.0f : k; fb = fmaf(nvs, __saturatef(fminf(k, 4.0f - k)), mx);
}

template <int NPX>
__device__ __forceinline__ void fwd_step(int op, const float* __restrict__ c, float (&R)[NPX], float (&G)[NPX],
                                         float (&B)[NPX]) {
    switch (op) {
    case AISP_OP_EXPOSURE: {
        const float s = c[0];
#pragma unroll
        for (int i = 0; i < NPX; ++i) { R[i] *= s; G[i] *= s; B[i] *= s; }
        break;
    }
    case AISP_OP_GAMMA: {  // pow(max(x, 0.001), p) via lg2/ex2 (MUFU): |err| << 1e-5 on [0,1]
        const float p = c[0];
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            R[i] = exp2f(p * __log2f(fmaxf(R[i], 0.001f)));
            G[i] = exp2f(p * __log2f(fmaxf(G[i], 0.001f)));
            B[i] = exp2f(p * __log2f(fmaxf(B[i], 0.001f)));
        }
        break;
    }
    case AISP_OP_WB: {
        const float s0 = c[0], s1 = c[1], s2 = c[2];
#pragma unroll
        for (int i = 0; i < NPX; ++i) { R[i] *= s0; G[i] *= s1; B[i] *= s2; }
        break;
    }
    case AISP_OP_CCM: {  // out_i = sum_j M[i][j] x_j   (isp/filters.py:666-672)
        float m[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) m[k] = c[k];
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            const float r = R[i], g = G[i], b = B[i];
            R[i] = (r * m[0] + g * m[1]) + b * m[2];
            G[i] = (r * m[3] + g * m[4]) + b * m[5];
            B[i] = (r * m[6] + g * m[7]) + b * m[8];
        }
        break;
    }
    case AISP_OP_TONE: {
        const float sc = c[8] * 0.125f;  // exact: undoes the factor 8 carried by curve8()
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            R[i] = curve8(R[i], c, 1) * sc;
            G[i] = curve8(G[i], c, 1) * sc;
            B[i] = curve8(B[i], c, 1) * sc;
        }
        break;
    }
    case AISP_OP_COLOR: {
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            R[i] = curve8(R[i], c + 0, 3) * (c[24] * 0.125f);
            G[i] = curve8(G[i], c + 1, 3) * (c[25] * 0.125f);
            B[i] = curve8(B[i], c + 2, 3) * (c[26] * 0.125f);
        }
        break;
    }
    case AISP_OP_CONTRAST: {  // isp/filters.py:415-419
        const float p = c[0], ip = 1.f - p;
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            const float l = clip01(lum_isp(R[i], G[i], B[i]));
            const float cl = -__cosf(AISP_PIF * l) * 0.5f + 0.5f;
            const float inv = 1.0f / (l + 1e-6f);
            R[i] = ip * R[i] + p * (R[i] * inv * cl);
            G[i] = ip * G[i] + p * (G[i] * inv * cl);
            B[i] = ip * B[i] + p * (B[i] * inv * cl);
        }
        break;
    }
    case AISP_OP_WNB: {  // isp/filters.py:435-437
        const float p = c[0], ip = 1.f - p;
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            const float l = lum_isp(R[i], G[i], B[i]);
            R[i] = ip * R[i] + p * l;
            G[i] = ip * G[i] + p * l;
            B[i] = ip * B[i] + p * l;
        }
        break;
    }
    case AISP_OP_SATPLUS: {
        const float p = c[0], ip = 1.f - p;
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            const float r = clip01(R[i]), g = clip01(G[i]), b = clip01(B[i]);
            float fr, fg, fb;
            satplus_forward(r, g, b, fr, fg, fb);
            R[i] = r * ip + fr * p;
            G[i] = g * ip + fg * p;
            B[i] = b * ip + fb * p;
        }
        break;
    }
    default: break;
    }
}

// =============================================================================================
// per-pixel backward of one step.  (gr,gg,gb): upstream gradient in, image gradient out (GIMG).
// acc: raw per-thread partial sums, turned into parameter gradients by finalize_grads().
// =============================================================================================
template <int OP>
struct PwBwd;

#define AISP_MASK_CLIP(yr, yg, yb)                                   \
    if (clip) { gr *= pass01(yr); gg *= pass01(yg); gb *= pass01(yb); }

template <>
struct PwBwd<AISP_OP_EXPOSURE> {
    static constexpr int NACC = 1;
    template <bool GIMG>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, int clip, float* acc) {
        const float s = c[0];
        AISP_MASK_CLIP(r * s, g * s, b * s)
        acc[0] = fmaf(gr, r, fmaf(gg, g, fmaf(gb, b, acc[0])));
        if (GIMG) { gr *= s; gg *= s; gb *= s; }
    }
};

template <>
struct PwBwd<AISP_OP_GAMMA> {
    static constexpr int NACC = 1;
    template <bool GIMG>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, int clip, float* acc) {
        const float p = c[0];
        const float xr = fmaxf(r, 0.001f), xg = fmaxf(g, 0.001f), xb = fmaxf(b, 0.001f);
        const float lr = __log2f(xr), lg = __log2f(xg), lb = __log2f(xb);
        const float yr = exp2f(p * lr), yg = exp2f(p * lg), yb = exp2f(p * lb);
        AISP_MASK_CLIP(yr, yg, yb)
        acc[0] = fmaf(gr * yr, lr, fmaf(gg * yg, lg, fmaf(gb * yb, lb, acc[0])));  // x ln2 in finalize
        if (GIMG) {  // p * x^(p-1), only where the min-clamp passed (x >= 0.001, inclusive)
            gr = (r >= 0.001f) ? gr * p * __fdividef(yr, xr) : 0.f;
            gg = (g >= 0.001f) ? gg * p * __fdividef(yg, xg) : 0.f;
            gb = (b >= 0.001f) ? gb * p * __fdividef(yb, xb) : 0.f;
        }
    }
};

template <>
struct PwBwd<AISP_OP_WB> {
    static constexpr int NACC = 3;
    template <bool GIMG>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, int clip, float* acc) {
        AISP_MASK_CLIP(r * c[0], g * c[1], b * c[2])
        acc[0] = fmaf(gr, r, acc[0]); acc[1] = fmaf(gg, g, acc[1]); acc[2] = fmaf(gb, b, acc[2]);
        if (GIMG) { gr *= c[0]; gg *= c[1]; gb *= c[2]; }
    }
};

template <>
struct PwBwd<AISP_OP_CCM> {
    static constexpr int NACC = 9;
    template <bool GIMG>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, int clip, float* acc) {
        const float yr = (r * c[0] + g * c[1]) + b * c[2];
        const float yg = (r * c[3] + g * c[4]) + b * c[5];
        const float yb = (r * c[6] + g * c[7]) + b * c[8];
        AISP_MASK_CLIP(yr, yg, yb)
        acc[0] = fmaf(gr, r, acc[0]); acc[1] = fmaf(gr, g, acc[1]); acc[2] = fmaf(gr, b, acc[2]);
        acc[3] = fmaf(gg, r, acc[3]); acc[4] = fmaf(gg, g, acc[4]); acc[5] = fmaf(gg, b, acc[5]);
        acc[6] = fmaf(gb, r, acc[6]); acc[7] = fmaf(gb, g, acc[7]); acc[8] = fmaf(gb, b, acc[8]);
        if (GIMG) {  // M^T gy
            const float xr = c[0] * gr + c[3] * gg + c[6] * gb;
            const float xg = c[1] * gr + c[4] * gg + c[7] * gb;
            const float xb = c[2] * gr + c[5] * gg + c[8] * gb;
            gr = xr; gg = xg; gb = xb;
        }
    }
};

// one channel of a curve filter.  u_k = sat(8x - k) = 8 * clip(x - k/8, 0, 1/8) (see curve8);
// accumulates g*u_k (8x the segment sums, undone in finalize_grads) and g*y.
template <bool GIMG>
__device__ __forceinline__ void curve8_bwd(float x, const float* c, int stride, float sc, float& g, int clip,
                                           float* acc, int astride, float& yacc) {
    float u[8], v[8];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = fmaf(x, 8.0f, -(float)k);
        u[k] = __saturatef(v[k]);
        sum = fmaf(u[k], c[k * stride], sum);
    }
    const float y = sum * (sc * 0.125f);
    if (clip) g *= pass01(y);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k * astride] = fmaf(g, u[k], acc[k * astride]);
    yacc = fmaf(g, y, yacc);
    if (GIMG) {  // clamp backward is inclusive at both ends: on a knot two segments pass
        float slope = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) slope += (v[k] >= 0.f && v[k] <= 1.0f) ? c[k * stride] : 0.f;
        g = g * sc * slope;
    }
}

template <>
struct PwBwd<AISP_OP_TONE> {
    static constexpr int NACC = 9;
    template <bool GIMG>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, int clip, float* acc) {
        curve8_bwd<GIMG>(r, c, 1, c[8], gr, clip, acc, 1, acc[8]);
        curve8_bwd<GIMG>(g, c, 1, c[8], gg, clip, acc, 1, acc[8]);
        curve8_bwd<GIMG>(b, c, 1, c[8], gb, clip, acc, 1, acc[8]);
    }
};

template <>
struct PwBwd<AISP_OP_COLOR> {
    static constexpr int NACC = 27;
    template <bool GIMG>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, int clip, float* acc) {
        curve8_bwd<GIMG>(r, c + 0, 3, c[24], gr, clip, acc + 0, 3, acc[24]);
        curve8_bwd<GIMG>(g, c + 1, 3, c[25], gg, clip, acc + 1, 3, acc[25]);
        curve8_bwd<GIMG>(b, c + 2, 3, c[26], gb, clip, acc + 2, 3, acc[26]);
    }
};

template <>
struct PwBwd<AISP_OP_CONTRAST> {
    static constexpr int NACC = 1;
    template <bool GIMG>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, int clip, float* acc) {
        const float p = c[0], ip = 1.f - p;
        const float l0 = lum_isp(r, g, b);
        const float l = clip01(l0);
        float sn, cs;
        __sincosf(AISP_PIF * l, &sn, &cs);
        const float cl = -cs * 0.5f + 0.5f;
        const float den = l + 1e-6f;
        const float inv = 1.0f / den;
        const float cr = r * inv * cl, cg = g * inv * cl, cb = b * inv * cl;
        AISP_MASK_CLIP(ip * r + p * cr, ip * g + p * cg, ip * b + p * cb)
        acc[0] = fmaf(gr, cr - r, fmaf(gg, cg - g, fmaf(gb, cb - b, acc[0])));
        if (GIMG) {
            const float ratio = cl * inv;
            const float dratio = (0.5f * AISP_PIF * sn * den - cl) * inv * inv;
            const float common = p * dratio * pass01(l0) * (gr * r + gg * g + gb * b);
            const float k = ip + p * ratio;
            gr = gr * k + common * 0.27f;
            gg = gg * k + common * 0.67f;
            gb = gb * k + common * 0.06f;
        }
    }
};

template <>
struct PwBwd<AISP_OP_WNB> {
    static constexpr int NACC = 1;
    template <bool GIMG>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, int clip, float* acc) {
        const float p = c[0], ip = 1.f - p;
        const float l = lum_isp(r, g, b);
        AISP_MASK_CLIP(ip * r + p * l, ip * g + p * l, ip * b + p * l)
        acc[0] = fmaf(gr, l - r, fmaf(gg, l - g, fmaf(gb, l - b, acc[0])));
        if (GIMG) {
            const float s = p * (gr + gg + gb);
            gr = ip * gr + s * 0.27f;
            gg = ip * gg + s * 0.67f;
            gb = ip * gb + s * 0.06f;
        }
    }
};

template <>
struct PwBwd<AISP_OP_SATPLUS> {
    static constexpr int NACC = 1;
    template <bool GIMG>
    static __device__ __forceinline__ void px(const float* c, float r0, float g0, float b0, float& gr, float& gg,
                                              float& gb, int clip, float* acc) {
        const float p = c[0], ip = 1.f - p;
        const float r = clip01(r0), g = clip01(g0), b = clip01(b0);
        float fr, fg, fb;
        HsvState st;
        if (GIMG) satplus_full(r, g, b, fr, fg, fb, st);     // the reverse sweep needs the sextant / f / s
        else satplus_forward(r, g, b, fr, fg, fb);            // d y / d p = full - x only needs the colour
        AISP_MASK_CLIP(r * ip + fr * p, g * ip + fg * p, b * ip + fb * p)
        acc[0] = fmaf(gr, fr - r, fmaf(gg, fg - g, fmaf(gb, fb - b, acc[0])));
        if (GIMG) {
            // reverse sweep through hsv2rgb -> enhanced saturation -> rgb2hsv -> leading clip
            float cr = gr * ip, cg = gg * ip, cb = gb * ip;
            const float ar = gr * p, ag = gg * p, ab = gb * p;
            float gv = 0.f, gp = 0.f, gq = 0.f, gt = 0.f;
            switch (st.sextant) {
            case 0: gv = ar; gt = ag; gp = ab; break;
            case 1: gq = ar; gv = ag; gp = ab; break;
            case 2: gp = ar; gv = ag; gt = ab; break;
            case 3: gp = ar; gq = ag; gv = ab; break;
            case 4: gt = ar; gp = ag; gv = ab; break;
            case 5: gv = ar; gp = ag; gq = ab; break;
            default: break;
            }
            const float s = st.s, f = st.f, vv = st.vv;
            float gvv = gv + gp * (1.f - s) + gq * (1.f - f * s) + gt * (1.f - (1.f - f) * s);
            const float gs = -vv * (gp + gq * f + gt * (1.f - f));
            const float ghue6 = vv * s * (gt - gq);      // d/df; h*6 and hue/6 cancel, % has slope 1
            const float gs2 = gs * pass01(st.s2);
            float gmx = gvv * pass01(st.mx);
            const float gsat = gs2 * (1.f - st.m * 0.8f);
            const float gm = gs2 * (1.f - st.sat) * 0.8f;
            const float sgn = (st.u > 0.f) ? 1.f : ((st.u < 0.f) ? -1.f : 0.f);  // abs'(0) = 0
            gmx += gm * sgn;
            float gmn = 0.f;
            if (!st.satzero) {
                const float den = st.mx + 1e-8f;
                gmx += gsat * (1.0f / den - (st.mx - st.mn) / (den * den));
                gmn -= gsat / den;
            }
            if (st.branch >= 0) {
                const float gnum = ghue6 / st.d;
                const float gd = -ghue6 * st.num / (st.d * st.d);
                gmx += gd;
                gmn -= gd;
                if (st.branch == 0) { cg += gnum; cb -= gnum; }
                else if (st.branch == 1) { cb += gnum; cr -= gnum; }
                else { cr += gnum; cg -= gnum; }
            }
            // max / min route their gradient to the first arg-extremum in R,G,B order
            if (r == st.mx) cr += gmx; else if (g == st.mx) cg += gmx; else cb += gmx;
            if (r == st.mn) cr += gmn; else if (g == st.mn) cg += gmn; else cb += gmn;
            gr = cr * pass01(r0);
            gg = cg * pass01(g0);
            gb = cb * pass01(b0);
        }
    }
};

// =============================================================================================
// kernels
// =============================================================================================
// Bulk L2 prefetch of one CTA chunk of a 3-plane image (threads 0..2, one plane each): the rounds
// after the first then see L2 latency instead of DRAM latency.  Needs 16-byte alignment (VEC == 4).
__device__ __forceinline__ void prefetch_chunk_l2(const float* __restrict__ base3, int N, int chunk0) {
    if (threadIdx.x < 3 && chunk0 + kPwChunkPx <= N) {
        const float* p = base3 + (size_t)threadIdx.x * N + chunk0;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((unsigned)(kPwChunkPx * sizeof(float))) : "memory");
    }
}
__device__ __forceinline__ void stage_consts(const float* __restrict__ params, const int32_t* __restrict__ ops,
                                             int b, int S, int len, float (*raw)[kConst], float (*sc)[kConst],
                                             int* sop, const BankMap& bm) {
    // one warp per step loads the raw row, lane 0 derives the constants
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < len) {
        raw[warp][lane] = (lane < AISP_PSTRIDE) ? params[((size_t)b * S + warp) * AISP_PSTRIDE + lane] : 0.f;
        sc[warp][lane] = 0.f;
        __syncwarp();
        if (lane == 0) {
            const int op = ops ? ops[(size_t)b * S + warp] : bank_op(bm, b);
            sop[warp] = op;
            derive_consts(op, raw[warp], sc[warp]);
        }
    }
    __syncthreads();
}

template <int VEC>
__global__ void __launch_bounds__(kThreads, 4)
pw_fwd_kernel(const float* __restrict__ img, float* __restrict__ out, const float* __restrict__ params,
              const int32_t* __restrict__ ops, const int32_t* __restrict__ seq_len, int N, int S, int clip_each,
              BankMap bm) {
    pdl_prologue();
    static_assert(AISP_MAX_STEPS <= kWarps, "one warp per step stages the constants");
    __shared__ float raw[AISP_MAX_STEPS][kConst];
    __shared__ float sc[AISP_MAX_STEPS][kConst];
    __shared__ int sop[AISP_MAX_STEPS];
    const int b = bank_sample(bm, blockIdx.y);
    int len = seq_len ? min(max(seq_len[b], 0), S) : S;
    const int op0 = sample_op(ops, bm, b, S);
    if (len > 0 && !is_pointwise(op0)) {
        // another family owns this sample -- except AISP_OP_NONE: the all-zero one-hot row of
        // agent.py:18-23,154 (pdf_sample returned -1), whose gathered image is exactly zero
        if (op0 == AISP_OP_NONE) {
            float* q = out + (size_t)b * 3 * (size_t)N;
            const int c0 = blockIdx.x * kPwChunkPx;
            for (int pl = 0; pl < 3; ++pl)
                for (int i = c0 + threadIdx.x; i < min(c0 + kPwChunkPx, N); i += kThreads) q[(size_t)pl * N + i] = 0.f;
        }
        return;
    }
    stage_consts(params, ops, b, S, len, raw, sc, sop, bm);
    // a stencil op inside a sequence terminates it (documented in the header)
    for (int k = 0; k < len; ++k)
        if (!is_pointwise(sop[k])) { len = k; break; }

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

// =============================================================================================
// Fused multi-step backward: a per-sample sequence of up to kChainMax per-pixel filters is
// differentiated in ONE pass over HBM.  Per pixel the forward chain is recomputed in registers
// (keeping every stage's input), then the stages are swept in reverse: each stage turns the
// upstream gradient into its parameter-gradient partial sums and into the gradient w.r.t. its
// input, which is the upstream gradient of the previous stage.  Traffic: image 12 + grad_out 12
// (+12 for grad_img) bytes per pixel for the whole sequence -- the same as ONE single-step backward.
// The 24-knot ColorFilter (27 partial sums) does not fit the 9-accumulator stage budget: its
// gradient row is returned as NaN (the image gradient still flows through it correctly).
// =============================================================================================
constexpr int kChainMax = 4;
constexpr int kChainAcc = 9;

template <int NPX, bool GX>
__device__ __forceinline__ void bwd_step(int op, const float* __restrict__ c, const float (&R)[NPX],
                                         const float (&G)[NPX], const float (&B)[NPX], float (&gr)[NPX],
                                         float (&gg)[NPX], float (&gb)[NPX], int clip, float* acc) {
    switch (op) {
#define AISP_CHAIN_CASE(OPC)                                                                         \
    case OPC: {                                                                                      \
        _Pragma("unroll") for (int i = 0; i < NPX; ++i)                                              \
            PwBwd<OPC>::template px<GX>(c, R[i], G[i], B[i], gr[i], gg[i], gb[i], clip, acc);        \
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
#undef AISP_CHAIN_CASE
    case AISP_OP_COLOR: {  // image gradient only; the parameter-gradient row is flagged NaN
        float scratch[PwBwd<AISP_OP_COLOR>::NACC];
#pragma unroll
        for (int k = 0; k < PwBwd<AISP_OP_COLOR>::NACC; ++k) scratch[k] = 0.f;
#pragma unroll
        for (int i = 0; i < NPX; ++i)
            PwBwd<AISP_OP_COLOR>::template px<GX>(c, R[i], G[i], B[i], gr[i], gg[i], gb[i], clip, scratch);
        acc[0] = __int_as_float(0x7fc00000);
        break;
    }
    default: break;
    }
}

template <int VEC, bool GIMG>
__global__ void __launch_bounds__(kThreads, 2)
pw_chain_bwd_kernel(const float* __restrict__ img, const float* __restrict__ gout, const float* __restrict__ params,
                    const int32_t* __restrict__ ops, const int32_t* __restrict__ seq_len, int N, int S, int clip_each,
                    float* __restrict__ gimg, float* __restrict__ partial) {
    pdl_prologue();
    __shared__ float raw[kChainMax][kConst];
    __shared__ float sc[kChainMax][kConst];
    __shared__ int sop[kChainMax];
    __shared__ float red[kWarps * AISP_ACC_STRIDE];
    const int b = blockIdx.y;
    int len = seq_len ? min(max(seq_len[b], 0), S) : S;
    if (len > 0 && !is_pointwise(ops[(size_t)b * S])) {
        if (GIMG && ops[(size_t)b * S] == AISP_OP_NONE) {
            float* q = gimg + (size_t)b * 3 * (size_t)N;
            const int c0 = blockIdx.x * kPwChunkPx;
            for (int pl = 0; pl < 3; ++pl)
                for (int i = c0 + threadIdx.x; i < min(c0 + kPwChunkPx, N); i += kThreads) q[(size_t)pl * N + i] = 0.f;
        }
        return;
    }
    stage_consts(params, ops, b, S, len, raw, sc, sop, BankMap{1, 0, 0ull, 0ull});
    for (int k = 0; k < len; ++k)
        if (!is_pointwise(sop[k])) { len = k; break; }

    constexpr int GROUPS = kPwChunkPx / (kThreads * VEC);
    const size_t base = (size_t)b * 3 * (size_t)N;
    const float* pr = img + base;
    const float* pg = gout + base;
    float* gi = GIMG ? gimg + base : nullptr;
    const int chunk0 = blockIdx.x * kPwChunkPx;
    if (VEC == 4) {   // long compute per round (recompute + reverse sweep): later rounds then find their pixels in L2
        prefetch_chunk_l2(pr, N, chunk0);
        prefetch_chunk_l2(pg, N, chunk0);
    }
    float acc[kChainMax][kChainAcc];
#pragma unroll
    for (int k = 0; k < kChainMax; ++k)
#pragma unroll
        for (int j = 0; j < kChainAcc; ++j) acc[k][j] = 0.f;

    for (int g0 = 0; g0 < GROUPS; ++g0) {
        const int i = chunk0 + (g0 * kThreads + threadIdx.x) * VEC;
        if (i >= N) break;
        Pack<VEC> t;
        float xs[kChainMax][3][VEC];
        float R[VEC], G[VEC], B[VEC], gr[VEC], gg[VEC], gb[VEC];
        t.load(pr + i);
#pragma unroll
        for (int v = 0; v < VEC; ++v) R[v] = t.v[v];
        t.load(pr + N + i);
#pragma unroll
        for (int v = 0; v < VEC; ++v) G[v] = t.v[v];
        t.load(pr + 2 * (size_t)N + i);
#pragma unroll
        for (int v = 0; v < VEC; ++v) B[v] = t.v[v];
        t.load(pg + i);
#pragma unroll
        for (int v = 0; v < VEC; ++v) gr[v] = t.v[v];
        t.load(pg + N + i);
#pragma unroll
        for (int v = 0; v < VEC; ++v) gg[v] = t.v[v];
        t.load(pg + 2 * (size_t)N + i);
#pragma unroll
        for (int v = 0; v < VEC; ++v) gb[v] = t.v[v];
        // forward recompute, remembering every stage's input
#pragma unroll
        for (int k = 0; k < kChainMax; ++k) {
            if (k < len) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { xs[k][0][v] = R[v]; xs[k][1][v] = G[v]; xs[k][2][v] = B[v]; }
                if (k + 1 < len) {  // the last stage's output is never needed
                    fwd_step<VEC>(sop[k], sc[k], R, G, B);
                    if (clip_each) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) { R[v] = clip01(R[v]); G[v] = clip01(G[v]); B[v] = clip01(B[v]); }
                    }
                }
            }
        }
        // reverse sweep
#pragma unroll
        for (int k = kChainMax - 1; k >= 0; --k) {
            if (k < len) {
                if (k == 0 && !GIMG)
                    bwd_step<VEC, false>(sop[k], sc[k], xs[k][0], xs[k][1], xs[k][2], gr, gg, gb, clip_each, acc[k]);
                else
                    bwd_step<VEC, true>(sop[k], sc[k], xs[k][0], xs[k][1], xs[k][2], gr, gg, gb, clip_each, acc[k]);
            }
        }
        if (GIMG) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) t.v[v] = gr[v];
            t.store(gi + i);
#pragma unroll
            for (int v = 0; v < VEC; ++v) t.v[v] = gg[v];
            t.store(gi + N + i);
#pragma unroll
            for (int v = 0; v < VEC; ++v) t.v[v] = gb[v];
            t.store(gi + 2 * (size_t)N + i);
        }
    }
    // one scratch row per (sample, chunk, stage)
#pragma unroll
    for (int k = 0; k < kChainMax; ++k) {
        if (k < S) {
            __syncthreads();  // `red` is reused stage after stage
            block_reduce_store<kChainAcc>(acc[k], red,
                                          partial + (((size_t)b * gridDim.x + blockIdx.x) * S + k) * AISP_ACC_STRIDE);
        }
    }
}

// grid = (B, S): sums the rows of (sample b, stage k) in fp64 and applies the stage's chain rule
__global__ void __launch_bounds__(kThreads)
chain_finalize_kernel(const float* __restrict__ partial, int nchunks, int S, const float* __restrict__ params,
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
    int len = seq_len ? min(max(seq_len[b], 0), S) : S;
    bool live = (k < len) && is_pointwise(ops[(size_t)b * S]);
    for (int j = 0; live && j <= k; ++j) live = is_pointwise(ops[(size_t)b * S + j]);
    if (!live) {
        if (threadIdx.x < AISP_PSTRIDE) gp_row[threadIdx.x] = 0.f;
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
        if (op == AISP_OP_COLOR)
            for (int j = 0; j < AISP_PSTRIDE; ++j) gp[j] = __int_as_float(0x7fc00000);
        for (int j = 0; j < AISP_PSTRIDE; ++j) gp_row[j] = gp[j];
    }
}

cudaError_t launch_pointwise_chain_bwd(const float* img, const float* gout, const float* params, const int32_t* ops,
                                       const int32_t* seq_len, int B, int H, int W, int S, int clip_each,
                                       float* grad_params, float* grad_img, float* partial, cudaStream_t st) {
    const long long N = (long long)H * W;
    const int rows = pointwise_rows(H, W);
    dim3 grid((unsigned)rows, (unsigned)B);
    const bool vec = (N % 4 == 0) && aligned16(img) && aligned16(gout) && (!grad_img || aligned16(grad_img));
    if (vec) {
        if (grad_img) launch_pdl(pw_chain_bwd_kernel<4, true>, grid, kThreads, st, img, gout, params, ops, seq_len, (int)N, S, clip_each, grad_img, partial);
        else launch_pdl(pw_chain_bwd_kernel<4, false>, grid, kThreads, st, img, gout, params, ops, seq_len, (int)N, S, clip_each, nullptr, partial);
    } else {
        if (grad_img) launch_pdl(pw_chain_bwd_kernel<1, true>, grid, kThreads, st, img, gout, params, ops, seq_len, (int)N, S, clip_each, grad_img, partial);
        else launch_pdl(pw_chain_bwd_kernel<1, false>, grid, kThreads, st, img, gout, params, ops, seq_len, (int)N, S, clip_each, nullptr, partial);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    launch_pdl(chain_finalize_kernel, dim3(B, S), kThreads, st, partial, rows, S, params, ops, seq_len, grad_params);
    return cudaGetLastError();
}

int chain_bwd_max_steps() { return kChainMax; }

}  // namespace aisp
