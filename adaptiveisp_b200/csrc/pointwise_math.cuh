// Per-pixel filter arithmetic shared by the per-pixel, fused-sequence and stencil-prologue kernels:
// forward of one filter on one pixel (fwd_px / fwd_step) and the single-step backward bodies
// (PwBwd<OP>).  Every translation unit that includes this file is compiled with -fmad=false (see the
// rounding note below and the Makefile).
#pragma once

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
// Bare MUFU lg2 / ex2 (no denormal pre/post-scaling code around them): gamma only takes log2 of values
// >= 0.001 and its exponent p * log2(x) stays far above -126, where these are bit-identical to
// __log2f / exp2f and save four instructions per call.
__device__ __forceinline__ float lg2_mufu(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_mufu(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// bare MUFU reciprocal for arguments known to be normal (gamma: x >= 0.001): __fdividef wraps the same
// MUFU in denormal-range scaling code (five more instructions per call)
__device__ __forceinline__ float rcp_mufu(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// clamp backward as ATen does it (where(lo <= y <= hi, g, 0)): a select, so a non-finite upstream
// gradient on a clipped pixel is dropped, not turned into NaN
// 0 <= y <= 1  <=>  sat(y) == y (NaN: sat gives 0, the compare fails; -0 passes): one FADD.SAT on the FMA pipe
// and one compare instead of two compares and a predicate merge on the half-rate ALU pipe
__device__ __forceinline__ float mask01(float y, float g) { return (__saturatef(y) == y) ? g : 0.f; }

__device__ __forceinline__ float lum_isp(float r, float g, float b) {  // isp/filters.py:12-14
    return (0.27f * r + 0.67f * g) + 0.06f * b;
}

__device__ __forceinline__ float curve8(float x, const float* c, int stride) {
    // 8 * sum_k clip(x - k/8, 0, 1/8) * p_k, k ascending (isp/filters.py:342-344).
    // clip(x - k/8, 0, 1/8) == sat(8x - k) / 8 exactly (power-of-two scaling commutes with every
    // rounding), so each knot is one FFMA.SAT + FMUL + FADD on the FMA pipe instead of
    // FADD + 2 FMNMX on the half-rate ALU pipe; the factor 8 is folded into the scale (c[8]/8).
    // Knot 0 goes through the NaN-propagating clip (ATen's clamp keeps NaN, .SAT would flush it to 0):
    // a NaN pixel then poisons the whole sum, as in the reference, at the cost of one ALU op.
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float u = (k == 0) ? clip01(x * 8.0f) : __saturatef(fmaf(x, 8.0f, -(float)k));
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
    const float mx = max_nan(r, max_nan(g, b));   // image.max(1) / .min(1) keep NaN: a NaN channel
    const float mn = min_nan(r, min_nan(g, b));   // poisons the whole pixel, as in the reference
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
    const float mx = max_nan(r, max_nan(g, b));
    const float mn = min_nan(r, min_nan(g, b));
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

// One filter applied to one pixel: (r,g,b) in, (r,g,b) out.  POISON adds the `(1 - mask) * img` term of
// the reference's lerp (isp/filters.py:115,138 with mask == 1): 0 * x + y is y bit for bit for finite
// x, and NaN when x is inf / NaN -- which is how a non-finite pixel survives every filter in the
// reference (and trips the pool-refill guard of train.py:374).  One FFMA per channel.
template <bool POISON>
__device__ __forceinline__ void fwd_px(int op, const float* __restrict__ c, float& R, float& G, float& B) {
    const float r = R, g = G, b = B;
    float yr = r, yg = g, yb = b;
    switch (op) {
    case AISP_OP_EXPOSURE: {
        const float s = c[0];
        yr = r * s; yg = g * s; yb = b * s;
        break;
    }
    case AISP_OP_GAMMA: {  // pow(max(x, 0.001), p) via lg2/ex2 (MUFU): |err| << 1e-5 on [0,1]
        const float p = c[0];
        yr = ex2_mufu(p * lg2_mufu(max_nan(r, 0.001f)));   // torch.max(img, 0.001) keeps NaN
        yg = ex2_mufu(p * lg2_mufu(max_nan(g, 0.001f)));
        yb = ex2_mufu(p * lg2_mufu(max_nan(b, 0.001f)));
        break;
    }
    case AISP_OP_WB: {
        yr = r * c[0]; yg = g * c[1]; yb = b * c[2];
        break;
    }
    case AISP_OP_CCM: {  // out_i = sum_j M[i][j] x_j   (isp/filters.py:666-672)
        yr = (r * c[0] + g * c[1]) + b * c[2];
        yg = (r * c[3] + g * c[4]) + b * c[5];
        yb = (r * c[6] + g * c[7]) + b * c[8];
        break;
    }
    case AISP_OP_TONE: {
        const float sc = c[8] * 0.125f;  // exact: undoes the factor 8 carried by curve8()
        yr = curve8(r, c, 1) * sc;
        yg = curve8(g, c, 1) * sc;
        yb = curve8(b, c, 1) * sc;
        break;
    }
    case AISP_OP_COLOR: {
        yr = curve8(r, c + 0, 3) * (c[24] * 0.125f);
        yg = curve8(g, c + 1, 3) * (c[25] * 0.125f);
        yb = curve8(b, c + 2, 3) * (c[26] * 0.125f);
        break;
    }
    case AISP_OP_CONTRAST: {  // isp/filters.py:415-419
        const float p = c[0], ip = 1.f - p;
        const float l = clip01(lum_isp(r, g, b));
        const float cl = -__cosf(AISP_PIF * l) * 0.5f + 0.5f;
        const float inv = 1.0f / (l + 1e-6f);
        yr = ip * r + p * (r * inv * cl);
        yg = ip * g + p * (g * inv * cl);
        yb = ip * b + p * (b * inv * cl);
        break;
    }
    case AISP_OP_WNB: {  // isp/filters.py:435-437
        const float p = c[0], ip = 1.f - p;
        const float l = lum_isp(r, g, b);
        yr = ip * r + p * l;
        yg = ip * g + p * l;
        yb = ip * b + p * l;
        break;
    }
    case AISP_OP_SATPLUS: {
        const float p = c[0], ip = 1.f - p;
        const float cr = clip01(r), cg = clip01(g), cb = clip01(b);
        float fr, fg, fb;
        satplus_forward(cr, cg, cb, fr, fg, fb);
        yr = cr * ip + fr * p;
        yg = cg * ip + fg * p;
        yb = cb * ip + fb * p;
        break;
    }
    default: break;
    }
    if (POISON) { yr = fmaf(0.f, r, yr); yg = fmaf(0.f, g, yg); yb = fmaf(0.f, b, yb); }
    R = yr; G = yg; B = yb;
}

// Two pixels at a time as packed fp32 (FMUL2 / FADD2 / FFMA2: one issue slot per two results).  Lane-wise
// this is fwd_px: the same operations in the same order with the same (IEEE round-to-nearest) rounding, so
// the results are bit-identical; MUFU, min / max, compares and the HSV round trip stay per lane.  A forward
// pass keeps no accumulators, so packing costs no registers.
__device__ __forceinline__ f32x2 lum_isp2(f32x2 r, f32x2 g, f32x2 b) {
    return add2_sep(add2_sep(mul2(splat2(0.27f), r), mul2(splat2(0.67f), g)), mul2(splat2(0.06f), b));
}
__device__ __forceinline__ f32x2 curve8_2(f32x2 x, const float* c, int stride) {
    f32x2 acc = splat2(0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const f32x2 u = (k == 0) ? pack2(clip01(lo2(x) * 8.0f), clip01(hi2(x) * 8.0f))
                                 : pack2(__saturatef(fmaf(lo2(x), 8.0f, -(float)k)), __saturatef(fmaf(hi2(x), 8.0f, -(float)k)));
        acc = fma2(u, splat2(c[k * stride]), acc);
    }
    return acc;
}

template <bool POISON>
__device__ __forceinline__ void fwd_px2(int op, const float* __restrict__ c, f32x2& R, f32x2& G, f32x2& B) {
    const f32x2 r = R, g = G, b = B;
    f32x2 yr = r, yg = g, yb = b;
    switch (op) {
    case AISP_OP_EXPOSURE: {
        const f32x2 s = splat2(c[0]);
        yr = mul2(r, s); yg = mul2(g, s); yb = mul2(b, s);
        break;
    }
    case AISP_OP_GAMMA: {
        const f32x2 p = splat2(c[0]);
        const f32x2 tr = mul2(p, pack2(lg2_mufu(max_nan(lo2(r), 0.001f)), lg2_mufu(max_nan(hi2(r), 0.001f))));
        const f32x2 tg = mul2(p, pack2(lg2_mufu(max_nan(lo2(g), 0.001f)), lg2_mufu(max_nan(hi2(g), 0.001f))));
        const f32x2 tb = mul2(p, pack2(lg2_mufu(max_nan(lo2(b), 0.001f)), lg2_mufu(max_nan(hi2(b), 0.001f))));
        yr = pack2(ex2_mufu(lo2(tr)), ex2_mufu(hi2(tr)));
        yg = pack2(ex2_mufu(lo2(tg)), ex2_mufu(hi2(tg)));
        yb = pack2(ex2_mufu(lo2(tb)), ex2_mufu(hi2(tb)));
        break;
    }
    case AISP_OP_WB: {
        yr = mul2(r, splat2(c[0])); yg = mul2(g, splat2(c[1])); yb = mul2(b, splat2(c[2]));
        break;
    }
    case AISP_OP_CCM: {
        yr = add2_sep(add2_sep(mul2(r, splat2(c[0])), mul2(g, splat2(c[1]))), mul2(b, splat2(c[2])));
        yg = add2_sep(add2_sep(mul2(r, splat2(c[3])), mul2(g, splat2(c[4]))), mul2(b, splat2(c[5])));
        yb = add2_sep(add2_sep(mul2(r, splat2(c[6])), mul2(g, splat2(c[7]))), mul2(b, splat2(c[8])));
        break;
    }
    case AISP_OP_TONE: {
        const f32x2 sc = splat2(c[8] * 0.125f);
        yr = mul2(curve8_2(r, c, 1), sc);
        yg = mul2(curve8_2(g, c, 1), sc);
        yb = mul2(curve8_2(b, c, 1), sc);
        break;
    }
    case AISP_OP_COLOR: {
        yr = mul2(curve8_2(r, c + 0, 3), splat2(c[24] * 0.125f));
        yg = mul2(curve8_2(g, c + 1, 3), splat2(c[25] * 0.125f));
        yb = mul2(curve8_2(b, c + 2, 3), splat2(c[26] * 0.125f));
        break;
    }
    case AISP_OP_CONTRAST: {
        const float p = c[0], ip = 1.f - p;
        const f32x2 l0 = lum_isp2(r, g, b);
        const float la = clip01(lo2(l0)), lb = clip01(hi2(l0));
        const f32x2 cl = pack2(-__cosf(AISP_PIF * la) * 0.5f + 0.5f, -__cosf(AISP_PIF * lb) * 0.5f + 0.5f);
        const f32x2 inv = pack2(1.0f / (la + 1e-6f), 1.0f / (lb + 1e-6f));
        const f32x2 ip2 = splat2(ip), p2 = splat2(p);
        yr = add2_sep(mul2(ip2, r), mul2(p2, mul2(mul2(r, inv), cl)));
        yg = add2_sep(mul2(ip2, g), mul2(p2, mul2(mul2(g, inv), cl)));
        yb = add2_sep(mul2(ip2, b), mul2(p2, mul2(mul2(b, inv), cl)));
        break;
    }
    case AISP_OP_WNB: {
        const float p = c[0], ip = 1.f - p;
        const f32x2 pl = mul2(splat2(p), lum_isp2(r, g, b)), ip2 = splat2(ip);
        yr = add2_sep(mul2(ip2, r), pl);
        yg = add2_sep(mul2(ip2, g), pl);
        yb = add2_sep(mul2(ip2, b), pl);
        break;
    }
    case AISP_OP_SATPLUS: {
        const float p = c[0], ip = 1.f - p;
        const float ar = clip01(lo2(r)), ag = clip01(lo2(g)), ab = clip01(lo2(b));
        const float br = clip01(hi2(r)), bg = clip01(hi2(g)), bb = clip01(hi2(b));
        float f0r, f0g, f0b, f1r, f1g, f1b;
        satplus_forward(ar, ag, ab, f0r, f0g, f0b);
        satplus_forward(br, bg, bb, f1r, f1g, f1b);
        const f32x2 ip2 = splat2(ip), p2 = splat2(p);
        yr = add2_sep(mul2(pack2(ar, br), ip2), mul2(pack2(f0r, f1r), p2));
        yg = add2_sep(mul2(pack2(ag, bg), ip2), mul2(pack2(f0g, f1g), p2));
        yb = add2_sep(mul2(pack2(ab, bb), ip2), mul2(pack2(f0b, f1b), p2));
        break;
    }
    default: break;
    }
    if (POISON) {
        const f32x2 z = splat2(0.f);
        yr = fma2(z, r, yr); yg = fma2(z, g, yg); yb = fma2(z, b, yb);
    }
    R = yr; G = yg; B = yb;
}

// NPX pixels of one thread through one filter; the op switch is CTA-uniform and hoisted by the
// compiler out of the (unrolled) pixel loop.  Even NPX: pixel pairs (2i, 2i+1) go through fwd_px2.
template <int NPX, bool POISON = true>
__device__ __forceinline__ void fwd_step(int op, const float* __restrict__ c, float (&R)[NPX], float (&G)[NPX],
                                         float (&B)[NPX]) {
    switch (op) {
#define AISP_FWD_CASE(OPC)                                                                   \
    case OPC: {                                                                              \
        if constexpr (NPX % 2 == 0) {                                                        \
            _Pragma("unroll") for (int i = 0; i < NPX; i += 2) {                             \
                f32x2 r2 = pack2(R[i], R[i + (NPX > 1)]), g2 = pack2(G[i], G[i + (NPX > 1)]), b2 = pack2(B[i], B[i + (NPX > 1)]); \
                fwd_px2<POISON>(OPC, c, r2, g2, b2);                                         \
                R[i] = lo2(r2); R[i + (NPX > 1)] = hi2(r2);                                  \
                G[i] = lo2(g2); G[i + (NPX > 1)] = hi2(g2);                                  \
                B[i] = lo2(b2); B[i + (NPX > 1)] = hi2(b2);                                  \
            }                                                                                \
        } else {                                                                             \
            _Pragma("unroll") for (int i = 0; i < NPX; ++i) fwd_px<POISON>(OPC, c, R[i], G[i], B[i]); \
        }                                                                                    \
        break;                                                                               \
    }
        AISP_FWD_CASE(AISP_OP_EXPOSURE)
        AISP_FWD_CASE(AISP_OP_GAMMA)
        AISP_FWD_CASE(AISP_OP_WB)
        AISP_FWD_CASE(AISP_OP_CCM)
        AISP_FWD_CASE(AISP_OP_TONE)
        AISP_FWD_CASE(AISP_OP_COLOR)
        AISP_FWD_CASE(AISP_OP_CONTRAST)
        AISP_FWD_CASE(AISP_OP_WNB)
        AISP_FWD_CASE(AISP_OP_SATPLUS)
#undef AISP_FWD_CASE
    default: break;
    }
}

// =============================================================================================
// per-pixel backward of one step.  (gr,gg,gb): upstream gradient in, image gradient out (GIMG).
// acc: raw per-thread partial sums, turned into parameter gradients by finalize_grads().
// =============================================================================================
template <int OP>
struct PwBwd;

// CLIP: 0 = no clip after the step (Filter.run), 1 = clip, the body masks the upstream gradient with the
// step's own output, 2 = clip, the caller has masked the gradient already (compile-time sequences)
#define AISP_MASK_CLIP(yr, yg, yb)                                   \
    if (CLIP == 1) { gr = mask01(yr, gr); gg = mask01(yg, gg); gb = mask01(yb, gb); }

template <>
struct PwBwd<AISP_OP_EXPOSURE> {
    static constexpr int NACC = 1;
    template <bool GIMG, int CLIP>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, float* acc) {
        const float s = c[0];
        AISP_MASK_CLIP(r * s, g * s, b * s)
        acc[0] = fmaf(gr, r, fmaf(gg, g, fmaf(gb, b, acc[0])));
        if (GIMG) { gr *= s; gg *= s; gb *= s; }
    }
};

template <>
struct PwBwd<AISP_OP_GAMMA> {
    static constexpr int NACC = 1;
    template <bool GIMG, int CLIP>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, float* acc) {
        const float p = c[0];
        const float xr = max_nan(r, 0.001f), xg = max_nan(g, 0.001f), xb = max_nan(b, 0.001f);
        const float lr = lg2_mufu(xr), lg = lg2_mufu(xg), lb = lg2_mufu(xb);
        const float yr = ex2_mufu(p * lr), yg = ex2_mufu(p * lg), yb = ex2_mufu(p * lb);
        AISP_MASK_CLIP(yr, yg, yb)
        const float tr = gr * yr, tg = gg * yg, tb = gb * yb;
        acc[0] = fmaf(tr, lr, fmaf(tg, lg, fmaf(tb, lb, acc[0])));  // x ln2 in finalize
        if (GIMG) {  // g * p * x^(p-1) = (g * y) * (p / x), only where the min-clamp passed (x >= 0.001, inclusive)
            gr = (r >= 0.001f) ? tr * (p * rcp_mufu(xr)) : 0.f;
            gg = (g >= 0.001f) ? tg * (p * rcp_mufu(xg)) : 0.f;
            gb = (b >= 0.001f) ? tb * (p * rcp_mufu(xb)) : 0.f;
        }
    }
};

template <>
struct PwBwd<AISP_OP_WB> {
    static constexpr int NACC = 3;
    template <bool GIMG, int CLIP>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, float* acc) {
        AISP_MASK_CLIP(r * c[0], g * c[1], b * c[2])
        acc[0] = fmaf(gr, r, acc[0]); acc[1] = fmaf(gg, g, acc[1]); acc[2] = fmaf(gb, b, acc[2]);
        if (GIMG) { gr *= c[0]; gg *= c[1]; gb *= c[2]; }
    }
};

template <>
struct PwBwd<AISP_OP_CCM> {
    static constexpr int NACC = 9;
    template <bool GIMG, int CLIP>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, float* acc) {
        const float yr = (r * c[0] + g * c[1]) + b * c[2];
        const float yg = (r * c[3] + g * c[4]) + b * c[5];
        const float yb = (r * c[6] + g * c[7]) + b * c[8];
        AISP_MASK_CLIP(yr, yg, yb)
        acc[0] = fmaf(gr, r, acc[0]); acc[1] = fmaf(gr, g, acc[1]); acc[2] = fmaf(gr, b, acc[2]);
        acc[3] = fmaf(gg, r, acc[3]); acc[4] = fmaf(gg, g, acc[4]); acc[5] = fmaf(gg, b, acc[5]);
        acc[6] = fmaf(gb, r, acc[6]); acc[7] = fmaf(gb, g, acc[7]); acc[8] = fmaf(gb, b, acc[8]);
        if (GIMG) {  // M^T gy
            const float xr = fmaf(c[6], gb, fmaf(c[3], gg, c[0] * gr));
            const float xg = fmaf(c[7], gb, fmaf(c[4], gg, c[1] * gr));
            const float xb = fmaf(c[8], gb, fmaf(c[5], gg, c[2] * gr));
            gr = xr; gg = xg; gb = xb;
        }
    }
};

// one channel of a curve filter.  u_k = sat(8x - k) = 8 * clip(x - k/8, 0, 1/8) (see curve8);
// accumulates g*u_k (8x the segment sums, undone in finalize_grads) and g*y.
template <bool GIMG, int CLIP>
__device__ __forceinline__ void curve8_bwd(float x, const float* c, int stride, float sc, float& g,
                                           float* acc, int astride, float& yacc) {
    float u[8], v[8];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = fmaf(x, 8.0f, -(float)k);
        u[k] = (k == 0) ? clip01(v[k]) : __saturatef(v[k]);   // NaN-propagating first knot, see curve8()
        sum = fmaf(u[k], c[k * stride], sum);
    }
    const float y = sum * (sc * 0.125f);
    if (CLIP == 1) g = mask01(y, g);
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
    template <bool GIMG, int CLIP>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, float* acc) {
        curve8_bwd<GIMG, CLIP>(r, c, 1, c[8], gr, acc, 1, acc[8]);
        curve8_bwd<GIMG, CLIP>(g, c, 1, c[8], gg, acc, 1, acc[8]);
        curve8_bwd<GIMG, CLIP>(b, c, 1, c[8], gb, acc, 1, acc[8]);
    }
};

template <>
struct PwBwd<AISP_OP_COLOR> {
    static constexpr int NACC = 27;
    template <bool GIMG, int CLIP>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, float* acc) {
        curve8_bwd<GIMG, CLIP>(r, c + 0, 3, c[24], gr, acc + 0, 3, acc[24]);
        curve8_bwd<GIMG, CLIP>(g, c + 1, 3, c[25], gg, acc + 1, 3, acc[25]);
        curve8_bwd<GIMG, CLIP>(b, c + 2, 3, c[26], gb, acc + 2, 3, acc[26]);
    }
};

template <>
struct PwBwd<AISP_OP_CONTRAST> {
    static constexpr int NACC = 1;
    template <bool GIMG, int CLIP>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, float* acc) {
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
    template <bool GIMG, int CLIP>
    static __device__ __forceinline__ void px(const float* c, float r, float g, float b, float& gr, float& gg,
                                              float& gb, float* acc) {
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
    template <bool GIMG, int CLIP>
    static __device__ __forceinline__ void px(const float* c, float r0, float g0, float b0, float& gr, float& gg,
                                              float& gb, float* acc) {
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

// position of the (first) stencil step of sample b's sequence, or -1; *end = effective length (a
// second stencil step ends the sequence: the host-side planner splits such sequences into phases)
__device__ __forceinline__ int find_stencil(const int32_t* __restrict__ o, int len, int* end) {
    int pos = -1;
    *end = len;
    for (int k = 0; k < len; ++k) {
        const int op = o[k];
        if (!is_pointwise(op)) {
            if (pos < 0 && (is_sharpen(op) || op == AISP_OP_NLM)) pos = k;
            else { *end = k; break; }
        }
    }
    return pos;
}

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

}  // namespace aisp
