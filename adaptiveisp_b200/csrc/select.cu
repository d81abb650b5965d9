// Device-side filter selection and agent-state update.
//
// Reference: pdf_sample / one_hot (agent.py:12-23), the train / eval / forced-id selection
// (agent.py:126-149), the one-hot gather of the selected filter's parameters (agent.py:154 keeps one
// of ten images; here the parameter row is gathered BEFORE the filter runs) and the new agent state
// (agent.py:234-259).  The reference spends ~30 tiny ATen launches and ten boolean-index host syncs
// on this per step; here it is ONE launch with no host round trip, so selection -> apply is
// graph-capturable.  One thread per sample; every sum is a sequential fp32 loop in index order
// (what torch.cumsum / a 10-element torch.sum do), so the knife-edge compare `cdf < u` matches.
#include "aisp_common.cuh"

namespace aisp {

__global__ void __launch_bounds__(128)
select_kernel(const float* __restrict__ pdf, const float* __restrict__ noise, int mode, int forced,
              const float* __restrict__ states, const float* __restrict__ packed_all,
              const int32_t* __restrict__ op_table, int B, int F, int S, float test_steps, float early_stop_c,
              long long* __restrict__ sel_out, long long* __restrict__ one_hot, int32_t* __restrict__ ops_out,
              float* __restrict__ rows, float* __restrict__ new_states, float* __restrict__ penalties) {
    pdl_prologue();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* p = pdf + (size_t)b * F;
    int sel;
    if (mode == AISP_SELECT_FORCED) {
        sel = forced;
    } else if (mode == AISP_SELECT_ARGMAX) {      // torch.argmax: first maximum, NaN counts as maximal
        sel = 0;
        float best = p[0];
        for (int i = 1; i < F; ++i) {
            const float v = p[i];
            if (v > best || (v != v && best == best)) { best = v; sel = i; }
        }
    } else {                                      // pdf_sample, agent.py:12-16
        float tot = 0.f;
        for (int i = 0; i < F; ++i) tot += p[i];
        const float den = tot + 1e-36f;
        const float u = noise[b];
        float run = 0.f;
        int cnt = 0;
        for (int i = 0; i < F; ++i) {
            const float q = __fdiv_rn(p[i], den);
            run += q;                             // inclusive cumsum ...
            cnt += ((run - q) < u) ? 1 : 0;       // ... minus the entry itself, compared with the noise
        }
        sel = cnt - 1;
    }
    sel_out[b] = sel;
    const bool valid = (sel >= 0) && (sel < F);
    for (int i = 0; i < F; ++i) one_hot[(size_t)b * F + i] = (i == sel) ? 1 : 0;
    ops_out[b] = valid ? op_table[sel] : AISP_OP_NONE;
    // parameter row of the selected filter (zeros for the all-zero one-hot row of sel == -1)
    const float* src = packed_all + ((size_t)b * F + (valid ? sel : 0)) * AISP_PSTRIDE;
    for (int k = 0; k < AISP_PSTRIDE; ++k) rows[(size_t)b * AISP_PSTRIDE + k] = valid ? src[k] : 0.f;

    // new state (agent.py:234-259): [submitted, submitted, step + 1, max(usage, one_hot)]
    const float* s = states + (size_t)b * S;
    float* ns = new_states + (size_t)b * S;
    const float step = s[2];
    const float is_last = (fabsf((step + 1.f) - test_steps) < 1e-4f) ? 1.f : 0.f;
    ns[0] = is_last;
    ns[1] = is_last;
    ns[2] = step + 1.f;
    float usage_pen = 0.f;
    for (int i = 0; i < F; ++i) {
        const float hot = (i == sel) ? 1.f : 0.f;
        const float use = s[3 + i];
        usage_pen += use * hot;
        ns[3 + i] = (use != use) ? use : fmaxf(use, hot);   // torch.maximum propagates NaN
    }
    penalties[2 * b] = usage_pen;
    penalties[2 * b + 1] = ((1.f - is_last) * is_last) * early_stop_c;
}

// d L / d packed_all of the row gather: the upstream row lands in the selected slot, zeros elsewhere
// (exact zeros for the unselected filters, as torch.gather's backward gives the reference).
__global__ void __launch_bounds__(128)
select_bwd_kernel(const float* __restrict__ grad_rows, const long long* __restrict__ sel, int B, int F,
                  float* __restrict__ grad_packed_all) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (sample, slot, entry)
    if (i >= B * F * AISP_PSTRIDE) return;
    const int k = i % AISP_PSTRIDE, f = (i / AISP_PSTRIDE) % F, b = i / (AISP_PSTRIDE * F);
    grad_packed_all[i] = (sel[b] == f) ? grad_rows[(size_t)b * AISP_PSTRIDE + k] : 0.f;
}

cudaError_t launch_select(const float* pdf, const float* noise, int mode, int forced, const float* states,
                          const float* packed_all, const int32_t* op_table, int B, int F, int S, float test_steps,
                          float early_stop_c, long long* sel, long long* one_hot, int32_t* ops, float* rows,
                          float* new_states, float* penalties, cudaStream_t st) {
    launch_pdl(select_kernel, dim3((B + 127) / 128), dim3(128), st, pdf, noise, mode, forced, states, packed_all,
               op_table, B, F, S, test_steps, early_stop_c, sel, one_hot, ops, rows, new_states, penalties);
    return cudaGetLastError();
}

cudaError_t launch_select_bwd(const float* grad_rows, const long long* sel, int B, int F, float* grad_packed_all,
                              cudaStream_t st) {
    const int n = B * F * AISP_PSTRIDE;
    launch_pdl(select_bwd_kernel, dim3((n + 127) / 128), dim3(128), st, grad_rows, sel, B, F, grad_packed_all);
    return cudaGetLastError();
}

}  // namespace aisp
