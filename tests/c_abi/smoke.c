/* Plain-C consumer of the C ABI: no Python, no torch -- only include/aisp_b200.h, libaisp_b200.so and the
 * CUDA runtime.  Exercises a fused 3-step per-pixel sequence, its single-step backward, the 3x3 sharpen
 * and NLM on a small batch and checks closed-form expectations.  Exit code 0 = all checks passed.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include tests/c_abi/smoke.c -o smoke \
 *       -L adaptiveisp_b200/csrc -laisp_b200 -L /usr/local/cuda/lib64 -lcudart -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "aisp_b200.h"

#define CK(x)                                                                       \
    do {                                                                            \
        int rc_ = (int)(x);                                                         \
        if (rc_ != 0) { fprintf(stderr, "%s -> %d (%s)\n", #x, rc_, aisp_status_string(rc_)); return 2; } \
    } while (0)

static int fail(const char* what, double got, double want) {
    fprintf(stderr, "FAIL %s: got %.9g want %.9g\n", what, got, want);
    return 1;
}

int main(void) {
    const int B = 2, H = 24, W = 40, N = H * W, n = B * 3 * N;
    float *h_img = malloc(sizeof(float) * n), *h_out = malloc(sizeof(float) * n), *h_g = malloc(sizeof(float) * n);
    for (int i = 0; i < n; ++i) { h_img[i] = (float)((i * 37) % 101) / 100.0f; h_g[i] = 1.0f; }
    float *d_img, *d_out, *d_g, *d_par, *d_gp, *d_stash;
    int32_t* d_ops;
    void* d_scr;
    CK(cudaMalloc((void**)&d_img, sizeof(float) * n));
    CK(cudaMalloc((void**)&d_out, sizeof(float) * n));
    CK(cudaMalloc((void**)&d_g, sizeof(float) * n));
    CK(cudaMalloc((void**)&d_stash, sizeof(float) * n));
    CK(cudaMalloc((void**)&d_par, sizeof(float) * B * AISP_MAX_STEPS * AISP_PSTRIDE));
    CK(cudaMalloc((void**)&d_gp, sizeof(float) * B * AISP_PSTRIDE));
    CK(cudaMalloc((void**)&d_ops, sizeof(int32_t) * B * AISP_MAX_STEPS));
    size_t scr = aisp_bwd_scratch_bytes(B, H, W);
    CK(cudaMalloc(&d_scr, scr));
    CK(cudaMemset(d_scr, 0, scr));
    CK(cudaMemcpy(d_img, h_img, sizeof(float) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_g, h_g, sizeof(float) * n, cudaMemcpyHostToDevice));
    if (aisp_version() != 3 || aisp_op_num_params(AISP_OP_CCM) != 9) return fail("version/params", aisp_version(), 3);

    /* 1. fused sequence: exposure +1 EV -> white balance (0.5,0.5,0.5) -> exposure 0 == identity */
    const int S = 3;
    float par[2 * 3 * AISP_PSTRIDE];
    int32_t ops[2 * 3];
    memset(par, 0, sizeof(par));
    for (int b = 0; b < B; ++b) {
        ops[b * S + 0] = AISP_OP_EXPOSURE; par[(b * S + 0) * AISP_PSTRIDE] = 1.0f;
        ops[b * S + 1] = AISP_OP_WB;
        par[(b * S + 1) * AISP_PSTRIDE + 0] = par[(b * S + 1) * AISP_PSTRIDE + 1] = par[(b * S + 1) * AISP_PSTRIDE + 2] = 0.5f;
        ops[b * S + 2] = AISP_OP_EXPOSURE; par[(b * S + 2) * AISP_PSTRIDE] = 0.0f;
    }
    CK(cudaMemcpy(d_par, par, sizeof(par), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ops, ops, sizeof(ops), cudaMemcpyHostToDevice));
    CK(aisp_pointwise_fwd(d_img, d_out, d_par, d_ops, NULL, B, H, W, S, 0, NULL));
    CK(cudaMemcpy(h_out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i)
        if (fabsf(h_out[i] - h_img[i]) > 1e-6f) return fail("fused E,WB,E identity", h_out[i], h_img[i]);

    /* 2. backward of exposure p=0 with g=1, no clip: dL/dp = ln2 * sum(x) per sample */
    float par1[2 * AISP_PSTRIDE];
    int32_t ops1[2] = {AISP_OP_EXPOSURE, AISP_OP_EXPOSURE};
    memset(par1, 0, sizeof(par1));
    CK(cudaMemcpy(d_par, par1, sizeof(par1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ops, ops1, sizeof(ops1), cudaMemcpyHostToDevice));
    CK(aisp_pointwise_bwd(d_img, d_g, d_par, d_ops, B, H, W, 0, d_gp, NULL, d_scr, scr, NULL));
    float gp[2 * AISP_PSTRIDE];
    CK(cudaMemcpy(gp, d_gp, sizeof(gp), cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b) {
        double s = 0;
        for (int i = 0; i < 3 * N; ++i) s += h_img[b * 3 * N + i];
        if (fabs(gp[b * AISP_PSTRIDE] - s * 0.6931471805599453) > 1e-4 * s) return fail("exposure grad", gp[b * AISP_PSTRIDE], s * 0.693147);
    }

    /* 3. 3x3 sharpen with factor 1 is the identity (y = x*1 + blur*0); NLM keeps a constant image */
    int32_t ops2[2] = {AISP_OP_SHARPEN, AISP_OP_NLM};
    par1[0] = 1.0f; par1[AISP_PSTRIDE] = 0.5f;
    for (int i = 0; i < 3 * N; ++i) h_img[3 * N + i] = 0.25f;       /* sample 1: constant */
    CK(cudaMemcpy(d_img, h_img, sizeof(float) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_par, par1, sizeof(par1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ops, ops2, sizeof(ops2), cudaMemcpyHostToDevice));
    CK(aisp_select_apply_fwd(d_img, d_out, d_par, d_ops, B, H, W, 1, d_stash, NULL, NULL));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 3 * N; ++i)
        if (fabsf(h_out[i] - h_img[i]) > 1e-6f) return fail("sharpen factor 1", h_out[i], h_img[i]);
    for (int i = 3 * N; i < n; ++i)
        if (fabsf(h_out[i] - 0.25f) > 1e-6f) return fail("nlm constant image", h_out[i], 0.25);

    /* 3b. filter bank: {exposure 0 EV, white balance (1,1,1), 3x3 sharpen factor 1} on the same batch
     *     -> a [B,3,3,H,W] stack of three copies of the image; the backward of the exposure slot with
     *     g = 1 is ln2 * sum(x) again, the two other slots receive their own rows */
    {
        const int F = 3;
        const int32_t fops[3] = {AISP_OP_EXPOSURE, AISP_OP_WB, AISP_OP_SHARPEN};   /* HOST array */
        float bpar[2 * 3 * AISP_PSTRIDE];
        float *d_stack, *d_gstack, *d_bpar, *d_bgp;
        void* d_bscr;
        size_t bscr = aisp_bwd_scratch_bytes(B * F, H, W);
        memset(bpar, 0, sizeof(bpar));
        for (int b = 0; b < B; ++b) {
            bpar[(b * F + 1) * AISP_PSTRIDE + 0] = bpar[(b * F + 1) * AISP_PSTRIDE + 1] = bpar[(b * F + 1) * AISP_PSTRIDE + 2] = 1.0f;
            bpar[(b * F + 2) * AISP_PSTRIDE] = 1.0f;
        }
        CK(cudaMalloc((void**)&d_stack, sizeof(float) * n * F));
        CK(cudaMalloc((void**)&d_gstack, sizeof(float) * n * F));
        CK(cudaMalloc((void**)&d_bpar, sizeof(bpar)));
        CK(cudaMalloc((void**)&d_bgp, sizeof(bpar)));
        CK(cudaMalloc(&d_bscr, bscr));
        CK(cudaMemcpy(d_bpar, bpar, sizeof(bpar), cudaMemcpyHostToDevice));
        float* h_stack = malloc(sizeof(float) * n * F);
        for (int i = 0; i < n * F; ++i) h_stack[i] = 1.0f;
        CK(cudaMemcpy(d_gstack, h_stack, sizeof(float) * n * F, cudaMemcpyHostToDevice));
        CK(aisp_bank_fwd(d_img, d_stack, d_bpar, fops, B, F, H, W, 0, NULL, NULL));
        CK(aisp_bank_bwd(d_img, d_gstack, d_bpar, fops, B, F, H, W, 0, NULL, d_bgp, d_bscr, bscr, NULL));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h_stack, d_stack, sizeof(float) * n * F, cudaMemcpyDeviceToHost));
        for (int b = 0; b < B; ++b)
            for (int f = 0; f < F; ++f)
                for (int i = 0; i < 3 * N; ++i) {
                    const float want = h_img[b * 3 * N + i];
                    const float got = h_stack[((size_t)(b * F + f) * 3) * N + i];
                    /* the reference's WB divides by (1e-5 + 0.27+0.67+0.06) */
                    if (fabsf(got - want) > 2e-5f) return fail("bank identity stack", got, want);
                }
        float bgp[2 * 3 * AISP_PSTRIDE];
        CK(cudaMemcpy(bgp, d_bgp, sizeof(bgp), cudaMemcpyDeviceToHost));
        for (int b = 0; b < B; ++b) {
            double s = 0;
            for (int i = 0; i < 3 * N; ++i) s += h_img[b * 3 * N + i];
            if (fabs(bgp[(b * F) * AISP_PSTRIDE] - s * 0.6931471805599453) > 1e-4 * s)
                return fail("bank exposure grad", bgp[(b * F) * AISP_PSTRIDE], s * 0.693147);
        }
        const int32_t two_nlm[2] = {AISP_OP_NLM, AISP_OP_NLM};
        if (aisp_bank_fwd(d_img, d_stack, d_bpar, two_nlm, B, 2, H, W, 0, NULL, NULL) != AISP_ERR_UNSUPPORTED)
            return fail("bank: two NLM slots must be refused", 0, 0);
    }

    /* 4. argument errors come back as negative status codes, not crashes */
    if (aisp_pointwise_fwd(NULL, d_out, d_par, d_ops, NULL, B, H, W, 1, 0, NULL) != AISP_ERR_NULL) return fail("null check", 0, 0);
    if (aisp_pointwise_fwd(d_img, d_img, d_par, d_ops, NULL, B, H, W, 1, 0, NULL) != AISP_ERR_UNSUPPORTED) return fail("alias check", 0, 0);
    printf("c_abi smoke: ok\n");
    return 0;
}
