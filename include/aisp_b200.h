/*
 * aisp_b200.h -- C ABI of the B200-native (sm_100a) differentiable ISP filter chain.
 *
 * This is the drop-in boundary for the hot path of OpenImagingLab/AdaptiveISP: the arithmetic
 * behind `Filter.process` / `Filter.forward` / `Filter.run` of isp/filters.py (with isp/denoise.py
 * and isp/sharpen.py) and the "apply the selected filter" step of agent.py:103-116,154.  The
 * reference has no FFI of its own (it is pure PyTorch); each entry point below cites the reference
 * code it replaces, and INTEGRATION.md shows the Python (ctypes) stub a maintainer binds it with.
 *
 * Conventions
 *   - all image tensors are device pointers to fp32, NCHW, contiguous, C == 3: [B,3,H,W];
 *   - `params` is a packed row per (sample, step): AISP_PSTRIDE floats, holding the filter's
 *     parameters exactly as the reference's `filter_param_regressor` returns them, flattened
 *     (Tone [B,8,1,1,1] -> 8 floats; Color [B,8,3,1,1] -> 24 floats knot-major; CCM 9 floats
 *     row-major; USM (sigma, amount); everything else 1 or 3 floats); unused tail is ignored;
 *   - `ops` holds one enum aisp_op per (sample, step): the policy-selected filter;
 *   - every buffer is caller-owned and only borrowed for the call; no hidden allocation, scratch
 *     is passed in; all work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = the
 *     legacy default stream) of the current device; functions are stateless and re-entrant;
 *   - return value: 0 on success, a negative aisp_status for argument errors, or a positive
 *     cudaError_t from the launch.  There is no CPU fallback anywhere.
 */
#ifndef AISP_B200_H
#define AISP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Filter op codes.  0..9 follow the order of cfg.filters (config.py:19-22). */
enum aisp_op {
    AISP_OP_NONE       = -1, /* no filter selected (all-zero one-hot row, agent.py:18-23,154):
                                the output image is exactly 0 and all gradients are 0            */
    AISP_OP_EXPOSURE   = 0,  /* ExposureFilter              isp/filters.py:215-224   1 param  */
    AISP_OP_GAMMA      = 1,  /* GammaFilter                 isp/filters.py:235-245   1        */
    AISP_OP_CCM        = 2,  /* CCMFilter                   isp/filters.py:694-708   9        */
    AISP_OP_SHARPEN    = 3,  /* SharpenFilter (3x3)         isp/filters.py:621-631   1        */
    AISP_OP_NLM        = 4,  /* DenoiseFilter (NLM gray)    isp/filters.py:571-586   1        */
    AISP_OP_TONE       = 5,  /* ToneFilter                  isp/filters.py:326-347   8        */
    AISP_OP_CONTRAST   = 6,  /* ContrastFilter              isp/filters.py:406-419   1        */
    AISP_OP_SATPLUS    = 7,  /* SaturationPlusFilter        isp/filters.py:536-560   1        */
    AISP_OP_WNB        = 8,  /* WNBFilter (desaturation)    isp/filters.py:427-437   1        */
    AISP_OP_WB         = 9,  /* ImprovedWhiteBalanceFilter  isp/filters.py:253-272   3        */
    AISP_OP_USM        = 10, /* SharpenUSMFilter (5x5)      isp/filters.py:597-608   2        */
    AISP_OP_COLOR      = 11, /* ColorFilter                 isp/filters.py:281-303   24       */
    AISP_OP_SHARPEN_V2 = 12, /* SharpenFilterV2             isp/filters.py:644-653   1        */
    AISP_OP_COUNT      = 13
};

#define AISP_PSTRIDE   24 /* floats per (sample, step) parameter row */
#define AISP_MAX_STEPS 8  /* longest fused per-sample sequence        */
#define AISP_ACC_STRIDE 32 /* floats per partial-sum row in the backward scratch */
#define AISP_MAX_CHAIN_BWD 6 /* longest per-sample sequence differentiated in one fused pass */

/* `clip_each` of the sequence entry points is a bit field: */
#define AISP_SEQ_CLIP   1 /* clip to [0,1] after every step (Filter.forward); 0: Filter.run semantics */
#define AISP_SEQ_STRICT 2 /* a non-per-pixel op anywhere in a sample's sequence poisons that sample's
                             output / gradients with NaN instead of being skipped or ending the sequence */

enum aisp_status {
    AISP_OK              = 0,
    AISP_ERR_NULL        = -1, /* a required pointer is NULL                          */
    AISP_ERR_SHAPE       = -2, /* B/H/W/S out of range                                */
    AISP_ERR_SCRATCH     = -3, /* scratch buffer too small                            */
    AISP_ERR_UNSUPPORTED = -4, /* combination not implemented (never silently wrong)  */
    AISP_ERR_ALIGN       = -5  /* pointer not 4-byte aligned                          */
};

/* Library ABI version (bumped on any signature change) and status text. */
int         aisp_version(void);
const char* aisp_status_string(int status);

/* Number of parameters of an op (aisp_op), or -1. */
int aisp_op_num_params(int op);

/*
 * Fused per-pixel pass, forward.  Replaces, for the per-pixel filters
 * (E, G, CCM, T, Ct, S+, BW, W, C), `lerp(img, process(img,p), 1)` [+ clip] of
 * isp/filters.py:115,125 / :138, applied `seq_len[b]` times in ONE pass over HBM with the
 * per-sample op sequence ops[b, 0..seq_len[b]) -- the heterogeneous "policy-selected filter
 * sequence" of agent.py:103-154 / yolov3/val_adaptiveisp.py:291-309 replay.
 *   img, out   [B,3,H,W]; out must not alias img
 *   params     [B,S,AISP_PSTRIDE]     ops [B,S] int32     seq_len [B] int32 or NULL (= S everywhere)
 *   clip_each  1: clip to [0,1] after every step (Filter.forward semantics)
 *              0: never clip (Filter.run semantics, isp/filters.py:128-139)
 * Samples whose FIRST op is a stencil op (SHARPEN, SHARPEN_V2, USM, NLM) are skipped entirely
 * (their `out` rows are left untouched) so that the three family entry points can be issued
 * back to back on a heterogeneous batch; a stencil op later in a sequence ends the sequence there.
 * With AISP_SEQ_STRICT set in `clip_each` such samples get an all-NaN `out` instead: callers that
 * own the whole sequence (fused chains, replay) then never read uninitialised memory.
 */
int aisp_pointwise_fwd(const float* img, float* out, const float* params, const int32_t* ops,
                       const int32_t* seq_len, int B, int H, int W, int S, int clip_each, void* stream);

/* Scratch bytes required by any *_bwd entry point for a [B,3,H,W] batch. */
size_t aisp_bwd_scratch_bytes(int B, int H, int W);

/*
 * Fused per-pixel pass, backward of ONE step per sample (what autograd does for
 * isp/filters.py:115,125 when train.py:341-342 calls backward()): the forward value is
 * recomputed in registers from `img`; parameter gradients are reduced warp-shuffle -> block ->
 * per-sample (fixed order, fp64 final combine: deterministic run to run).
 *   grad_out    [B,3,H,W]  upstream dL/dout
 *   clip        1 if the forward was Filter.forward (clip backward: pass iff 0 <= y <= 1)
 *   grad_params [B,AISP_PSTRIDE]   written (all PSTRIDE entries) for every non-skipped sample
 *   grad_img    [B,3,H,W] or NULL (NULL in training: train.py:255 makes img a leaf without grad)
 *   scratch     >= aisp_bwd_scratch_bytes(B,H,W) bytes of device memory
 * Samples whose op is a stencil op are skipped (see aisp_pointwise_fwd).
 */
int aisp_pointwise_bwd(const float* img, const float* grad_out, const float* params, const int32_t* ops,
                       int B, int H, int W, int clip, float* grad_params, float* grad_img,
                       void* scratch, size_t scratch_bytes, void* stream);

/*
 * Backward of a fused per-sample SEQUENCE of per-pixel filters (the forward is aisp_pointwise_fwd with
 * the same params / ops / seq_len / clip_each) in ONE pass over HBM: the chain is recomputed per
 * pixel (stage inputs parked in shared memory) and swept in reverse, so the traffic is that of a
 * single-step backward however many stages are fused.  S <= AISP_MAX_CHAIN_BWD (longer chains: split
 * them and pass grad_img on).  Every per-pixel op is differentiated, the 24-knot COLOR included
 * (sequences containing it take a slower one-CTA-per-sample path).
 *   grad_params [B,S,AISP_PSTRIDE]   rows of steps >= seq_len[b] (and of skipped samples) are zero
 *   grad_img    [B,3,H,W] or NULL    rows of skipped samples are left untouched
 *   scratch     >= aisp_bwd_scratch_bytes(B,H,W)
 * AISP_SEQ_STRICT: samples with a non-per-pixel op get NaN grad_params rows and a NaN grad_img.
 */
int aisp_pointwise_chain_bwd(const float* img, const float* grad_out, const float* params, const int32_t* ops,
                             const int32_t* seq_len, int B, int H, int W, int S, int clip_each,
                             float* grad_params, float* grad_img, void* scratch, size_t scratch_bytes,
                             void* stream);

/*
 * 3x3 sharpen (SHARPEN: adjust_sharpness isp/sharpen.py:105-142; SHARPEN_V2: sharpness :145-182)
 * and 5x5 unsharp mask (USM: unsharp_mask isp/sharpen.py:84-102 with reflect padding, per-sample
 * sigma/amount -- the reference loops over the batch in Python, :91-96).  Halo-tiled in shared
 * memory.  All three clip to [0,1] internally, so there is no clip flag.
 * Samples whose op is not one of the three are skipped.
 */
int aisp_sharpen_fwd(const float* img, float* out, const float* params, const int32_t* ops,
                     int B, int H, int W, void* stream);

/*
 * Backward of the above.  grad_params: SHARPEN/SHARPEN_V2 -> [0] = d/dfactor;
 * USM -> [0] = d/dsigma, [1] = d/damount.
 * If grad_img != NULL, `gy_scratch` ([B,3,H,W] floats) must be given: the masked upstream gradient
 * g*[0<=y<=1] is staged there and the transposed stencil is gathered from it in a second pass.
 */
int aisp_sharpen_bwd(const float* img, const float* grad_out, const float* params, const int32_t* ops,
                     int B, int H, int W, float* grad_params, float* grad_img, float* gy_scratch,
                     void* scratch, size_t scratch_bytes, void* stream);

/*
 * Non-local-means denoise, gray-distance variant with 11x11 search / 5x5 patch and circular
 * boundaries (DenoiseFilter.process isp/filters.py:582-586 -> NonLocalMeansGray
 * isp/denoise.py:93-119, BoxFilter :46-65, rgb_to_luminance :11-17).
 *   dout_dh  [B,3,H,W] or NULL.  When given, the kernel also writes d out / d h per pixel and
 *            channel (clamp mask folded in), the closed form of SURVEY.md §8a row A11, so that the
 *            backward w.r.t. h is a single dot product instead of a second 121-shift pass.
 *   wsum     [B,H,W] or NULL.  When given, the per-pixel sum of weights is stored; it is the one
 *            extra input aisp_nlm_bwd_img needs.
 * Samples whose op is not NLM are skipped.
 */
int aisp_nlm_fwd(const float* img, float* out, const float* params, const int32_t* ops,
                 int B, int H, int W, float* dout_dh, float* wsum, void* stream);

/*
 * The bare non-local-means MODULES of isp/denoise.py (11x11 search, 5x5 patch), i.e. without the
 * DenoiseFilter wrapper (no clip of the input, no lerp term):
 *   gray != 0: NonLocalMeansGray.forward(rgb, h) :93-119 -- distances on the luma of the clipped image
 *              (rgb_to_luminance clips, :14), averages of the image as given;
 *   gray == 0: NonLocalMeans.forward(rgb, h) :68-90 -- per-channel distances and weights [B,3,H,W]
 *              (three passes of the same kernel, one per channel).
 * params[b,0] = h; ops[b] must be AISP_OP_NLM.  dout_dh as in aisp_nlm_fwd (aisp_nlm_bwd is its backward).
 * No image gradient for the modules (AispError in the Python layer).
 */
int aisp_nlm_module_fwd(const float* img, float* out, const float* params, const int32_t* ops, int B, int H, int W,
                        float* dout_dh, int gray, void* stream);

/*
 * Shot / read noise of the synthetic-RAW model, isp/unprocess_np.py:131-181 (adjust_random_brightness +
 * add_read_and_shot_noise; the per-image noise levels of random_noise_levels_log / _linear are drawn on the host):
 *     out = gain[b] * img + sqrt(gain[b] * img * shot[b] + read[b]) * z,   z ~ N(0, 1)
 * img / out [B, n_per_image] (in place allowed); shot, read [B] device; gain [B] device or NULL (1).
 * z [B, n_per_image] device: the standard normals to use (the result is then a deterministic function of the
 * inputs); NULL: generated in the kernel (Philox4x32-10 keyed by `seed`, counter = offset + element group).
 */
int aisp_shot_read_noise(const float* img, const float* z, float* out, const float* shot, const float* read,
                         const float* gain, int B, long long n_per_image, unsigned long long seed,
                         unsigned long long offset, void* stream);

/*
 * NonLocalMeansParam.forward(rgb) of isp/denoise.py:122-157: reflect-padded search window `window` (odd,
 * window / 2 < min(H, W)), patch box as large as the search window (:145-146), ONE scalar h for the whole
 * batch (`h`: device pointer to one float, the module's nn.Parameter).  luma [B,H,W] = rgb_to_luminance(rgb)
 * (isp/denoise.py:11-17).  dout_dh [B,3,H,W] or NULL: d out / d h per element; grad_h = sum(grad_out * dout_dh).
 * Instantiated nowhere in the reference: a plain one-pixel-per-thread kernel, not a tuned one.
 */
int aisp_nlm_param_fwd(const float* rgb, const float* luma, float* out, float* dout_dh, const float* h, int B, int H,
                       int W, int window, void* stream);

/* Backward of NLM w.r.t. h:  grad_params[b,0] = sum_{c,y,x} grad_out * dout_dh. */
int aisp_nlm_bwd(const float* grad_out, const float* dout_dh, const int32_t* ops, int B, int H, int W,
                 float* grad_params, void* scratch, size_t scratch_bytes, void* stream);

/*
 * Backward of NLM w.r.t. the image (autograd through isp/denoise.py:93-119 and the leading clip
 * of isp/filters.py:583): both the shifted-RGB gather and the dependence of every weight on the
 * luma patches.  Needs the forward's `out` and `wsum`.  Gather-only and deterministic; written for
 * correctness, not speed -- the reference's training never requests it (train.py:255).
 */
int aisp_nlm_bwd_img(const float* img, const float* out, const float* wsum, const float* grad_out,
                     const float* params, const int32_t* ops, int B, int H, int W, float* grad_img,
                     void* stream);

/*
 * Block-mean down-sampling [B,3,H,W] -> [B,3,out_h,out_w] in one read-only pass; requires
 * H % out_h == 0 and W % out_w == 0 (then identical to nn.AdaptiveAvgPool2d((out_h,out_w)),
 * agent.py:97 / value.py:63; otherwise AISP_ERR_UNSUPPORTED).  The per-image mean and the NaN/Inf
 * guard of train.py:288-290,374 are functions of this small image, so the five or six
 * full-resolution reads the reference spends on them collapse into this one.
 */
int aisp_block_mean(const float* img, float* down, int B, int H, int W, int out_h, int out_w, void* stream);

/*
 * The critic's statistics of the pooled image (value.py:64-75): stats[b] = (mean luminance, unbiased
 * luminance variance, mean saturation) of down[b] ([B,3,h,w], e.g. the block means above), one launch.
 * Forward only: the caller keeps the PyTorch statement where a gradient must flow through them.
 */
int aisp_value_stats(const float* down, int B, int h, int w, float* stats, void* stream);

/*
 * Apply the selected filter of each sample: the B200 form of agent.py:103-116,154, where the
 * reference runs all 10 filters on the whole batch, stacks [B,10,3,H,W] and keeps one of ten.
 * Issues the three family kernels back to back on `stream` (no host sync; graph-capturable);
 * every sample is processed by exactly one of them.  clip as in aisp_pointwise_fwd.
 * nlm_dout_dh / nlm_wsum: stashes for the NLM samples (see aisp_nlm_fwd), NULL if not needed.
 */
int aisp_select_apply_fwd(const float* img, float* out, const float* params, const int32_t* ops,
                          int B, int H, int W, int clip, float* nlm_dout_dh, float* nlm_wsum, void* stream);

/* Backward of the above.  `out` and `nlm_wsum` are only read when grad_img != NULL. */
int aisp_select_apply_bwd(const float* img, const float* out, const float* grad_out, const float* params,
                          const int32_t* ops, int B, int H, int W, int clip, const float* nlm_dout_dh,
                          const float* nlm_wsum, float* grad_params, float* grad_img, float* gy_scratch,
                          void* scratch, size_t scratch_bytes, void* stream);

/*
 * Sequence launch set: the general forward of the path.  Every sample b runs its own op sequence
 * ops[b, 0..seq_len[b]) that may hold per-pixel steps and AT MOST ONE stencil step (SHARPEN, SHARPEN_V2,
 * USM or NLM) anywhere in it -- the fixed chain of isp/filters.py:753-815 (E -> G -> WB -> CCM -> Shr),
 * the policy-selected single step of agent.py:103-154 (S == 1) and the replayed pipelines of
 * yolov3/val_adaptiveisp.py:291-327.  Three kernels are issued back to back (per-pixel, sharpen family,
 * NLM); each sample is processed by exactly one of them, per-pixel steps before the stencil step run on
 * the staged tile, steps after it on the outputs in registers: 24 B/px for the whole sequence.  A second
 * stencil step ends the sequence there (a host-side planner splits such pipelines into several calls).
 *   hr_img / hr_out  [B,3,hr_H,hr_W] or NULL: the high-resolution twin of the batch gets the SAME
 *                    sequences and parameters (isp/filters.py:116-122, agent.py:155-157, train.py:541) as
 *                    extra tiles of the same per-pixel / sharpen launches (NLM: a second launch)
 *   down             [B,3,down_h,down_w] or NULL: block means of `out` (== nn.AdaptiveAvgPool2d for evenly
 *                    dividing sizes: what agent.py:97 and value.py:63 compute next from the retouched
 *                    image), emitted from the kernels' store path where the pooling blocks tile the CTA's
 *                    work (512x512 -> 64x64 does) and completed by one masked pass for the rest (NLM samples)
 *   nlm_dout_dh / nlm_wsum   training stashes of the NLM samples (see aisp_nlm_fwd); S == 1 only
 * clip_each: AISP_SEQ_CLIP or 0.
 */
int aisp_sequence_fwd(const float* img, float* out, const float* params, const int32_t* ops, const int32_t* seq_len,
                      int B, int H, int W, int S, int clip_each, const float* hr_img, float* hr_out, int hr_H, int hr_W,
                      float* down, int down_h, int down_w, float* nlm_dout_dh, float* nlm_wsum, void* stream);

/*
 * aisp_select_apply_bwd when a gradient ALSO reaches the block means of the output (aisp_sequence_fwd's
 * `down`: the critic pools the retouched image, value.py:63, so train.py:341-342 differentiates through the
 * pooled image too).  grad_down [B,3,down_h,down_w] is added, divided by the block area, to grad_out inside
 * the kernels' loads: no up-sampled gradient image is materialised.  Parameter gradients only (grad_img
 * must be NULL); W and the pooling block sides must be powers of two (512 -> 64 is), else
 * AISP_ERR_UNSUPPORTED and the caller adds the up-sampled gradient itself.
 */
int aisp_select_apply_bwd_pooled(const float* img, const float* out, const float* grad_out, const float* grad_down,
                                 int down_h, int down_w, const float* params, const int32_t* ops, int B, int H, int W,
                                 int clip, const float* nlm_dout_dh, const float* nlm_wsum, float* grad_params,
                                 float* grad_img, float* gy_scratch, void* scratch, size_t scratch_bytes, void* stream);

/*
 * Device-side filter selection + agent-state update in one launch, no host round trip:
 * pdf_sample / argmax / forced id (agent.py:12-16,126-149), one_hot (agent.py:18-23), the gather
 * of the selected filter's parameter row (the B200 form of agent.py:154: pick the row before the
 * filter runs instead of one of ten images after) and the new state (agent.py:234-259).
 *   pdf        [B,F]   selection probabilities (after the exploration mix and renormalisation)
 *   noise      [B]     uniform noise z[:,0] (AISP_SELECT_SAMPLE only, else may be NULL)
 *   states     [B,S]   S = 3 + F: (reward, stopped, step, usage[F])            (util.py:15-18)
 *   packed_all [B,F,AISP_PSTRIDE]  every filter's packed parameter row
 *   op_table   [F]     int32, filter index -> aisp_op (device)
 * outputs: sel [B] int64 (-1 possible: pdf_sample with noise == 0), one_hot [B,F] int64,
 *   ops [B] int32 (AISP_OP_NONE for sel == -1), rows [B,AISP_PSTRIDE] (zeros for sel == -1),
 *   new_states [B,S], penalties [B,2] = (usage penalty, early-stop penalty).
 * Sums are sequential fp32 in index order, like torch.cumsum / torch.sum over a 10-entry row.
 */
enum aisp_select_mode { AISP_SELECT_SAMPLE = 0, AISP_SELECT_ARGMAX = 1, AISP_SELECT_FORCED = 2 };

int aisp_select(const float* pdf, const float* noise, int mode, int forced_id, const float* states,
                const float* packed_all, const int32_t* op_table, int B, int F, int S, float test_steps,
                float early_stop_penalty, int64_t* sel, int64_t* one_hot, int32_t* ops, float* rows,
                float* new_states, float* penalties, void* stream);

/* Backward of the row gather: grad_packed_all [B,F,AISP_PSTRIDE] = grad_rows scattered to the
 * selected slot, exact zeros elsewhere (what torch.gather's backward gives the reference). */
int aisp_select_bwd(const float* grad_rows, const int64_t* sel, int B, int F, float* grad_packed_all,
                    void* stream);

/*
 * Feature -> parameter regressors of every filter of a bank in ONE launch (forward) and one (backward):
 * `Filter.filter_param_regressor` of isp/filters.py:215-708 (tanh_range / exp / sigmoid, the white-balance
 * red-feature mask and luminance normalisation :253-272), producing the packed rows the image kernels read.
 *   raw         [B,Ntot]  the fc_filter outputs of the F filters side by side; filter f's n_f features
 *                         start at column offsets[f]
 *   filter_ops  [F] int32 (device)   offsets [F] int32 (device)
 *   cfg_ranges  HOST array of 15 floats, (lo, span) pairs formed in double precision as the reference
 *               forms r - l: exposure, log-gamma, tone curve, colour curve, colour-curve shift
 *               (atanh term of tanh_range's `initial`), USM, sharpen, CCM
 *   packed      [B,F,AISP_PSTRIDE]   (unused tail of a row: zeros)
 * Backward: grad_raw [B,Ntot] = J^T grad_packed (columns of no filter are left untouched).
 */
int aisp_regress_fwd(const float* raw, const int32_t* filter_ops, const int32_t* offsets, int B, int F, int Ntot,
                     const float* cfg_ranges, float* packed, void* stream);
int aisp_regress_bwd(const float* raw, const float* grad_packed, const int32_t* filter_ops, const int32_t* offsets,
                     int B, int F, int Ntot, const float* cfg_ranges, float* grad_raw, void* stream);

/*
 * Filter bank: apply F filters to the SAME batch and keep every result -- the stack of
 * agent.py:103-107 (`filtered_images.append(filter(...))` for every cfg.filter, then
 * torch.stack(dim=1)), which is also BASELINE.json configs[1] ("all 10 filters fwd+bwd").
 *   img     [B,3,H,W]                      out / grad_out  [B,F,3,H,W]
 *   params  [B,F,AISP_PSTRIDE]             filter_ops      HOST array of F op codes (F <= 16,
 *   grad_params [B,F,AISP_PSTRIDE]                         at most ONE AISP_OP_NLM, no AISP_OP_NONE)
 *   nlm_dout_dh [B,3,H,W] device stash for the NLM slot, or NULL (forward: no stash is written;
 *               backward: the NLM slot's grad_params row is left untouched)
 * Every kernel family is launched once, over its own slots only ("virtual samples" v = b*F + f,
 * image index v / F).  The per-pixel slots of an image chunk are all applied by the same CTA (the
 * chunk is read once: into registers forward, into a shared-memory cache backward); the stencil
 * slots share the image through L2.  DRAM traffic is ~(1 + F) planes forward instead of 2F.
 * Results are bit-identical to F separate aisp_select_apply_* calls.
 * Parameter gradients only (the reference never differentiates the stack w.r.t. the input batch
 * without going through the selection, train.py:255,341-342).  scratch_bytes >=
 * aisp_bwd_scratch_bytes(B*F, H, W);  B*F <= 65535.
 */
int aisp_bank_fwd(const float* img, float* out, const float* params, const int32_t* filter_ops, int B, int F,
                  int H, int W, int clip, float* nlm_dout_dh, void* stream);

int aisp_bank_bwd(const float* img, const float* grad_out, const float* params, const int32_t* filter_ops, int B,
                  int F, int H, int W, int clip, const float* nlm_dout_dh, float* grad_params, void* scratch,
                  size_t scratch_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AISP_B200_H */
