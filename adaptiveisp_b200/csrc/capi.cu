// extern "C" boundary of libaisp_b200.so: argument validation + dispatch to the kernel launchers.
// Plain pointers and sizes only -- no torch types.  See include/aisp_b200.h for the contract.
#include "aisp_common.cuh"

namespace aisp {
cudaError_t launch_pointwise_fwd(const float*, float*, const float*, const int32_t*, const int32_t*, int, int, int, int,
                                 int, BankMap, cudaStream_t);
cudaError_t launch_pointwise_bank_fwd(const float*, float*, const float*, int, int, int, int, BankMap, cudaStream_t);
cudaError_t launch_pointwise_bank_bwd(const float*, const float*, const float*, int, int, int, int, float*, float*, BankMap,
                                      bool, cudaStream_t);
cudaError_t launch_pointwise_bwd(const float*, const float*, const float*, const int32_t*, int, int, int, int, float*,
                                 float*, float*, BankMap, PooledGrad, cudaStream_t);
cudaError_t launch_sharpen_fwd(const float*, float*, const float*, const int32_t*, int, int, int, BankMap, cudaStream_t);
cudaError_t launch_sharpen_bwd(const float*, const float*, const float*, const int32_t*, int, int, int, float*, float*,
                               float*, float*, BankMap, PooledGrad, cudaStream_t);
cudaError_t launch_nlm_fwd(const float*, float*, const float*, const int32_t*, int, int, int, float*, float*, BankMap,
                           cudaStream_t);
cudaError_t launch_nlm_bwd_img(const float*, const float*, const float*, const float*, const float*, const int32_t*, int,
                               int, int, float*, cudaStream_t);
cudaError_t launch_nlm_bwd(const float*, const float*, const float*, const int32_t*, int, int, int, float*, float*, BankMap,
                           PooledGrad, cudaStream_t);
cudaError_t launch_block_mean(const float*, float*, int, int, int, int, int, cudaStream_t);
cudaError_t launch_pointwise_chain_bwd(const float*, const float*, const float*, const int32_t*, const int32_t*, int, int,
                                       int, int, int, float*, float*, float*, cudaStream_t);
cudaError_t launch_select(const float*, const float*, int, int, const float*, const float*, const int32_t*, int, int, int,
                          float, float, long long*, long long*, int32_t*, float*, float*, float*, cudaStream_t);
cudaError_t launch_select_bwd(const float*, const long long*, int, int, float*, cudaStream_t);
cudaError_t launch_pointwise_seq_fwd(const float*, float*, const float*, const int32_t*, const int32_t*, int, int, int, int,
                                     int, const float*, float*, int, int, float*, int, int, cudaStream_t);
cudaError_t launch_sharpen_seq_fwd(const float*, float*, const float*, const int32_t*, const int32_t*, int, int, int, int,
                                   int, const float*, float*, int, int, float*, int, int, cudaStream_t);
cudaError_t launch_nlm_seq_fwd(const float*, float*, const float*, const int32_t*, const int32_t*, int, int, int, int, int,
                               cudaStream_t);
cudaError_t launch_block_mean_masked(const float*, float*, int, int, int, int, int, const int32_t*, const int32_t*, int, int,
                                     cudaStream_t);
cudaError_t launch_regress(const float*, const int32_t*, const int32_t*, int, int, int, const float*, float*, const float*,
                           float*, cudaStream_t);
cudaError_t launch_value_stats(const float*, int, int, float*, cudaStream_t);
cudaError_t launch_nlm_module_fwd(const float*, float*, const float*, const int32_t*, int, int, int, float*, int, cudaStream_t);
cudaError_t launch_nlm_param_fwd(const float*, const float*, float*, float*, const float*, int, int, int, int, cudaStream_t);
cudaError_t launch_shot_read_noise(const float*, const float*, float*, const float*, const float*, const float*, int,
                                   long long, unsigned long long, unsigned long long, cudaStream_t);
bool pointwise_can_emit(int, int, int, int);
bool sharpen_can_emit(int, int, int, int);
int chain_bwd_max_steps();
int pointwise_rows(int H, int W);
int sharpen_rows(int H, int W);
}  // namespace aisp

using namespace aisp;

namespace {
inline bool al4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }
inline int shape_ok(int B, int H, int W) {
    // int32 pixel indexing inside a plane; gridDim.y/z limit on the batch
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    if ((long long)H * W > (1LL << 30)) return 0;
    if (B > 65535) return 0;
    return 1;
}
}  // namespace

extern "C" {

int aisp_version(void) { return 3; }

const char* aisp_status_string(int s) {
    switch (s) {
    case AISP_OK: return "ok";
    case AISP_ERR_NULL: return "required pointer is NULL";
    case AISP_ERR_SHAPE: return "shape out of range";
    case AISP_ERR_SCRATCH: return "scratch buffer too small";
    case AISP_ERR_UNSUPPORTED: return "unsupported combination";
    case AISP_ERR_ALIGN: return "pointer not 4-byte aligned";
    default: return s > 0 ? cudaGetErrorString((cudaError_t)s) : "unknown status";
    }
}

int aisp_op_num_params(int op) {
    static const int n[AISP_OP_COUNT] = {1, 1, 9, 1, 1, 8, 1, 1, 1, 3, 2, 24, 1};
    return (op >= 0 && op < AISP_OP_COUNT) ? n[op] : -1;
}

size_t aisp_bwd_scratch_bytes(int B, int H, int W) {
    if (!shape_ok(B, H, W)) return 0;
    const int rows = pointwise_rows(H, W) > sharpen_rows(H, W) ? pointwise_rows(H, W) : sharpen_rows(H, W);
    // x AISP_MAX_CHAIN_BWD: the fused multi-step backward keeps one row per (chunk, stage)
    return (size_t)B * rows * AISP_ACC_STRIDE * sizeof(float) * AISP_MAX_CHAIN_BWD;
}

int aisp_pointwise_fwd(const float* img, float* out, const float* params, const int32_t* ops, const int32_t* seq_len,
                       int B, int H, int W, int S, int clip_each, void* stream) {
    if (!img || !out || !params || !ops) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W) || S < 1 || S > AISP_MAX_STEPS) return AISP_ERR_SHAPE;
    if (!al4(img) || !al4(out)) return AISP_ERR_ALIGN;
    if (img == out) return AISP_ERR_UNSUPPORTED;
    return (int)launch_pointwise_fwd(img, out, params, ops, seq_len, B, H, W, S, clip_each & 3, plain_batch(),
                                     (cudaStream_t)stream);
}

int aisp_pointwise_bwd(const float* img, const float* grad_out, const float* params, const int32_t* ops, int B, int H,
                       int W, int clip, float* grad_params, float* grad_img, void* scratch, size_t scratch_bytes,
                       void* stream) {
    if (!img || !grad_out || !params || !ops || !grad_params || !scratch) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W)) return AISP_ERR_SHAPE;
    if (scratch_bytes < aisp_bwd_scratch_bytes(B, H, W)) return AISP_ERR_SCRATCH;
    if (!al4(img) || !al4(grad_out) || !al4(grad_img)) return AISP_ERR_ALIGN;
    return (int)launch_pointwise_bwd(img, grad_out, params, ops, B, H, W, clip ? 1 : 0, grad_params, grad_img,
                                     (float*)scratch, plain_batch(), no_pooled_grad(), (cudaStream_t)stream);
}

int aisp_pointwise_chain_bwd(const float* img, const float* grad_out, const float* params, const int32_t* ops,
                             const int32_t* seq_len, int B, int H, int W, int S, int clip_each, float* grad_params,
                             float* grad_img, void* scratch, size_t scratch_bytes, void* stream) {
    if (!img || !grad_out || !params || !ops || !grad_params || !scratch) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W) || S < 1) return AISP_ERR_SHAPE;
    if (S > AISP_MAX_CHAIN_BWD || S > chain_bwd_max_steps()) return AISP_ERR_UNSUPPORTED;
    if (scratch_bytes < aisp_bwd_scratch_bytes(B, H, W)) return AISP_ERR_SCRATCH;
    if (!al4(img) || !al4(grad_out) || !al4(grad_img)) return AISP_ERR_ALIGN;
    return (int)launch_pointwise_chain_bwd(img, grad_out, params, ops, seq_len, B, H, W, S, clip_each & 3,
                                           grad_params, grad_img, (float*)scratch, (cudaStream_t)stream);
}

int aisp_sharpen_fwd(const float* img, float* out, const float* params, const int32_t* ops, int B, int H, int W,
                     void* stream) {
    if (!img || !out || !params || !ops) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W)) return AISP_ERR_SHAPE;
    if (img == out) return AISP_ERR_UNSUPPORTED;
    return (int)launch_sharpen_fwd(img, out, params, ops, B, H, W, plain_batch(), (cudaStream_t)stream);
}

int aisp_sharpen_bwd(const float* img, const float* grad_out, const float* params, const int32_t* ops, int B, int H,
                     int W, float* grad_params, float* grad_img, float* gy_scratch, void* scratch,
                     size_t scratch_bytes, void* stream) {
    if (!img || !grad_out || !params || !ops || !grad_params || !scratch) return AISP_ERR_NULL;
    if (grad_img && !gy_scratch) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W)) return AISP_ERR_SHAPE;
    if (scratch_bytes < aisp_bwd_scratch_bytes(B, H, W)) return AISP_ERR_SCRATCH;
    return (int)launch_sharpen_bwd(img, grad_out, params, ops, B, H, W, grad_params, grad_img, gy_scratch,
                                   (float*)scratch, plain_batch(), no_pooled_grad(), (cudaStream_t)stream);
}

int aisp_nlm_fwd(const float* img, float* out, const float* params, const int32_t* ops, int B, int H, int W,
                 float* dout_dh, float* wsum, void* stream) {
    if (!img || !out || !params || !ops) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W)) return AISP_ERR_SHAPE;
    if (img == out) return AISP_ERR_UNSUPPORTED;
    return (int)launch_nlm_fwd(img, out, params, ops, B, H, W, dout_dh, wsum, plain_batch(), (cudaStream_t)stream);
}

int aisp_nlm_module_fwd(const float* img, float* out, const float* params, const int32_t* ops, int B, int H, int W,
                        float* dout_dh, int gray, void* stream) {
    if (!img || !out || !params || !ops) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W)) return AISP_ERR_SHAPE;
    if (img == out) return AISP_ERR_UNSUPPORTED;
    return (int)launch_nlm_module_fwd(img, out, params, ops, B, H, W, dout_dh, gray ? 1 : 0, (cudaStream_t)stream);
}

int aisp_nlm_param_fwd(const float* rgb, const float* luma, float* out, float* dout_dh, const float* h, int B, int H,
                       int W, int window, void* stream) {
    if (!rgb || !luma || !out || !h) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W)) return AISP_ERR_SHAPE;
    // reflect padding needs pad < size (as F.pad does); windows are odd
    if (window < 1 || (window & 1) == 0 || window / 2 >= H || window / 2 >= W) return AISP_ERR_SHAPE;
    if (rgb == out) return AISP_ERR_UNSUPPORTED;
    return (int)launch_nlm_param_fwd(rgb, luma, out, dout_dh, h, B, H, W, window, (cudaStream_t)stream);
}

int aisp_nlm_bwd(const float* grad_out, const float* dout_dh, const int32_t* ops, int B, int H, int W,
                 float* grad_params, void* scratch, size_t scratch_bytes, void* stream) {
    if (!grad_out || !dout_dh || !ops || !grad_params || !scratch) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W)) return AISP_ERR_SHAPE;
    if (scratch_bytes < aisp_bwd_scratch_bytes(B, H, W)) return AISP_ERR_SCRATCH;
    return (int)launch_nlm_bwd(grad_out, dout_dh, nullptr, ops, B, H, W, grad_params, (float*)scratch, plain_batch(),
                               no_pooled_grad(), (cudaStream_t)stream);
}

int aisp_nlm_bwd_img(const float* img, const float* out, const float* wsum, const float* grad_out,
                     const float* params, const int32_t* ops, int B, int H, int W, float* grad_img, void* stream) {
    if (!img || !out || !wsum || !grad_out || !params || !ops || !grad_img) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W)) return AISP_ERR_SHAPE;
    return (int)launch_nlm_bwd_img(img, out, wsum, grad_out, params, ops, B, H, W, grad_img, (cudaStream_t)stream);
}

int aisp_block_mean(const float* img, float* down, int B, int H, int W, int out_h, int out_w, void* stream) {
    if (!img || !down) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W) || out_h <= 0 || out_w <= 0 || (long long)B * 3 > 65535) return AISP_ERR_SHAPE;
    if (H % out_h != 0 || W % out_w != 0) return AISP_ERR_UNSUPPORTED;  // adaptive pooling with uneven windows
    return (int)launch_block_mean(img, down, B, H, W, out_h, out_w, (cudaStream_t)stream);
}

int aisp_shot_read_noise(const float* img, const float* z, float* out, const float* shot, const float* read,
                         const float* gain, int B, long long n_per_image, unsigned long long seed,
                         unsigned long long offset, void* stream) {
    if (!img || !out || !shot || !read) return AISP_ERR_NULL;
    if (B <= 0 || B > 65535 || n_per_image <= 0) return AISP_ERR_SHAPE;
    return (int)launch_shot_read_noise(img, z, out, shot, read, gain, B, n_per_image, seed, offset, (cudaStream_t)stream);
}

int aisp_value_stats(const float* down, int B, int h, int w, float* stats, void* stream) {
    if (!down || !stats) return AISP_ERR_NULL;
    if (B <= 0 || h <= 0 || w <= 0 || (long long)h * w > (1LL << 28)) return AISP_ERR_SHAPE;
    return (int)launch_value_stats(down, B, h * w, stats, (cudaStream_t)stream);
}

int aisp_select_apply_fwd(const float* img, float* out, const float* params, const int32_t* ops, int B, int H, int W,
                          int clip, float* nlm_dout_dh, float* nlm_wsum, void* stream) {
    int e = aisp_pointwise_fwd(img, out, params, ops, nullptr, B, H, W, 1, clip, stream);
    if (e) return e;
    e = aisp_sharpen_fwd(img, out, params, ops, B, H, W, stream);
    if (e) return e;
    return aisp_nlm_fwd(img, out, params, ops, B, H, W, nlm_dout_dh, nlm_wsum, stream);
}

int aisp_select_apply_bwd(const float* img, const float* out, const float* grad_out, const float* params,
                          const int32_t* ops, int B, int H, int W, int clip, const float* nlm_dout_dh,
                          const float* nlm_wsum, float* grad_params, float* grad_img, float* gy_scratch, void* scratch,
                          size_t scratch_bytes, void* stream) {
    // the NLM samples of a heterogeneous batch need their stashes: d out/d h for grad_params,
    // the weight sums (and the forward output) for grad_img
    // (nlm_dout_dh == NULL: the caller does not need parameter gradients of the NLM samples; their
    //  grad_params rows are left untouched)
    if (grad_img && (!nlm_wsum || !out)) return AISP_ERR_NULL;
    int e = aisp_pointwise_bwd(img, grad_out, params, ops, B, H, W, clip, grad_params, grad_img, scratch,
                               scratch_bytes, stream);
    if (e) return e;
    e = aisp_sharpen_bwd(img, grad_out, params, ops, B, H, W, grad_params, grad_img, gy_scratch, scratch,
                         scratch_bytes, stream);
    if (e) return e;
    if (nlm_dout_dh) e = aisp_nlm_bwd(grad_out, nlm_dout_dh, ops, B, H, W, grad_params, scratch, scratch_bytes, stream);
    if (e) return e;
    if (grad_img) e = aisp_nlm_bwd_img(img, out, nlm_wsum, grad_out, params, ops, B, H, W, grad_img, stream);
    return e;
}

// Backward of aisp_sequence_fwd with S == 1 (the select-apply step) when a gradient also reaches the block
// means: it is added to the upstream gradient inside the kernels' loads (PooledGrad).
static inline int ilog2_exact(int v) {
    if (v <= 0 || (v & (v - 1))) return -1;
    int k = 0;
    while ((1 << k) < v) ++k;
    return k;
}

int aisp_select_apply_bwd_pooled(const float* img, const float* out, const float* grad_out, const float* grad_down,
                                 int down_h, int down_w, const float* params, const int32_t* ops, int B, int H, int W,
                                 int clip, const float* nlm_dout_dh, const float* nlm_wsum, float* grad_params,
                                 float* grad_img, float* gy_scratch, void* scratch, size_t scratch_bytes, void* stream) {
    if (!img || !grad_out || !grad_down || !params || !ops || !grad_params || !scratch) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W) || down_h <= 0 || down_w <= 0 || H % down_h || W % down_w) return AISP_ERR_SHAPE;
    if (scratch_bytes < aisp_bwd_scratch_bytes(B, H, W)) return AISP_ERR_SCRATCH;
    // the NLM image-gradient kernel (rare path) has no pooled input: parameter gradients only here
    (void)out; (void)nlm_wsum; (void)gy_scratch;
    if (grad_img) return AISP_ERR_UNSUPPORTED;
    PooledGrad pg;
    pg.g = grad_down;
    pg.ws = ilog2_exact(W);
    pg.bhs = ilog2_exact(H / down_h);
    pg.bws = ilog2_exact(W / down_w);
    pg.oh = down_h;
    pg.ow = down_w;
    pg.inv_area = 1.0f / (float)((H / down_h) * (W / down_w));
    // (a pooling block must not split a 4-pixel vector: block width >= 4 or the scalar path)
    if (pg.ws < 0 || pg.bhs < 0 || pg.bws < 0 || ((W / down_w) < 4 && (W % 4) == 0)) return AISP_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = launch_pointwise_bwd(img, grad_out, params, ops, B, H, W, clip ? 1 : 0, grad_params, nullptr,
                                         (float*)scratch, plain_batch(), pg, st);
    if (e != cudaSuccess) return (int)e;
    e = launch_sharpen_bwd(img, grad_out, params, ops, B, H, W, grad_params, nullptr, nullptr, (float*)scratch,
                           plain_batch(), pg, st);
    if (e != cudaSuccess) return (int)e;
    if (nlm_dout_dh)
        e = launch_nlm_bwd(grad_out, nlm_dout_dh, nullptr, ops, B, H, W, grad_params, (float*)scratch, plain_batch(), pg, st);
    return (int)e;
}

int aisp_select(const float* pdf, const float* noise, int mode, int forced_id, const float* states,
                const float* packed_all, const int32_t* op_table, int B, int F, int S, float test_steps,
                float early_stop_penalty, int64_t* sel, int64_t* one_hot, int32_t* ops, float* rows, float* new_states,
                float* penalties, void* stream) {
    if (!pdf || !states || !packed_all || !op_table || !sel || !one_hot || !ops || !rows || !new_states || !penalties)
        return AISP_ERR_NULL;
    if (mode == AISP_SELECT_SAMPLE && !noise) return AISP_ERR_NULL;
    if (B <= 0 || F <= 0 || S != 3 + F || (long long)B * F * AISP_PSTRIDE > (1LL << 30)) return AISP_ERR_SHAPE;
    if (mode < AISP_SELECT_SAMPLE || mode > AISP_SELECT_FORCED) return AISP_ERR_UNSUPPORTED;
    if (mode == AISP_SELECT_FORCED && (forced_id < 0 || forced_id >= F)) return AISP_ERR_SHAPE;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
    return (int)launch_select(pdf, noise, mode, forced_id, states, packed_all, op_table, B, F, S, test_steps,
                              early_stop_penalty, (long long*)sel, (long long*)one_hot, ops, rows, new_states,
                              penalties, (cudaStream_t)stream);
}

int aisp_select_bwd(const float* grad_rows, const int64_t* sel, int B, int F, float* grad_packed_all, void* stream) {
    if (!grad_rows || !sel || !grad_packed_all) return AISP_ERR_NULL;
    if (B <= 0 || F <= 0 || (long long)B * F * AISP_PSTRIDE > (1LL << 30)) return AISP_ERR_SHAPE;
    return (int)launch_select_bwd(grad_rows, (const long long*)sel, B, F, grad_packed_all, (cudaStream_t)stream);
}

// ---- sequence launch set: per-sample sequences with at most one stencil step, optional high-resolution
// twin, optional block means of the output (see the header).  Every sample is processed by exactly one
// of the three family kernels; no host knowledge of the ops is needed.
int aisp_sequence_fwd(const float* img, float* out, const float* params, const int32_t* ops, const int32_t* seq_len,
                      int B, int H, int W, int S, int clip_each, const float* hr_img, float* hr_out, int hr_H, int hr_W,
                      float* down, int down_h, int down_w, float* nlm_dout_dh, float* nlm_wsum, void* stream) {
    if (!img || !out || !params || !ops) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W) || S < 1 || S > AISP_MAX_STEPS) return AISP_ERR_SHAPE;
    if (!al4(img) || !al4(out)) return AISP_ERR_ALIGN;
    if (img == out) return AISP_ERR_UNSUPPORTED;
    if ((hr_img == nullptr) != (hr_out == nullptr)) return AISP_ERR_NULL;
    if (hr_img && (!shape_ok(B, hr_H, hr_W) || hr_img == hr_out)) return AISP_ERR_SHAPE;
    if (down && (down_h <= 0 || down_w <= 0 || H % down_h || W % down_w || (long long)B * 3 > 65535)) return AISP_ERR_UNSUPPORTED;
    if ((nlm_dout_dh || nlm_wsum) && S != 1) return AISP_ERR_UNSUPPORTED;   // closed-form d/dh needs NLM to be the whole step
    cudaStream_t st = (cudaStream_t)stream;
    const int flags = (clip_each & AISP_SEQ_CLIP);
    // which families can emit the block means from their store path; the rest is completed by one masked pass
    const bool pw_emit = down && pointwise_can_emit(H, W, down_h, down_w) && (((uintptr_t)img | (uintptr_t)out) & 15u) == 0;
    const bool sh_emit = down && sharpen_can_emit(H, W, down_h, down_w);
    cudaError_t e = launch_pointwise_seq_fwd(img, out, params, ops, seq_len, B, H, W, S, flags, hr_img, hr_out, hr_H, hr_W,
                                             pw_emit ? down : nullptr, down_h, down_w, st);
    if (e != cudaSuccess) return (int)e;
    e = launch_sharpen_seq_fwd(img, out, params, ops, seq_len, B, H, W, S, flags, hr_img, hr_out, hr_H, hr_W,
                               sh_emit ? down : nullptr, down_h, down_w, st);
    if (e != cudaSuccess) return (int)e;
    if (S == 1) {   // the select-apply case keeps the plain NLM kernel and its training stashes
        e = launch_nlm_fwd(img, out, params, ops, B, H, W, nlm_dout_dh, nlm_wsum, plain_batch(), st);
        if (e == cudaSuccess && hr_img)
            e = launch_nlm_fwd(hr_img, hr_out, params, ops, B, hr_H, hr_W, nullptr, nullptr, plain_batch(), st);
    } else {
        e = launch_nlm_seq_fwd(img, out, params, ops, seq_len, B, H, W, S, flags, st);
        if (e == cudaSuccess && hr_img)
            e = launch_nlm_seq_fwd(hr_img, hr_out, params, ops, seq_len, B, hr_H, hr_W, S, flags, st);
    }
    if (e != cudaSuccess) return (int)e;
    if (down) {
        const int families = (pw_emit ? 0 : (1 << FAMILY_POINTWISE)) | (sh_emit ? 0 : (1 << FAMILY_SHARPEN)) | (1 << FAMILY_NLM);
        e = launch_block_mean_masked(out, down, B, H, W, down_h, down_w, ops, seq_len, S, families, st);
    }
    return (int)e;
}

// ---- feature -> parameter regressors of a whole filter bank, one launch each way
int aisp_regress_fwd(const float* raw, const int32_t* filter_ops, const int32_t* offsets, int B, int F, int Ntot,
                     const float* cfg_ranges, float* packed, void* stream) {
    if (!raw || !filter_ops || !offsets || !cfg_ranges || !packed) return AISP_ERR_NULL;
    if (B <= 0 || F <= 0 || Ntot <= 0 || (long long)B * F > (1LL << 30)) return AISP_ERR_SHAPE;
    return (int)launch_regress(raw, filter_ops, offsets, B, F, Ntot, cfg_ranges, packed, nullptr, nullptr,
                               (cudaStream_t)stream);
}

int aisp_regress_bwd(const float* raw, const float* grad_packed, const int32_t* filter_ops, const int32_t* offsets,
                     int B, int F, int Ntot, const float* cfg_ranges, float* grad_raw, void* stream) {
    if (!raw || !grad_packed || !filter_ops || !offsets || !cfg_ranges || !grad_raw) return AISP_ERR_NULL;
    if (B <= 0 || F <= 0 || Ntot <= 0 || (long long)B * F > (1LL << 30)) return AISP_ERR_SHAPE;
    return (int)launch_regress(raw, filter_ops, offsets, B, F, Ntot, cfg_ranges, nullptr, grad_packed, grad_raw,
                               (cudaStream_t)stream);
}

// ---- filter bank: F filters applied to the SAME batch (agent.py:103-107 runs every cfg.filter on
// the input and stacks the results).  The op list is a HOST array; each family launches only over
// its own slots (BankMap), so there are no idle CTAs and no device-side ops array.
static int make_bank_maps(const int32_t* fops, int F, BankMap m[3]) {
    if (F < 1 || F > kMaxBankFilters) return AISP_ERR_SHAPE;
    unsigned long long opsn = 0;
    for (int f = 0; f < F; ++f) {
        if (fops[f] < 0 || fops[f] >= AISP_OP_COUNT) return AISP_ERR_UNSUPPORTED;
        opsn |= (unsigned long long)fops[f] << (4 * f);
    }
    for (int k = 0; k < 3; ++k) m[k] = BankMap{F, 0, 0ull, opsn};
    for (int f = 0; f < F; ++f) {
        const int k = is_pointwise(fops[f]) ? FAMILY_POINTWISE : is_sharpen(fops[f]) ? FAMILY_SHARPEN : FAMILY_NLM;
        m[k].slots |= (unsigned long long)f << (4 * m[k].n);
        ++m[k].n;
    }
    return m[FAMILY_NLM].n > 1 ? AISP_ERR_UNSUPPORTED : AISP_OK;  // one compact stash per image
}

int aisp_bank_fwd(const float* img, float* out, const float* params, const int32_t* filter_ops, int B, int F, int H,
                  int W, int clip, float* nlm_dout_dh, void* stream) {
    if (!img || !out || !params || !filter_ops) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W) || F < 1 || (long long)B * F > 65535) return AISP_ERR_SHAPE;
    if (!al4(img) || !al4(out)) return AISP_ERR_ALIGN;
    if (img == out) return AISP_ERR_UNSUPPORTED;
    BankMap m[3];
    int rc = make_bank_maps(filter_ops, F, m);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (m[FAMILY_POINTWISE].n)
        e = launch_pointwise_bank_fwd(img, out, params, B, H, W, clip ? 1 : 0, m[FAMILY_POINTWISE], st);
    if (e == cudaSuccess && m[FAMILY_SHARPEN].n)
        e = launch_sharpen_fwd(img, out, params, nullptr, B * m[FAMILY_SHARPEN].n, H, W, m[FAMILY_SHARPEN], st);
    if (e == cudaSuccess && m[FAMILY_NLM].n)
        e = launch_nlm_fwd(img, out, params, nullptr, B * m[FAMILY_NLM].n, H, W, nlm_dout_dh, nullptr, m[FAMILY_NLM], st);
    return (int)e;
}

int aisp_bank_bwd(const float* img, const float* grad_out, const float* params, const int32_t* filter_ops, int B,
                  int F, int H, int W, int clip, const float* nlm_dout_dh, float* grad_params, void* scratch,
                  size_t scratch_bytes, void* stream) {
    if (!img || !grad_out || !params || !filter_ops || !grad_params || !scratch) return AISP_ERR_NULL;
    if (!shape_ok(B, H, W) || F < 1 || (long long)B * F > 65535) return AISP_ERR_SHAPE;
    if (scratch_bytes < aisp_bwd_scratch_bytes(B * F, H, W)) return AISP_ERR_SCRATCH;
    if (!al4(img) || !al4(grad_out)) return AISP_ERR_ALIGN;
    BankMap m[3];
    int rc = make_bank_maps(filter_ops, F, m);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    // the per-pixel slots go out in two launches: ColorFilter slots have their own instantiation
    // (27 partial sums: see pw_bwd_kernel); cfg.filters holds none, so normally this is one launch
    BankMap plain = m[FAMILY_POINTWISE], color = m[FAMILY_POINTWISE];
    plain.n = color.n = 0;
    plain.slots = color.slots = 0ull;
    for (int j = 0; j < m[FAMILY_POINTWISE].n; ++j) {
        const unsigned long long f = (m[FAMILY_POINTWISE].slots >> (4 * j)) & 15ull;
        BankMap& t = (filter_ops[f] == AISP_OP_COLOR) ? color : plain;
        t.slots |= f << (4 * t.n);
        ++t.n;
    }
    if (plain.n)
        e = launch_pointwise_bank_bwd(img, grad_out, params, B, H, W, clip ? 1 : 0, grad_params, (float*)scratch, plain,
                                      false, st);
    if (e == cudaSuccess && color.n)
        e = launch_pointwise_bank_bwd(img, grad_out, params, B, H, W, clip ? 1 : 0, grad_params, (float*)scratch, color,
                                      true, st);
    if (e == cudaSuccess && m[FAMILY_SHARPEN].n)
        e = launch_sharpen_bwd(img, grad_out, params, nullptr, B * m[FAMILY_SHARPEN].n, H, W, grad_params, nullptr,
                               nullptr, (float*)scratch, m[FAMILY_SHARPEN], no_pooled_grad(), st);
    if (e == cudaSuccess && m[FAMILY_NLM].n && nlm_dout_dh)
        e = launch_nlm_bwd(grad_out, nlm_dout_dh, nullptr, nullptr, B * m[FAMILY_NLM].n, H, W, grad_params,
                           (float*)scratch, m[FAMILY_NLM], no_pooled_grad(), st);
    return (int)e;
}

}  // extern "C"
