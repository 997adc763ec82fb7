/*
 * rtsds_b200.h — C ABI of librtsds_b200.so (hand-written sm_100a kernels).
 *
 * The reference (sina-behnam/RTSDS) has no FFI of its own: its hot path is
 * the Python nn.Module API.  This header is the boundary the
 * drop-in modules under models/ and utils.fast_hist bind through ctypes
 * (rtsds_b200/_lib.py).  Every entry point
 *   - takes plain device pointers, sizes and a cudaStream_t (as void*),
 *   - is asynchronous on that stream, never allocates, never synchronises,
 *     and is CUDA-graph capturable,
 *   - returns 0 on success or a negative RTSDS_E* code; the message is
 *     available from rtsds_last_error_string() (thread-local),
 *   - has NO CPU fallback: a non-sm_100 device is RTSDS_EARCH.
 *
 * Activation layout inside the path is NHWC (channels innermost): 16-bit in
 * the production modes (fp16 for eval-mode inference, bf16 for training) and
 * fp32 in the "fp32 check" mode (BASELINE.json tolerance 1e-4).  API-boundary tensors (input image, returned logits,
 * labels) keep the reference's NCHW fp32 / int64 layout.
 *
 * Reference file:line citations are relative to /root/reference.
 */
#ifndef RTSDS_B200_H
#define RTSDS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTSDS_ABI_VERSION 3

/* error codes */
#define RTSDS_OK        0
#define RTSDS_EINVAL   -1   /* bad shape / argument */
#define RTSDS_EARCH    -2   /* device is not sm_100 */
#define RTSDS_ECUDA    -3   /* CUDA runtime / launch error */
#define RTSDS_EUNSUP   -4   /* configuration not supported by this kernel */
#define RTSDS_EWS      -5   /* workspace too small */

/* element types of activation buffers */
#define RTSDS_F32   0
#define RTSDS_BF16  1
#define RTSDS_F16   2   /* IEEE half: the inference (eval-mode) storage/operand type — same tcgen05 kind::f16 rate as bf16,
                           11-bit significand; on BASELINE config 1 it keeps argmax agreement >= 99.9 % where bf16 gives 99.78 % */

/* activations fused into conv epilogues */
#define RTSDS_ACT_NONE   0
#define RTSDS_ACT_RELU   1
#define RTSDS_ACT_LRELU  2   /* LeakyReLU(slope) — models/domain_shift/adversarial/model.py:50 */

typedef void* rtsds_stream_t;   /* cudaStream_t */

int         rtsds_abi_version(void);
const char* rtsds_last_error_string(void);
/* RTSDS_OK if the current device is compute capability 10.x, else RTSDS_EARCH. */
int         rtsds_check_device(void);
/* number of kernel launches issued by this library since load (all threads). */
int64_t     rtsds_launch_count(void);
/* Deterministic mode (default off; RTSDS_DETERMINISTIC=1 in the environment turns it on at load).  SURVEY.md §5/§7.2 ask
 * for run-to-run reproducible parity runs: ATen's CPU path, which the reference's results come from, sums in a fixed
 * order.  When on, the cross-CTA floating-point reductions of the training path — train-mode BatchNorm sum / sum of
 * squares in every conv and stem epilogue, the weight-gradient partials of every wgrad kernel, the BatchNorm-backward
 * sums, the split global pool, bias gradients, the ARM gate / FFM head gradients and the loss-gradient scatter of the CE
 * kernels — are accumulated exactly (64.64 fixed point, integer atomics, one rounding to fp32 at the end) in library-owned
 * scratch instead of through fp32 atomics, so results do not depend on CTA arrival order.  Buffers and signatures are
 * unchanged; each affected call costs one memset and one small finishing launch more. */
void        rtsds_set_deterministic(int on);
int         rtsds_get_deterministic(void);

/* ------------------------------------------------------------------------
 * Metric: utils.fast_hist (utils.py:52-58)
 *   k = (a >= 0) & (a < n); bincount(n*a[k] + b[k], minlength=n*n)
 * hist (int64 [n_cls*n_cls], row = label, col = prediction) is ACCUMULATED
 * into (caller zeroes it), which is what validation.py:55 does with
 * total_hist.  Predictions must lie in [0, n_cls) for kept pixels, as they do
 * for an argmax over n_cls channels; out-of-range predictions are counted in
 * *n_bad (int64, may be NULL) instead of corrupting a neighbouring bin.
 * Bit-exact integer arithmetic.
 * ---------------------------------------------------------------------- */
int rtsds_confusion_hist(const int64_t* label, const int64_t* pred, int64_t n_pix,
                         int n_cls, int64_t* hist, int64_t* n_bad, rtsds_stream_t s);

/* Fused torch.argmax(logits, 1) (validation.py:51) + fast_hist.
 * logits: fp32 NCHW [n, n_cls, hw]; label int64 [n, hw];
 * pred_out: optional int64 [n, hw] (argmax, first max on ties);
 * hist accumulated as above. */
int rtsds_argmax_hist(const float* logits, const int64_t* label, int n, int n_cls,
                      int64_t hw, int64_t* pred_out, int64_t* hist, rtsds_stream_t s);
/* Same with the class map written as uint8 [n, hw] (n_cls <= 256): the prediction a serving loop copies back to the host
 * is 1 byte per pixel instead of the 8 of torch.argmax's int64 (validation.py:51 `.cpu()`). */
int rtsds_argmax_hist_u8(const float* logits, const int64_t* label, int n, int n_cls,
                         int64_t hw, uint8_t* pred_u8_out, int64_t* hist, rtsds_stream_t s);

/* ------------------------------------------------------------------------
 * Device-side input pipeline: the per-sample CPU work of the reference's Dataset + transforms
 * (main.py:60-108; datasets/cityscapes.py:66-72; datasets/gta5.py:68-82; utils.py:67-75) on the GPU, from the RAW
 * uint8 planes.
 *   image_u8_to_f32: read_image(...).float() -> transforms.Resize(size, antialias=True) -> transforms.Normalize:
 *       src uint8 NCHW [n,c,h,w] (c <= 3) -> dst fp32 NCHW [n,c,oh,ow] = resize(float(src)) * scale3[ch] + bias3[ch]
 *       (scale = 1/std, bias = -mean/std; HOST pointers).  Resize = ATen _upsample_bilinear2d_aa (triangle filter whose
 *       support grows with the down-scaling factor; identity for equal sizes).
 *   label_resize_clamp: read_image(...).long() -> Resize -> IntRangeTransformer(lo, hi): src uint8 or int64 [n,h,w] ->
 *       int64 [n,oh,ow] = clamp(round_half_even(resize(float(src))), lo, hi)   (clamp == 0: no clamp, GTA5 labels).
 * For frames that already have the network's size the image kernel is a pure convert + normalise pass (16 pixels per
 * thread): ~3 us for a 512x1024 frame, after which the fused stems run unchanged.  (A variant of the stem kernel that read
 * the uint8 frame itself was measured 20 us SLOWER per frame: the conversion lands on the gather warps, which are that
 * kernel's critical path — profiles/r02_summary.md.)
 * ---------------------------------------------------------------------- */
int rtsds_image_u8_to_f32(const uint8_t* src, int n, int c, int h, int w, int oh, int ow, const float* scale3,
                          const float* bias3, float* dst, rtsds_stream_t s);
int rtsds_label_resize_clamp(const void* src, int src_is_u8, int n, int h, int w, int oh, int ow, int clamp,
                             int64_t lo, int64_t hi, int64_t* dst, rtsds_stream_t s);
/* F.adaptive_avg_pool2d on NCHW fp32 (train.py:410,438,445: adversarial_train_2 pools the generator's logits to the target
 * label size before softmax -> discriminator); planes = n*c.  Windows as ATen: [floor(o*in/out), ceil((o+1)*in/out)).
 * bwd: dx[i] = sum over the windows containing i of dy / area (gather form, deterministic). */
int rtsds_adaptive_avgpool_nchw_fwd(const float* x, int64_t planes, int h, int w, int oh, int ow, float* y, rtsds_stream_t s);
int rtsds_adaptive_avgpool_nchw_bwd(const float* dy, int64_t planes, int h, int w, int oh, int ow, float* dx, rtsds_stream_t s);


/* ------------------------------------------------------------------------
 * Convolution (nn.Conv2d as used by models/bisenet/build_bisenet.py:11-12,
 * torchvision BasicBlock, models/deeplabv2/deeplabv2.py:13-24,59-61,
 * models/domain_shift/adversarial/model.py:45-49).
 *
 * y[n,oh,ow,co] = act( scale[co] * (sum_{r,s,ci} x[n, oh*stride-pad+r*dil,
 *                 ow*stride-pad+s*dil, ci] * w[co,r,s,ci]) + shift[co]
 *                 + residual[n,oh,ow,co] )
 * x: NHWC, pixel pitch in_ld elements (>= cin: views into concat buffers);
 * y / residual: NHWC, pixel pitch out_ld / res_ld.
 * w: packed [cout_pad][kh*kw][cin] (rtsds_pack_conv_weight), dtype = x dtype.
 * scale/shift: fp32 [cout] or NULL (1 / 0) — folded eval BatchNorm or bias.
 * stats: fp32 [2*cout] or NULL; when given, per-channel sum and sum of
 *        squares of the RAW convolution result (before scale/shift) are
 *        atomically accumulated: train-mode BatchNorm statistics.
 * ---------------------------------------------------------------------- */
typedef struct RtsdsConvDesc {
    int n, h, w, cin, in_ld;
    int cout, out_ld, res_ld;
    int kh, kw, stride, pad, dil;
    int oh, ow;
    int act;
    float slope;
    int in_dtype;      /* RTSDS_BF16 / RTSDS_F16 (tensor-core path; the packed weights have the same type) or RTSDS_F32 (check path) */
    int out_dtype;     /* in_dtype or RTSDS_F32 */
    int split_k;       /* >1: split the reduction over this many CTAs (tensor-core path) */
} RtsdsConvDesc;

/* tcgen05/TMEM implicit GEMM fed by TMA; x, w bf16; cin % 64 == 0.
 * workspace: fp32 [n*oh*ow*cout_pad32] zero-filled by the call when split_k>1. */
int rtsds_conv2d_tc_fwd(const RtsdsConvDesc* d, const void* x, const void* w,
                        const float* scale, const float* shift, const void* residual,
                        float* stats, void* y, void* workspace, size_t ws_bytes,
                        rtsds_stream_t s);
/* Same with AdaptiveAvgPool2d(1) of the layer's OUTPUT fused into the epilogue (build_bisenet.py:46 ARM pooling,
 * build_contextpath.py:27-28 tail), deterministically: gap_out fp32 [n][parts][cout], parts =
 * rtsds_conv2d_tc_gap_parts(d); every CTA WRITES the partial mean (sum of the final post-activation values of its rows /
 * (oh*ow)) of its part; the mean is the sum over parts, taken by the consumer in index order.  Nothing to zero. */
int rtsds_conv2d_tc_gap_parts(const RtsdsConvDesc* d);
int rtsds_conv2d_tc_fwd_gap(const RtsdsConvDesc* d, const void* x, const void* w,
                            const float* scale, const float* shift, const void* residual,
                            void* y, float* gap_out, void* workspace, size_t ws_bytes, rtsds_stream_t s);
size_t rtsds_conv2d_tc_workspace_bytes(const RtsdsConvDesc* d);
/* output-channel padding of the packed weight layout (32, 64 or a multiple of 128). */
int rtsds_conv_cout_pad(int cout);
/* process-wide tuning override for experiments: block_n in {0=auto,32,64,128}, stages (0=auto). */
void rtsds_conv2d_tc_tune(int block_n, int stages);
/* debug: while `buf` (device memory, 256 bytes per CTA of the largest grid) is non-NULL, every non-persistent conv_tc launch
 * writes per-CTA time stamps into it (globaltimer at entry/exit, clock64 at: entry, setup done, dependency resolved, last TMA
 * issued, first operands landed, last MMA issued, accumulator complete, own epilogue done, exit).  tools/conv_timeline.py. */
void rtsds_debug_conv_trace(void* buf);
/* debug: 4 clock64 stamps per block of arm_gate_resize (64 bytes per block): start, pooled vector ready, gates ready, done. */
void rtsds_debug_arm_trace(void* buf);

/* CUDA-core implicit GEMM with fp32 accumulation; any cin; x/w dtype per
 * d->in_dtype.  This is the fp32 check mode (and a bf16 cross-check). */
int rtsds_conv2d_simt_fwd(const RtsdsConvDesc* d, const void* x, const void* w,
                          const float* scale, const float* shift, const void* residual,
                          float* stats, void* y, rtsds_stream_t s);

/* OIHW fp32 (nn.Conv2d.weight) -> [cout_pad][kh*kw][cin] of dtype; rows
 * cout..cout_pad-1 are zero. */
int rtsds_pack_conv_weight(const float* w_oihw, int cout, int cin, int kh, int kw,
                           int cout_pad, int dtype, void* w_packed, rtsds_stream_t s);

/* ------------------------------------------------------------------------
 * Convolution backward (autograd of nn.Conv2d, driven by loss.backward() at
 * train.py:95,213,233,251,262).  `d` is always the FORWARD geometry.
 *
 * dgrad: dx[n,h,w,ci] = sum_{r,s,co} dy[n,(h+pad-r*dil)/stride,(w+pad-s*dil)/stride,co]
 *                       * W[co,ci,r,s]  (+ residual[n,h,w,ci], which may alias dx)
 *   dy: NHWC, dtype d->in_dtype, pitch d->out_ld.  Tensor-core path: pitch >=
 *       roundup(cout,64) with the channels cout.. zero.
 *   w_dgrad: [cin_pad][kh*kw][ck] from rtsds_pack_conv_weight_dgrad with
 *       cin_pad = rtsds_conv_cout_pad(cin); ck = roundup(cout,64) (tensor-core)
 *       or cout (CUDA-core).
 *   dx: NHWC of dx_dtype, pitch d->in_ld; residual pitch d->res_ld (0: in_ld).
 * wgrad: dw_packed[co][r*kw+s][ci] (fp32) += sum_{n,oh,ow} dy[n,oh,ow,co] *
 *        x[n,oh*stride-pad+r*dil,ow*stride-pad+s*dil,ci]; the caller zeroes
 *        dw_packed; rtsds_unpack_conv_wgrad converts it to the OIHW layout of
 *        nn.Conv2d.weight.grad (assign or accumulate).
 * ---------------------------------------------------------------------- */
size_t rtsds_conv2d_tc_dgrad_workspace_bytes(const RtsdsConvDesc* d);
int rtsds_conv2d_tc_dgrad(const RtsdsConvDesc* d, const void* dy, const void* w_dgrad,
                          const void* residual, void* dx, int dx_dtype, void* workspace,
                          size_t ws_bytes, rtsds_stream_t s);
int rtsds_conv2d_tc_wgrad(const RtsdsConvDesc* d, const void* x, const void* dy,
                          float* dw_packed, rtsds_stream_t s);
int rtsds_conv2d_simt_dgrad(const RtsdsConvDesc* d, const void* dy, const void* w_dgrad,
                            const void* residual, void* dx, int dx_dtype, rtsds_stream_t s);
int rtsds_conv2d_simt_wgrad(const RtsdsConvDesc* d, const void* x, const void* dy,
                            float* dw_packed, rtsds_stream_t s);
int rtsds_pack_conv_weight_dgrad(const float* w_oihw, int cout, int cin, int kh, int kw,
                                 int cin_pad, int ck, int dtype, void* w_packed,
                                 rtsds_stream_t s);
int rtsds_unpack_conv_wgrad(float* dw_packed /* zeroed on return */, int cout, int cin, int kh, int kw,
                            int accumulate, float* grad_oihw, rtsds_stream_t s);

/* Batched forms of the weight (re)packing and gradient unpacking above: one launch for up to 40 / 48 layers
 * (the per-layer kernels are launch-latency bound: a training step repacks every conv weight after
 * optimizer.step(), train.py:96, and converts every wgrad result).  `jobs` is a HOST array.
 *   pack job:   kind 0 -> out[cout_pad][taps][cin_pad];  kind 1 -> dgrad operand out[cin_pad][taps][ck]
 *   unpack job: as rtsds_unpack_conv_wgrad_cpad (dw_packed is zeroed on return). */
typedef struct RtsdsPackJob {
    const float* w;     /* OIHW fp32 */
    void* out;
    int cout, cin, cin_pad, taps, cout_pad, kind, ck;
} RtsdsPackJob;
typedef struct RtsdsUnpackJob {
    float* dw_packed;   /* fp32 [cout][taps][cin_src] */
    float* grad;        /* OIHW fp32 */
    int cout, cin, cin_src, taps, accumulate;
} RtsdsUnpackJob;
int rtsds_pack_conv_weights_batch(const RtsdsPackJob* jobs, int n_jobs, int dtype, rtsds_stream_t s);
int rtsds_unpack_conv_wgrads_batch(const RtsdsUnpackJob* jobs, int n_jobs, rtsds_stream_t s);

/* ------------------------------------------------------------------------
 * Fused optimizer step: torch.optim.Adam / SGD as built at main.py:110-120 and stepped at train.py:96,
 * 269-270, for ALL parameter tensors in ONE launch, together with the refresh of the packed conv operands the next
 * forward / backward read (what rtsds_pack_conv_weights_batch did in separate launches).
 *   job:   p / g / m / v fp32 device pointers of one parameter tensor (g NULL: the tensor is only re-packed; m NULL for
 *          momentum-less SGD; v unused by SGD), `group` = index into the hyper-parameter arrays.
 *          taps == 0: plain tensor of `numel` elements.  taps > 0: conv weight OIHW [cout][cin][taps] whose new value is
 *          also written to out_fwd [cout_pad][taps][cin_pad_fwd] and / or out_dgrad [cin_pad_dgrad][taps][ck] (either may
 *          be NULL) in `pack_dtype`.
 *   hyper: per-group lr / weight_decay (L2, added to the gradient as torch does) read every step — poly_lr_scheduler
 *          (utils.py:33-48) rewrites param_groups[0]['lr'] only; Adam: beta1, beta2, eps and the bias corrections of THIS
 *          step (1/(1-beta1^t), 1/sqrt(1-beta2^t)); SGD: momentum, first_step != 0 on the step that initialises the
 *          momentum buffer (torch: buf = grad).
 * jobs / first_block (prefix sums of rtsds_optim_job_blocks over the jobs, n_jobs entries) live in DEVICE memory.
 * ---------------------------------------------------------------------- */
#define RTSDS_OPT_ADAM 0
#define RTSDS_OPT_SGD  1
#define RTSDS_OPT_MAX_GROUPS 8
typedef struct RtsdsOptJob {
    float* p; const float* g; float* m; float* v;
    int64_t numel;
    int group;
    int taps, cout, cin;
    int cout_pad, cin_pad_fwd, cin_pad_dgrad, ck;
    void* out_fwd; void* out_dgrad;
} RtsdsOptJob;
typedef struct RtsdsOptHyper {
    int kind;                                   /* RTSDS_OPT_ADAM / RTSDS_OPT_SGD */
    int first_step;
    float lr[RTSDS_OPT_MAX_GROUPS];
    float weight_decay[RTSDS_OPT_MAX_GROUPS];
    float beta1, beta2, eps, inv_bias_correction1, inv_bias_correction2_sqrt;
    float momentum;
} RtsdsOptHyper;
int rtsds_optim_job_blocks(const RtsdsOptJob* job);
int rtsds_optim_step(const RtsdsOptJob* jobs_dev, const int* first_block_dev, int n_jobs, int total_blocks,
                     const RtsdsOptHyper* hyper, int pack_dtype, rtsds_stream_t s);

/* ------------------------------------------------------------------------
 * "Taps as N": k x k stride-1 conv with few output channels (c <= 32; the FFM ConvBlock 3x3 1024 -> 19,
 * build_bisenet.py:64,74) evaluated so that the wide input is read once instead of once per tap:
 *   forward : T = 1x1 conv of x with w_fwd (virtual OIHW [k*k*c, cin, 1, 1], N index t*c+co), then
 *             y[p,co] = act(scale*sum_t T[p+off(t), t*c+co] + shift)  (tapn_gather; optional BN statistics of the raw sum)
 *   backward: G[q,t*c+co] = dy[q-off(t),co] (tapn_scatter; g_ld >= k*k*c, pad columns zeroed), then
 *             dx = 1x1 conv of G with w_bwd (virtual OIHW [cin, kpad, 1, 1]) and dW2 = 1x1 wgrad(G, x);
 *             tapn_weight_grad ACCUMULATES dW2 ([k*k*c, cin]) into the [c, cin, k, k] gradient.
 * ---------------------------------------------------------------------- */
int rtsds_tapn_weights(const float* w_oihw, int c, int cin, int k, int kpad, float* w_fwd, float* w_bwd,
                       rtsds_stream_t s);
int rtsds_tapn_weight_grad(const float* dw2, int c, int cin, int k, float* grad_oihw, rtsds_stream_t s);
int rtsds_tapn_gather(const float* t_buf, int t_ld, int n, int h, int w, int c, int k, int pad, int dil,
                      const float* scale, const float* shift, int act, float* stats, float* y, int y_ld,
                      float* gap_out /* NULL, or fp32 [n][rtsds_tapn_gather_parts(n,h,w)][c]: one partial mean of y per block,
                                        written (deterministic: the consumer adds them in order) */, rtsds_stream_t s);
int rtsds_tapn_gather_parts(int n, int h, int w);
int rtsds_tapn_scatter(const void* dy, int dy_ld, int dy_dtype, int n, int h, int w, int c, int k, int pad, int dil,
                       void* g, int g_ld, int g_dtype, rtsds_stream_t s);

/* Stem convolutions read the API-boundary image directly:
 * x fp32 NCHW [n,cin,h,w] (cin <= 32), w fp32 OIHW, y NHWC of out_dtype,
 * with scale/shift/act/stats as above.  softmax_in != 0 applies a softmax
 * over the cin channels of x while loading (train.py:225 F.softmax feeding
 * the discriminator's conv1). */
int rtsds_stem_conv_fwd(const float* x, const float* w_oihw, int n, int cin, int h, int w,
                        int cout, int k, int stride, int pad,
                        const float* scale, const float* shift, int act, float slope,
                        int softmax_in, float* stats, int out_dtype, void* y,
                        rtsds_stream_t s);

/* Tensor-core stems of BiSeNet: the context-path conv7x7 s2 p3 3->64 (build_contextpath.py:19) and the
 * spatial-path conv3x3 s2 p1 3->64 (build_bisenet.py:24) read the same image on the same output grid and are
 * fused into one tcgen05 implicit GEMM (N = 64 + 64) whose im2col tile is gathered into swizzled shared
 * memory.  dtype: RTSDS_BF16 or RTSDS_F16 (operands and outputs; the weight gradient is bf16 only).
 * wpk: [128][192] of dtype from rtsds_stem_pack_weights; scale/shift: fp32 [128] (context-path BN in
 * 0..63, spatial-path BN in 64..127) or NULL; stats_*: fp32 [2*64] train-mode sums or NULL.
 * wgrad: d_raw_*: NHWC bf16 [n,oh,ow,64]; dw_ws: fp32 [128*192] scratch, zero on entry and on return;
 * g7/g3: OIHW fp32 gradients, accumulated. */
int rtsds_stem_pack_weights(const float* w7_oihw, const float* w3_oihw, int dtype, void* wpk, rtsds_stream_t s);
int rtsds_stem_pair_tc_fwd(const float* x, int n, int h, int w, const void* wpk, const float* scale,
                           const float* shift, int relu, float* stats_cp, float* stats_sp, int dtype, void* y_cp,
                           void* y_sp, rtsds_stream_t s);
/* Eval-mode form with the context path's nn.MaxPool2d(3, 2, 1) (build_contextpath.py:21) fused into the epilogue: the
 * 1/2-resolution context-path map is never written; `pool` (NHWC [n, (oh-1)/2+1, (ow-1)/2+1, 64] of dtype) must be ZERO on
 * entry and receives the pooled map (16-byte max-reductions of post-ReLU values: bit-identical to pooling the stored map). */
int rtsds_stem_pair_tc_fwd_pool(const float* x, int n, int h, int w, const void* wpk, const float* scale,
                                const float* shift, int dtype, void* pool, void* y_sp, rtsds_stream_t s);
int rtsds_stem_pair_tc_wgrad(const float* x, int n, int h, int w, const void* d_raw_cp,
                             const void* d_raw_sp, float* dw_ws, float* g7_oihw, float* g3_oihw,
                             rtsds_stream_t s);

/* Space-to-depth stems (tensor-core path): a k x k stride-2 conv on the 3-channel image (7x7 p3 or 3x3 p1) is a 4-tap
 * implicit GEMM over P, the padded space-to-depth image:
 *   P[n, i, j, (py*2+px)*3 + c] = x[n, c, 2(i-2)+py, 2(j-2)+px]   bf16 [n, oh+3, ow+3, 16], oh = (h-1)/2+1, zero outside
 * (one window row = 4 pixels x 16 channels = 64 contiguous bf16; windows overlap with a pixel pitch of 16 elements,
 * which the TMA tensor map expresses directly).  stem_s2d_weight turns OIHW [cout,3,k,k] into the virtual OIHW
 * [cout,64,4,1] weight of that GEMM (pack it with rtsds_pack_conv_weight); stem_s2d_weight_grad ACCUMULATES the
 * gradient of the virtual weight into the [cout,3,k,k] gradient.  conv_fwd has the epilogue of rtsds_conv2d_tc_fwd
 * (scale/shift/act/stats); conv_wgrad accumulates dw_packed [cout][4][64] fp32 like rtsds_conv2d_tc_wgrad. */
int rtsds_stem_s2d_pack(const float* x, int n, int h, int w, void* P, rtsds_stream_t s);
/* General form: x fp32 NCHW or (x_is_u8) the RAW uint8 frame; P = scale3[c]*x + bias3[c] inside the image (HOST float[3]
 * each, NULL = 1 / 0: transforms.Normalize of main.py:70 folded into the pack), 0 outside; P of p_dtype
 * (RTSDS_BF16 / RTSDS_F16).  rtsds_stem_s2d_conv_fwd_dt: the conv on such a P (weights packed in the same type). */
int rtsds_stem_s2d_pack_ex(const void* x, int x_is_u8, const float* scale3, const float* bias3, int n, int h, int w,
                           int p_dtype, void* P, rtsds_stream_t s);
int rtsds_stem_s2d_conv_fwd_dt(const void* P, int p_dtype, int n, int oh, int ow, const void* w_packed, int cout,
                               const float* scale, const float* shift, int act, void* y, int out_ld, int out_dtype,
                               rtsds_stream_t s);
int rtsds_stem_s2d_weight(const float* w_oihw, int cout, int k, int pad, float* w2_oihw, rtsds_stream_t s);
int rtsds_stem_s2d_weight_grad(const float* g2_oihw, int cout, int k, int pad, float* grad_oihw, rtsds_stream_t s);
int rtsds_stem_s2d_conv_fwd(const void* P, int n, int oh, int ow, const void* w_packed, int cout, const float* scale,
                            const float* shift, int act, float* stats, void* y, int out_ld, int out_dtype,
                            rtsds_stream_t s);
int rtsds_stem_s2d_conv_wgrad(const void* P, int n, int oh, int ow, const void* dy, int dy_ld, int cout,
                              float* dw_packed, rtsds_stream_t s);

/* nn.MaxPool2d(3, 2, 1[, ceil_mode]) on NHWC. */
int rtsds_maxpool3x3s2_fwd(const void* x, int n, int h, int w, int c, int dtype,
                           int ceil_mode, void* y, rtsds_stream_t s);
/* Same for an input whose pixels are x_ld >= c elements apart (a channel slice of a wider NHWC buffer). */
int rtsds_maxpool3x3s2_fwd_ld(const void* x, int n, int h, int w, int c, int x_ld, int dtype,
                              int ceil_mode, void* y, rtsds_stream_t s);
/* Training form: also writes idx [n,oh,ow,c/8] uint32 = per channel (one nibble each) the window position
 * r*3+q of the first maximum, so that the backward pass (rtsds_maxpool3x3s2_bwd_idx) does not re-read x. */
int rtsds_maxpool3x3s2_fwd_idx(const void* x, int n, int h, int w, int c, int dtype,
                               int ceil_mode, void* y, uint32_t* idx, rtsds_stream_t s);

/* ------------------------------------------------------------------------
 * BatchNorm helpers (nn.BatchNorm2d, eps/momentum per build_bisenet.py:135-137)
 * ---------------------------------------------------------------------- */
/* eval: scale = gamma / sqrt(var+eps); shift = beta - mean*scale (+ scale*conv_bias). */
int rtsds_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
                  const float* conv_bias, float eps, int c, float* scale, float* shift,
                  rtsds_stream_t s);
/* train: from stats = [sum(c), sumsq(c)] over count samples produce
 * scale/shift (biased variance) and update running_mean/var (unbiased,
 * momentum) in place; save_mean/save_invstd (fp32 [c]) kept for backward. */
int rtsds_bn_finalize(const float* stats, double count, const float* gamma, const float* beta,
                      float eps, float momentum, int c, float* running_mean, float* running_var,
                      float* scale, float* shift, float* save_mean, float* save_invstd,
                      rtsds_stream_t s);
/* rtsds_bn_finalize followed by rtsds_scale_shift_act(x, scale, shift, residual, ...) in ONE launch (train-mode
 * nn.BatchNorm2d [+ residual] [+ ReLU] of ConvBlock / BasicBlock / Bottleneck, build_bisenet.py:14-17): every thread
 * derives its channels' scale/shift from the sums with bn_finalize's arithmetic; scale/shift/save_* and the running
 * statistics are written as bn_finalize writes them.  Falls back to the two launches for channel counts or pitches the
 * vector path does not cover. */
int rtsds_bn_finalize_apply(const float* stats, double count, const float* gamma, const float* beta, float eps,
                            float momentum, int c, float* running_mean, float* running_var, float* scale,
                            float* shift, float* save_mean, float* save_invstd, const void* x,
                            const void* residual, int64_t n_pix, int x_ld, int res_ld, int y_ld, int act,
                            float slope, int x_dtype, int y_dtype, void* y, rtsds_stream_t s);
/* y = act(scale[c]*x + shift[c] + residual) elementwise over NHWC. */
int rtsds_scale_shift_act(const void* x, const float* scale, const float* shift,
                          const void* residual, int64_t n_pix, int c, int x_ld, int res_ld,
                          int y_ld, int act, float slope, int x_dtype, int y_dtype, void* y,
                          rtsds_stream_t s);

/* BatchNorm(+ReLU) backward (autograd of nn.BatchNorm2d in train mode).
 *   g = dy * (relu ? y > 0 : 1);  xhat = (raw - mean) * invstd
 * reduce: sums[c] = sum g, sums[C+c] = sum g*xhat  (zeroed by the call)
 * apply : d_raw = gamma*invstd*(g - sums[c]/M - xhat*sums[C+c]/M); g_out = g
 *         (optional: gradient of a residual branch); dgamma += sums[C+c],
 *         dbeta += sums[c] (optional, accumulated). */
int rtsds_bn_bwd_reduce(const void* dy, int dy_ld, const void* y, int y_ld, const void* raw,
                        int raw_ld, const float* mean, const float* invstd, int64_t n_pix, int c,
                        int relu, int dtype, float* sums, rtsds_stream_t s);
int rtsds_bn_bwd_apply(const void* dy, int dy_ld, const void* y, int y_ld, const void* raw,
                       int raw_ld, const float* mean, const float* invstd, const float* gamma,
                       const float* sums, int64_t n_pix, int c, int relu, int dtype, void* d_raw,
                       int d_raw_ld, int d_raw_dtype, void* g_out, int g_ld, float* dgamma,
                       float* dbeta, rtsds_stream_t s);
/* Same for a BatchNorm+ReLU layer WITHOUT a residual input: the ReLU mask is recomputed from raw as
 * fma(raw, fwd_scale, fwd_shift) > 0 with the scale/shift vectors rtsds_bn_finalize produced for the forward (the
 * very fma the forward evaluated, so the mask is bit-identical) and the activation y is not re-read. */
int rtsds_bn_bwd_reduce_rawmask(const void* dy, int dy_ld, const void* raw, int raw_ld, const float* mean,
                                const float* invstd, const float* fwd_scale, const float* fwd_shift, int64_t n_pix,
                                int c, int dtype, float* sums, rtsds_stream_t s);
int rtsds_bn_bwd_apply_rawmask(const void* dy, int dy_ld, const void* raw, int raw_ld, const float* mean,
                               const float* invstd, const float* gamma, const float* sums, const float* fwd_scale,
                               const float* fwd_shift, int64_t n_pix, int c, int dtype, void* d_raw, int d_raw_ld,
                               int d_raw_dtype, void* g_out, int g_ld, float* dgamma, float* dbeta, rtsds_stream_t s);
/* out[c] += sum over pixels of x[p][c] (conv bias gradients); caller zeroes or accumulates. */
int rtsds_channel_sum(const void* x, int ld, int64_t n_pix, int c, int dtype, float* out,
                      rtsds_stream_t s);
/* nn.MaxPool2d(3,2,1) backward: gradient goes to the first maximum of each window. */
int rtsds_maxpool3x3s2_bwd(const void* x, const void* dy, int n, int h, int w, int c, int dtype,
                           int ceil_mode, void* dx, rtsds_stream_t s);
int rtsds_maxpool3x3s2_bwd_idx(const uint32_t* idx, const void* dy, int n, int h, int w, int c, int dtype,
                               int ceil_mode, void* dx, rtsds_stream_t s);
/* weight gradient of a stem conv (x NCHW fp32, d_raw NHWC [n,oh,ow,64]); dw OIHW fp32 accumulated. */
int rtsds_stem_conv_wgrad(const float* x, const void* d_raw, int d_dtype, int n, int cin, int h,
                          int w, int cout, int k, int stride, int pad, float* dw_oihw,
                          rtsds_stream_t s);

/* ------------------------------------------------------------------------
 * BiSeNet glue (models/bisenet/build_bisenet.py)
 * ---------------------------------------------------------------------- */
/* Global average pool over H*W per (n, c): AdaptiveAvgPool2d(1) (:42,:46,:75)
 * and the context-path tail (build_contextpath.py:27-28).  out fp32 [n, c]. */
int rtsds_global_avgpool(const void* x, int n, int64_t hw, int c, int ld, int dtype,
                         float* out, rtsds_stream_t s);

/* AttentionRefinementModule gate (:45-49): g = sigmoid(BN(W p + b)), p = pooled [n,c].
 * w fp32 [c,c] (conv 1x1 weight), b fp32 [c].  train==0: BN folded from
 * running stats (gamma,beta,mean,var); train!=0: batch statistics over n
 * (biased var) and running stats updated in place (n must be >= 2, as in
 * torch).  mul (fp32 [n,c], may be NULL) is multiplied into the gate: the
 * `cx2 * tail` of :149. gate out fp32 [n,c].  Optional fp32 [n,c] outputs
 * lin_out (W p + b) and xhat_out (normalised) are kept for backward. */
int rtsds_arm_gate(const float* pooled, const float* w, const float* b, const float* gamma,
                   const float* beta, float* running_mean, float* running_var, float eps,
                   float momentum, int train, int n, int c, const float* mul, float* gate,
                   float* lin_out, float* xhat_out, rtsds_stream_t s);

/* Bilinear resize (F.interpolate(mode='bilinear', align_corners=False), :151-152)
 * of src NHWC [n,h,w,c] scaled by gate[n,c] (NULL = 1) and the constant gate_scale into dst NHWC
 * [n,oh,ow,dst_ld] at channel offset dst_coff: writes straight into the
 * torch.cat buffer of :153/:72. */
int rtsds_gate_resize_nhwc(const void* src, int n, int h, int w, int c, int src_ld,
                           const float* gate, float gate_scale, int oh, int ow, void* dst, int dst_ld,
                           int dst_coff, int dtype, rtsds_stream_t s);
/* Multiply input channels [c0, c1) of a packed conv weight ([rows][cin] of dtype, rows = cout_pad*taps) by `factor`:
 * the block exponent of an activation slot stored scaled (fp16 inference keeps the `cx2 * tail` slot of the concat
 * buffer, build_bisenet.py:149 — quadratic in the activations — at 2^-8 and its FFM weights at 2^8). */
int rtsds_scale_packed_channels(void* w_packed, int dtype, int64_t rows, int cin, int c0, int c1, float factor,
                                rtsds_stream_t s);

/* Eval mode, both AttentionRefinementModules and the two gated resizes into the concat buffer (:147-153) in ONE launch:
 * every block evaluates the gates of its own 32 channels from the pooled vector (folded BatchNorm) and streams its share
 * of destination pixels.  pooled: fp32 [n,pooled_parts,c], the mean of src over its pixels as pooled_parts partial means
 * that are added in index order (rtsds_conv2d_tc_fwd_gap writes one per CTA; a plain mean is pooled_parts = 1);
 * mul_pooled != 0 multiplies the gate by pooled[n,c] (`cx2 * tail`, :149); out_scale: constant factor (block exponent of
 * the fp16 cx2 slot).  dst: NHWC [n,oh,ow,dst_ld] of dtype, channels dst_coff .. dst_coff+c of each side. */
typedef struct RtsdsArmSide {
    const void* src;            /* NHWC [n,h,w_in,c] of dtype (pitch c) */
    const float* pooled;
    const float* w; const float* b; const float* gamma; const float* beta;
    const float* running_mean; const float* running_var;
    float eps, out_scale;
    int h, w_in, c, dst_coff, mul_pooled, pooled_parts;
} RtsdsArmSide;
int rtsds_arm_gate_resize(const RtsdsArmSide* a3, const RtsdsArmSide* a4, int dtype, int n, int oh, int ow, void* dst,
                          int dst_ld, rtsds_stream_t s);
/* rtsds_ffm_head + rtsds_resize_to_nchw in one kernel: out fp32 NCHW [n,c,oh,ow] = bilinear resize of
 * Wc (f*a + f) + bc, z evaluated in shared memory for the source rows each block needs (f fp32, pitch >= 32).
 * pooled: fp32 [n][pooled_parts][c] partial means of f, added in index order (a plain mean: pooled_parts = 1). */
int rtsds_ffm_head_resize(const float* f, int f_ld, const float* pooled, int n, int h, int w, int c, const float* w1,
                          const float* b1, const float* w2, const float* b2, const float* wc, const float* bc,
                          float* attn_out, int pooled_parts, int oh, int ow, float* out, rtsds_stream_t s);

/* FeatureFusionModule attention (:75-80) + final 1x1 conv (:167), evaluated at
 * feature resolution (the 1x1 conv commutes with the bilinear resize):
 *   a = sigmoid(W2 relu(W1 mean(f) + b1) + b2);   g = f*a + f;
 *   z = Wc g + bc   (Wc NULL: z = g, the with_interpolation=False result)
 * f: NHWC [n,hw,f_ld] (first c channels valid), pooled = mean(f) fp32 [n,c].
 * z: fp32 NHWC [n,hw,z_ld].  c <= 32. */
int rtsds_ffm_head(const void* f, int f_dtype, int f_ld, const float* pooled, int n, int64_t hw,
                   int c, const float* w1, const float* b1, const float* w2, const float* b2,
                   const float* wc, const float* bc, float* attn_out, float* z, int z_ld,
                   rtsds_stream_t s);

/* Bilinear resize of z fp32 NHWC [n,h,w,z_ld] (c valid channels) to the
 * API-boundary logits fp32 NCHW [n,c,oh,ow] (:158-159,:166; deeplabv2.py:126). */
int rtsds_resize_to_nchw(const float* z, int n, int h, int w, int c, int z_ld, int oh, int ow,
                         float* out, rtsds_stream_t s);

/* Backward of the glue above.
 * resize_bwd_nhwc: adjoint of the bilinear resize of rtsds_gate_resize_nhwc:
 *   d_src[n,y,x,c] (fp32 dense) = sum_dst weight * d_dst[n,oy,ox,dst_coff+c];
 *   dgate[n,c] = sum_pix d_src * src (optional; zeroed by the call).
 * gate_bwd_finish: dx = d_gated * gate[n,c] + add[n,c]*add_scale (GAP branch).
 * arm_gate_bwd: backward of rtsds_arm_gate in train mode; dpooled [n,c] out,
 *   parameter gradients accumulated; dlin_ws/dmul_ws: fp32 [n,c] scratch.
 * ffm_head_bwd: backward of rtsds_ffm_head; df [n,hw,df_ld] out, parameter
 *   gradients accumulated; da_ws/dpooled_ws fp32 [n,c] scratch.
 * resize_to_nchw_bwd: adjoint of rtsds_resize_to_nchw. */
int rtsds_resize_bwd_nhwc(const void* d_dst, int dst_ld, int dst_coff, int n, int h, int w, int c,
                          int oh, int ow, const void* src, int dtype, float* d_src, float* dgate,
                          rtsds_stream_t s);
int rtsds_gate_bwd_finish(const float* d_gated, const float* gate, const float* add,
                          float add_scale, int n, int64_t hw, int c, int out_dtype, void* dx,
                          rtsds_stream_t s);
int rtsds_arm_gate_bwd(const float* dgate, const float* pooled, const float* lin,
                       const float* xhat, const float* w, const float* gamma, const float* beta,
                       const float* mul, float eps, int n, int c, float* dlin_ws, float* dmul_ws,
                       float* dpooled, float* dw, float* dbias, float* dgamma, float* dbeta,
                       rtsds_stream_t s);
int rtsds_ffm_head_bwd(const float* dz, int dz_ld, const float* f, int f_ld, const float* pooled,
                       const float* attn, int n, int64_t hw, int c, const float* w1,
                       const float* b1, const float* w2, const float* wc, float* da_ws,
                       float* dpooled_ws, float* df, int df_ld, float* dw1, float* db1,
                       float* dw2, float* db2, float* dwc, float* dbc, rtsds_stream_t s);
int rtsds_resize_to_nchw_bwd(const float* dout, int n, int c, int oh, int ow, int h, int w,
                             float* dz, int z_ld, rtsds_stream_t s);

/* Variants of pack/unpack with the input-channel dimension zero-padded to cin_pad (activation buffers that
 * carry padding channels). */
int rtsds_pack_conv_weight_cpad(const float* w_oihw, int cout, int cin, int cin_pad, int kh, int kw,
                                int cout_pad, int dtype, void* w_packed, rtsds_stream_t s);
int rtsds_unpack_conv_wgrad_cpad(float* dw_packed, int cout, int cin, int cin_src, int kh, int kw,
                                 int accumulate, float* grad_oihw, rtsds_stream_t s);

/* ------------------------------------------------------------------------
 * Discriminators (models/domain_shift/adversarial/model.py:30-83) and the
 * adversarial step around them (train.py:218-263).
 *
 * conv1 (num_classes -> 64, 4x4 s2 p1, bias, LeakyReLU 0.2; model.py:45,54 /
 * :72,78) runs as a 2x2 stride-1 conv over the space-to-depth form of the
 * padded input:
 *   xs[n,i,j,(ph*2+pw)*32+c] = x[n,c,2i+ph-1,2j+pw-1]   (zero outside; c < 32)
 * xs: NHWC [n, h/2+1, w/2+1, 128] of dtype.  s2d_fwd reads the fp32 NCHW
 * probabilities — or, with softmax != 0, the LOGITS, applying F.softmax(dim=1)
 * (train.py:225,245,256) on the fly.  s2d_bwd is its adjoint: g fp32 NHWC gradient
 * w.r.t. xs -> dx fp32 NCHW; p (the forward xs) non-NULL applies the softmax
 * backward p*(g - sum_k g_k p_k); g is always fp32 (that difference cancels), dtype
 * describes p.
 * s2d_weight: OIHW [cout,c,4,4] -> OIHW [cout,128,2,2] (the weight of the 2x2
 * conv, packed further with rtsds_pack_conv_weight*); s2d_weight_grad is its
 * adjoint, ACCUMULATING into the [cout,c,4,4] gradient.
 * ---------------------------------------------------------------------- */
int rtsds_s2d_out_size(int in_size);
int rtsds_s2d_fwd(const float* x, int n, int c, int h, int w, int softmax, int dtype, void* xs,
                  rtsds_stream_t s);
int rtsds_s2d_bwd(const float* g, const void* p, int n, int c, int h, int w, int dtype, float* dx,
                  rtsds_stream_t s);
int rtsds_s2d_weight(const float* w_oihw, int cout, int c, float* w2_oihw, rtsds_stream_t s);
int rtsds_s2d_weight_grad(const float* gw2_oihw, int cout, int c, float* gw_oihw, rtsds_stream_t s);

/* (Leaky)ReLU backward: d_raw = dy * (y > 0 ? 1 : slope) over NHWC (c % 8 == 0);
 * dbias[c] += sum d_raw (NULL: skip) — the conv bias gradient (model.py:45-48). */
int rtsds_act_bwd(const void* dy, int dy_ld, const void* y, int y_ld, int64_t n_pix, int c, int act,
                  float slope, int dtype, void* d_raw, int d_raw_ld, float* dbias, rtsds_stream_t s);

/* classifier conv (Cout=1, 4x4 s2 p1, bias; model.py:49,58 / :73,80) + AdaptiveAvgPool2d(1) (:52,60 / :76,81):
 *   out[n] = b + 1/P * sum_{t,c} w[c,t] * tapsum[n,t,c],  tapsum = per-tap sums of x (fp32 [n,16,c],
 *   written by fwd, consumed by bwd), P = (h/2)*(w/2).  x NHWC [n,h,w,ld], c % 8 == 0.
 * bwd: g fp32 [n] = dL/dout (times g_scale, e.g. -lambda for the gradient reversal layer, model.py:9-17);
 *   dw [1,c,4,4] / dbias [1] accumulated (NULL: skip); d_raw (NULL: skip) = dL/dx, multiplied by the
 *   LeakyReLU mask of x when masked != 0 (x is then the preceding layer's activation output) and
 *   channel-summed into dbias_prev (NULL: skip). */
int rtsds_disc_cls_fwd(const void* x, int ld, int dtype, int n, int h, int w, int c, const float* w_oihw,
                       const float* bias, float* tapsum, float* out, rtsds_stream_t s);
int rtsds_disc_cls_bwd(const float* g, float g_scale, const float* tapsum, const float* w_oihw, const void* x,
                       int ld, int dtype, int n, int h, int w, int c, int masked, float slope, void* d_raw,
                       int d_ld, float* dbias_prev, float* dw, float* dbias, rtsds_stream_t s);

/* nn.BCEWithLogitsLoss (main.py:132-134) of n logits against a constant target (train.py:228-229,
 * 246-247, 257-258): loss[0] = scale * mean(...), dlogit[i] = scale * (sigmoid(x_i) - target) / n.
 * Either output may be NULL. */
int rtsds_bce_logits(const float* logit, int n, float target, float scale, float* loss, float* dlogit,
                     rtsds_stream_t s);

/* Layout converters between the reference's NCHW fp32 tensors and the path's NHWC buffers (used where a sub-module is
 * called on its own — models/bisenet/build_bisenet.py ConvBlock / Spatial_path / ARM / FFM, deeplabv2.py Bottleneck /
 * ClassifierModule — the whole-model forwards never need them).  y/x NHWC [n,hw,ld], channels c_off .. c_off+c. */
int rtsds_nchw_to_nhwc(const float* x, int n, int c, int64_t hw, int dtype, void* y, int ld, int c_off, rtsds_stream_t s);
int rtsds_nhwc_to_nchw(const void* x, int dtype, int ld, int c_off, int n, int c, int64_t hw, float* y, rtsds_stream_t s);

/* ------------------------------------------------------------------------
 * Loss: bilinear resize + nn.CrossEntropyLoss(ignore_index) (main.py:124-130,
 * train.py:86-92) + argmax / pixel accuracy (train.py:102-106) in one pass
 * over the low-resolution logits z; the full-resolution logits are never
 * materialised.
 *   acc (double [4]): += { sum of -log softmax[target] over valid pixels,
 *                          number of valid pixels,
 *                          number of pixels with argmax == target,
 *                          0 }
 *   pred_out: optional int64 [n,oh,ow] argmax.
 * Backward: dz[n,h,w,c] = sum over output pixels of bilinear weight *
 *   (softmax - onehot) * grad_scale  (grad_scale = upstream/valid_count,
 *   read from device memory: fp32 scalar). dz must be zero-filled by caller.
 * ---------------------------------------------------------------------- */
int rtsds_resize_ce_argmax_fwd(const float* z, int n, int h, int w, int c, int z_ld, int oh,
                               int ow, const int64_t* target, int64_t ignore_index,
                               double* acc, int64_t* pred_out, rtsds_stream_t s);
int rtsds_resize_ce_bwd(const float* z, int n, int h, int w, int c, int z_ld, int oh, int ow,
                        const int64_t* target, int64_t ignore_index, const float* grad_scale,
                        float* dz, rtsds_stream_t s);

/* Fused forward+backward of the same loss for the ~x8 heads (both scales > 7.1, c <= 20; query with
 * rtsds_resize_ce_fused_supported): one pass produces acc / pred_out as above AND the UNNORMALISED gradient
 * dz_unnorm[n,h,w,c] += sum over valid output pixels of bilinear weight * (softmax - onehot) (caller zeroes it;
 * NULL: skip).  The backward pass is then rtsds_scale_by_device_scalar(dz, numel, upstream / valid_count). */
int rtsds_resize_ce_fused_supported(int h, int w, int c, int oh, int ow);
int rtsds_resize_ce_fused(const float* z, int n, int h, int w, int c, int z_ld, int oh, int ow,
                          const int64_t* target, int64_t ignore_index, double* acc, int64_t* pred_out,
                          float* dz_unnorm, rtsds_stream_t s);
int rtsds_scale_by_device_scalar(float* x, int64_t n, const float* scale, rtsds_stream_t s);

/* CrossEntropyLoss + argmax on materialised NCHW logits (stock call sites). */
int rtsds_ce_argmax_nchw_fwd(const float* logits, int n, int c, int64_t hw,
                             const int64_t* target, int64_t ignore_index, double* acc,
                             int64_t* pred_out, rtsds_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* RTSDS_B200_H */
