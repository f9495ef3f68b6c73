/*
 * btslpg.h -- C ABI of libbtslpg.so: the B200 (sm_100a) implementation of the BTS decoder's
 * Local-Planar-Guidance hot path.
 *
 * The reference (clarencechen/bts-fully-tf) is pure tf.keras and has NO FFI: the boundary it
 * exposes for this path is the Keras Layer protocol.  Each entry point below therefore cites
 * the reference Python it replaces; INTEGRATION.md shows the ctypes / tf.custom_gradient stub
 * a maintainer adds to custom_layers.py and bts_decoder.py to bind them.
 *
 * Conventions
 *  - Tensors are described by BtsTensor, which is layout-compatible with DLPack's DLTensor
 *    (dlpack.h, ABI v0.8/v1.0), so a DLPack capsule from TensorFlow
 *    (tf.experimental.dlpack.to_dlpack), PyTorch or CuPy maps 1:1 with zero copies.
 *  - All tensors are CUDA device tensors (device_type kDLCUDA=2) on ONE device; host pointers
 *    are an error.  There is no CPU fallback.
 *  - Layout is the reference's NHWC (Keras channels_last).  Single-channel maps may be passed
 *    as (B,H,W,1) or (B,H,W); any batch/row/column strides are honoured (strides in ELEMENTS,
 *    NULL strides = contiguous), so an output can be a slot of a larger concat buffer.
 *  - dtype: float32 (kDLFloat,32) or bfloat16 (kDLBfloat,16) -- all tensors of a call share it,
 *    except `kernel`/`g_kernel` of the reduce entry points, which are always float32.
 *    Arithmetic is float32 in both cases.
 *  - The caller owns every buffer (outputs and workspace included).  The library never
 *    allocates, frees or retains a pointer past return, never synchronises, and is safe to
 *    capture into a CUDA Graph.  `stream` is a cudaStream_t (NULL = legacy default stream).
 *  - Return value: BTSLPG_OK (0) or a negative BtsLpgStatus; btslpg_last_error() returns a
 *    thread-local human-readable message for the last failure on the calling thread.
 *  - Re-entrant; no global mutable state apart from the thread-local error string.
 */
#ifndef BTSLPG_H_
#define BTSLPG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define BTSLPG_API __declspec(dllexport)
#else
#define BTSLPG_API __attribute__((visibility("default")))
#endif

#define BTSLPG_VERSION 100 /* 0.1.0 */

/* == DLDevice */
typedef struct {
    int32_t device_type; /* 2 = kDLCUDA (13 = kDLCUDAManaged also accepted) */
    int32_t device_id;
} BtsDevice;

/* == DLDataType */
typedef struct {
    uint8_t code;  /* 2 = kDLFloat, 4 = kDLBfloat */
    uint8_t bits;  /* 32 or 16 */
    uint16_t lanes; /* 1 */
} BtsDataType;

/* == DLTensor */
typedef struct {
    void *data;
    BtsDevice device;
    int32_t ndim;
    BtsDataType dtype;
    int64_t *shape;
    int64_t *strides; /* in elements; NULL = compact row-major */
    uint64_t byte_offset;
} BtsTensor;

typedef enum {
    BTSLPG_OK = 0,
    BTSLPG_EINVAL = -1,   /* NULL where a tensor is required, bad upratio / ds_stride */
    BTSLPG_EDTYPE = -2,   /* dtype not float32/bfloat16, or tensors of one call disagree */
    BTSLPG_ESHAPE = -3,   /* rank/extent mismatch, e.g. out is not (B, h*r, w*r[,1]) */
    BTSLPG_EDEVICE = -4,  /* not a CUDA tensor / tensors on different devices */
    BTSLPG_ELAYOUT = -5,  /* stride pattern the kernels cannot address (e.g. coef channel stride != 1) */
    BTSLPG_EWORKSPACE = -6, /* workspace NULL or too small */
    BTSLPG_ECUDA = -7     /* CUDA runtime error at launch (message carries cudaGetErrorString) */
} BtsLpgStatus;

BTSLPG_API int btslpg_version(void);
BTSLPG_API const char *btslpg_last_error(void);
BTSLPG_API const char *btslpg_status_string(int status);

/* ---------------------------------------------------------------------------------------------
 * LocalPlanarGuidance.call  -- replaces reference custom_layers.py:47-56 (with the constant of
 * build(), custom_layers.py:30-45, held in __constant__ memory instead of a (1,H,W,3) tensor)
 * and, when out_ds != NULL, the down-sampling Lambda of bts_decoder.py:81 / :88.
 *
 *   coef      (B, h, w, 3)         [phi_raw, theta_raw, dist]  (output of reduction_NxN)
 *   out_full  (B, h*r, w*r[, 1])   depth_{r}x{r}_scaled; may be a strided concat slot (bts_decoder.py:99)
 *   out_ds    (B, h*r/d, w*r/d[,1]) = out_full[:, ::d, ::d]; NULL to skip; d = ds_stride must divide r
 *   upratio   r >= 1.  r in {2,4,8} with d in {0, r/2} run the vectorised kernels.
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API int btslpg_forward(const BtsTensor *coef, int upratio, BtsTensor *out_full,
                              BtsTensor *out_ds, int ds_stride, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Gradient of the above -- replaces what TF autodiff derives from custom_layers.py:49-56 and
 * from the strided slices bts_decoder.py:81,88.  Deterministic: the r x r patch reduction is a
 * fixed-order in-register sum, no atomics.
 *
 *   g_full (B, H, W[,1]) dL/d out_full, nullable;  g_ds (B, H/d, W/d[,1]) dL/d out_ds, nullable
 *   g_coef (B, h, w, 3)  dL/d coef (written, not accumulated)
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API int btslpg_backward(const BtsTensor *coef, const BtsTensor *g_full, const BtsTensor *g_ds,
                               int upratio, int ds_stride, BtsTensor *g_coef, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Several independent LPG layers in ONE launch (e.g. the three scales of the micro-benchmark, or
 * the three d(concat1) gradients of bts_decoder.py:99 that arrive together in backward).
 * Semantics identical to n calls of btslpg_forward / btslpg_backward.  n <= BTSLPG_MAX_MULTI.
 * ------------------------------------------------------------------------------------------- */
#define BTSLPG_MAX_MULTI 4
typedef struct {
    const BtsTensor *coef;
    int32_t upratio;
    int32_t ds_stride;
    BtsTensor *out_full;
    BtsTensor *out_ds; /* nullable */
} BtsLpgForwardArgs;

typedef struct {
    const BtsTensor *coef;
    const BtsTensor *g_full; /* nullable */
    const BtsTensor *g_ds;   /* nullable */
    int32_t upratio;
    int32_t ds_stride;
    BtsTensor *g_coef;
} BtsLpgBackwardArgs;

BTSLPG_API int btslpg_forward_multi(const BtsLpgForwardArgs *layers, int n, void *stream);
BTSLPG_API int btslpg_backward_multi(const BtsLpgBackwardArgs *layers, int n, void *stream);

/* ---------------------------------------------------------------------------------------------
 * reduction_NxN + LocalPlanarGuidance fused -- replaces bts_decoder.py:79-81 / 86-88 / 93-94:
 *   coef = sigmoid(Conv2D(3, 1x1, use_bias=False)(feat)) ; out_full = LPG_r(coef) ; out_ds = slice
 *
 *   feat     (B, h, w, C)   NHWC, channel stride 1
 *   kernel   [C][3] float32 == the Keras HWIO kernel (1,1,C,3) squeezed (any of ndim 2 or 4)
 *   coef_out (B, h, w, 3)   the sigmoid output, saved for backward (and for callers that need
 *                           the reduction tensor itself); required
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API int btslpg_reduce_forward(const BtsTensor *feat, const BtsTensor *kernel, int upratio,
                                     BtsTensor *coef_out, BtsTensor *out_full, BtsTensor *out_ds,
                                     int ds_stride, void *stream);

/* Gradient of the fused op (TF autodiff of bts_decoder.py:79-81 etc.):
 *   g_coef = LPG backward ; dz = g_coef * coef * (1 - coef)
 *   g_feat[b,i,j,c] = sum_k dz[b,i,j,k] * kernel[c,k]            (B,h,w,C), nullable to skip
 *   g_kernel[c,k]   = sum_{b,i,j} feat[b,i,j,c] * dz[b,i,j,k]    [C][3] float32, nullable to skip
 * g_kernel is reduced deterministically: per-CTA partial sums in `workspace`, summed in CTA
 * order by the last CTA to finish (no float atomics).  g_kernel may point into a flat gradient
 * bucket that is handed to ncclAllReduce afterwards (data-parallel training).
 * g_coef_out (B,h,w,3), nullable, additionally receives the LPG coefficient gradient.
 * workspace: btslpg_reduce_backward_workspace_bytes(npix, C) bytes of device memory; contents
 * need no initialisation beyond a one-time zeroing of its first 256 bytes (done by
 * the caller once after allocation; the kernel leaves them zero on exit).
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API size_t btslpg_reduce_backward_workspace_bytes(int64_t npix, int channels);
BTSLPG_API int btslpg_reduce_backward(const BtsTensor *feat, const BtsTensor *kernel, const BtsTensor *coef,
                                      const BtsTensor *g_full, const BtsTensor *g_ds, int upratio,
                                      int ds_stride, BtsTensor *g_feat, BtsTensor *g_kernel,
                                      BtsTensor *g_coef_out, void *workspace, size_t workspace_bytes,
                                      void *stream);

/* ---------------------------------------------------------------------------------------------
 * Decoder tail: the full-resolution passes after the last convolution.
 * All maps are contiguous, 16-byte aligned (B,H,W[,1]) tensors of one dtype (float32 / bfloat16,
 * float32 arithmetic, float64 cross-CTA sums).  `workspace` is btslpg_tail_workspace_bytes() bytes
 * of device memory whose first 256 bytes were zeroed once after allocation (every launch leaves
 * them zero).  Reductions are deterministic: fixed-order sums, no float atomics.
 *
 * btslpg_silog_forward -- replaces bts_decoder.py:102-103 (sigmoid activation of the last Conv2D
 * and the `depth_est` Lambda) and bts.py:27-41 (si_log_loss) in ONE pass:
 *   depth_est = sigmoid(logit) * max_depth                                    (written)
 *   mask = y_true > gt_threshold ; d = log(y_true + 1e-7) - log(depth_est + 1e-7) over the mask
 *   loss = sqrt(mean(d^2) - 0.85 * mean(d)^2) * 10                            (float32 device scalar)
 *   logit NULL  : the loss alone on a given depth_est (the reference's si_log_loss(y_true, y_pred))
 *   y_true NULL : depth_est only (inference); loss / workspace may be NULL
 * gt_threshold: 0.1 nyu / matterport, 1.0 kitti (bts.py:28).  An empty mask yields NaN like the
 * reference.  The forward leaves (n, mean d, variance term) in the workspace for the backward.
 *
 * btslpg_silog_backward -- TF autodiff of the above: g_out = g_loss * d loss / d logit (wrt_logit
 * != 0, through the sigmoid) or d loss / d depth_est (wrt_logit == 0).  g_loss: float32 device
 * scalar, NULL = 1.  `workspace` must be the forward's, untouched in between.
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API size_t btslpg_tail_workspace_bytes(void);
BTSLPG_API int btslpg_silog_forward(const BtsTensor *logit, const BtsTensor *y_true, float max_depth,
                                    float gt_threshold, BtsTensor *depth_est, BtsTensor *loss,
                                    void *workspace, size_t workspace_bytes, void *stream);
BTSLPG_API int btslpg_silog_backward(const BtsTensor *depth_est, const BtsTensor *y_true, float max_depth,
                                     float gt_threshold, const BtsTensor *g_loss, const void *workspace,
                                     size_t workspace_bytes, int wrt_logit, BtsTensor *g_out, void *stream);

/* btslpg_eval_metrics -- replaces custom_eval_metrics.py:24-88: the nine metrics, each of which the
 * reference computes with its own masked pass (pre_eval re-run every time), from ONE pass:
 *   mask = min_depth_eval < y_true < max_depth_eval
 *   pred = clip(where(isfinite(y_pred), y_pred, max_depth_eval), min_depth_eval, max_depth_eval)
 *   metrics[0..8] = silog, abs_rel, log10, rmse, sq_rel, rmse_log, d1, d2, d3 (the reference's list
 *   order, custom_eval_metrics.py:88); metrics[9] = number of valid pixels.  float32[>=10] device. */
BTSLPG_API int btslpg_eval_metrics(const BtsTensor *y_true, const BtsTensor *y_pred, float min_depth_eval,
                                   float max_depth_eval, BtsTensor *metrics, void *workspace,
                                   size_t workspace_bytes, void *stream);

/* btslpg_eval_metrics_png16 -- the same pass that additionally writes the 16-bit depth image bts_predict.py:140-141 saves
 * (`pred_depth * 65536 / args.max_depth` in float32, then `.astype(np.uint16)`), from the one read of y_pred (SURVEY 8(f)
 * N4, second half):
 *   png  uint16 (kDLUInt, 16 bits), as many elements as y_pred, contiguous, 16-byte aligned; NULL = metrics only
 *   png[i] = (uint16)(int32)trunc(y_pred[i] * 65536 / png_max_depth): numpy's C cast, so a prediction equal to max_depth
 *            wraps to 0 and NaN gives 0 exactly as the reference's line does (no clamping)
 *   y_true NULL (bts_predict.py has no ground truth): the scaling pass alone; metrics / workspace may then be NULL. */
BTSLPG_API int btslpg_eval_metrics_png16(const BtsTensor *y_true, const BtsTensor *y_pred, float min_depth_eval,
                                         float max_depth_eval, BtsTensor *metrics, float png_max_depth, BtsTensor *png,
                                         void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused activation + channel concat of NHWC tensors -- replaces bts_decoder.py:98-99
 *     upconv1 = Conv2D(..., activation='elu')(upsample1)        (the activation: pass the conv's linear output as `a`)
 *     concat1 = Concatenate(axis=3)([upconv1, depth_2x2_scaled, depth_4x4_scaled, depth_8x8_scaled])
 * and, with act = 0, the conv_block concat [upconv, skip, lpg_ds] of bts_decoder.py:42 -- one pass
 * instead of a separate ELU pass plus a re-copy of every input.  Channel order = argument order.
 *   a       (B,H,W,CA)  dense source; act = 1 applies ELU(alpha=1) to it on the way (0 = identity)
 *   a_subpixel  != 0: `a` is instead (B,H/2,W/2,4*CA), the output of the 3x3 upconv evaluated on the LOW-RES input with
 *           4*CA output channels ordered (row parity, column parity, channel) -- algebraically UpSampling2D(2,'nearest') +
 *           Conv2D(3x3) (bts_decoder.py:97-98) without the up-sampled tensor; the pixel shuffle is done by this kernel's
 *           addressing (and g_a of the backward comes back in the same layout)
 *   scale, shift  float32 [CA], nullable (together): per-channel affine on `a` AFTER the activation -- an
 *           inference-mode BatchNormalization folded in (bts_decoder.py:33-34 / :40-41: upconv -> elu -> BN -> concat);
 *           forward only (the backward entry point differentiates the un-affined form)
 *   b       (B,H,W,CB)  second dense source, nullable
 *   planes  n_planes (<= 3) single-channel maps (B,H,W[,1]), e.g. the LPG outputs
 *   pad_channels  0..7 zero channels appended so that the total is a multiple of 4: the consumer convolution gets
 *           matching zero input channels in its kernel and cuDNN skips its own padding copy of the whole tensor
 *   out     (B,H,W,CA+CB+n_planes+pad_channels)
 * All tensors contiguous, 16-byte aligned, one dtype (float32 / bfloat16).  CA, CB multiples of 4 (8 for
 * bfloat16) take the vectorised path; other channel counts are accepted and run with scalar accesses.
 *
 * Backward (TF autodiff of the above): g_a = g_out[..., :CA] * elu'(.) with elu' taken from the saved
 * output y (required iff act = 1: y > 0 ? 1 : y + 1), g_b and g_planes[k] are slices of g_out.
 * g_b / g_planes[k] may be NULL to skip them; the pad channels' gradient is dropped.
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API int btslpg_concat_forward(const BtsTensor *a, int a_subpixel, int act, const BtsTensor *scale,
                                     const BtsTensor *shift, const BtsTensor *b, const BtsTensor *const *planes,
                                     int n_planes, int pad_channels, BtsTensor *out, void *stream);
BTSLPG_API int btslpg_concat_backward(const BtsTensor *g_out, const BtsTensor *y, int act, BtsTensor *g_a, int a_subpixel,
                                      BtsTensor *g_b, BtsTensor *const *g_planes, int n_planes, int pad_channels,
                                      void *stream);

/* ---------------------------------------------------------------------------------------------
 * TRAINING-mode conv block glue -- bts_decoder.py:30-44:
 *     upconv = Conv2D(nf, 3, activation='elu')(upsample) ; upconv = BatchNormalization(momentum=0.99, epsilon=1.1e-5)(upconv, training)
 *     concat = Concatenate(axis=3)([upconv, skip(, lpg)])
 * as a statistics pass plus the fused concat pass, forward and backward (the framework: ELU, BatchNorm, cat = three read+write
 * passes each way).  `pack` is a float32 [8][C] device buffer owned by the caller and shared by the four calls of a block:
 * {scale, shift, mean, std, 1/gamma, beta, c1, c2}.
 *   btslpg_bn_elu_stats            raw (B,H,W,C) = the convolution's LINEAR output; act = 1: statistics of elu(raw).  Writes
 *                                  pack[0..5] (scale = gamma / sqrt(var + eps) with the BIASED batch variance, shift = beta - mean * scale)
 *                                  and, when given, the moving averages: running = (1 - momentum) * running + momentum * batch value
 *                                  (momentum = 1 - Keras' 0.99; unbiased variance, as fused batch norm does).  Feed pack[0], pack[1]
 *                                  to btslpg_concat_forward as scale / shift.
 *   btslpg_bn_elu_backward_stats   g_out, y: gradient and saved output of the concat (B,H,W,CT), the block's `channels` first.
 *                                  Writes g_gamma = sum g * xhat, g_beta = sum g (xhat = (y - beta) / gamma: recovered from the output,
 *                                  nothing else is kept alive) and pack[6..7] = their means.
 *   btslpg_concat_backward_bn      btslpg_concat_backward with the pack: g_a = scale * (g - c1 - xhat * c2) * elu'(elu).
 * float32, C a power of two in [4, 1024], CT a multiple of 4.  Deterministic (fixed-order sums, no atomics).  Because xhat is recovered
 * from the output, a channel whose gamma is exactly 0 has no recoverable xhat (1/gamma): the backward then yields non-finite values for
 * that channel.  Keras initialises gamma = 1 and nothing in the reference's training drives it to exactly 0; a caller that prunes
 * channels by zeroing gamma must use the framework's layers for that block (the host code's fallback path).
 * workspace: btslpg_bn_workspace_bytes(C) bytes, 16-byte aligned, no initialisation needed.
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API size_t btslpg_bn_workspace_bytes(int channels);
BTSLPG_API int btslpg_bn_elu_stats(const BtsTensor *raw, int act, const BtsTensor *gamma, const BtsTensor *beta, BtsTensor *running_mean,
                                   BtsTensor *running_var, float momentum, float eps, BtsTensor *pack, void *workspace,
                                   size_t workspace_bytes, void *stream);
BTSLPG_API int btslpg_bn_elu_backward_stats(const BtsTensor *g_out, const BtsTensor *y, int channels, BtsTensor *pack, BtsTensor *g_gamma,
                                            BtsTensor *g_beta, void *workspace, size_t workspace_bytes, void *stream);
BTSLPG_API int btslpg_concat_backward_bn(const BtsTensor *g_out, const BtsTensor *y, int act, const BtsTensor *bn_pack, BtsTensor *g_a,
                                         int a_subpixel, BtsTensor *g_b, BtsTensor *const *g_planes, int n_planes, int pad_channels,
                                         void *stream);

/* ---------------------------------------------------------------------------------------------
 * Weight gradient of a full-resolution 3x3 convolution (stride 1, padding='same', no bias) on the tcgen05 tensor cores
 *  -- the d kernel of upconv1 / iconv1 (bts_decoder.py:98, :100), the part of their backward the library runs at 4-9
 * times its traffic floor:
 *     g_kernel[ky][kx][ci][co] = sum_{b,y,x} x[b, y+ky-1, x+kx-1, ci] * g[b, y, x, co]          (x = 0 outside the image)
 *   x         (B,H,W,Cin)  float32 contiguous NHWC, Cin a multiple of 4 in [4, 256]: the convolution's INPUT
 *   g         (B,H,W,Cout) float32 contiguous NHWC, Cout a multiple of 4 in [4, 128]: gradient of its (linear) output
 *             (channels beyond 64 on either side are handled in passes of 64 x 64 that stage their operands again)
 *   g_kernel  float32 [3][3][Cin][Cout], the Keras HWIO layout of layer.kernel
 * TF32 operands (the tensor core ignores the low 13 mantissa bits of the float32 inputs), float32 accumulation; deterministic
 * (per-CTA partials summed in a fixed order).  workspace: btslpg_conv3x3_wgrad_workspace_bytes(Cin, Cout) bytes, 16-byte aligned; like
 * the other workspaces of this library it may be shared with them: its first 256 bytes (their counters) are never touched.
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API size_t btslpg_conv3x3_wgrad_workspace_bytes(int cin, int cout);
BTSLPG_API int btslpg_conv3x3_wgrad(const BtsTensor *x, const BtsTensor *g, BtsTensor *g_kernel, void *workspace, size_t workspace_bytes,
                                    void *stream);

/* ---------------------------------------------------------------------------------------------
 * TRAINING-mode BatchNormalization (+ ReLU) over channel slices of NHWC buffers -- the glue of the DenseASPP
 * (bts_decoder.py:46-54 dense_aspp_block, :61-76) when is_training: [Concatenate ->] BatchNormalization(training) -> ReLU in front of
 * every convolution, on one (B,h,w,896) buffer that the blocks append to.  Batch statistics are per channel, so the moments of a
 * channel are taken ONCE when it is appended and shared by every later BatchNormalization over a concat that contains it.
 *   btslpg_bn_moments             x (B,H,W,C) channel slice -> mean[C], var[C] (biased), float32
 *   btslpg_bn_fold                moments + gamma / beta -> scale = gamma * rstd, shift = beta - mean * scale, rstd = 1/sqrt(var + eps);
 *                                 updates the moving averages when given (torch momentum = 1 - Keras' 0.99; unbiased variance,
 *                                 `count` values per channel).  The forward is then btslpg_affine_act(src, scale, shift, ReLU, dst).
 *   btslpg_bn_act_backward_stats  y = act(x * scale + shift); gm = g * [y > 0] (relu) (+ g2, nullable: a gradient that reaches the
 *                                 NORMALISED value directly): g_beta = sum gm, g_gamma = sum gm * xhat, xhat = (x - mean) * rstd
 *   btslpg_bn_act_backward        dst (+)= scale * (gm - g_beta / n - xhat * g_gamma / n): d loss / d x, written or ACCUMULATED into a
 *                                 slice of the shared gradient buffer (the concat's backward is that accumulation); dst_init
 *                                 (nullable, accumulate = 0): dst = dst_init + value, the first contribution to a slice whose
 *                                 upstream part lives in another tensor (no copy of the upstream gradient is made)
 * Every tensor argument is float32, channel stride 1, uniformly strided 16-byte aligned pixels, C a multiple of 4 in [4, 1024];
 * per-channel vectors are contiguous float32 [C], 16-byte aligned.  Deterministic (fixed-order sums, no atomics).
 * workspace: btslpg_bn_workspace_bytes(C) bytes, 16-byte aligned, no initialisation needed.
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API int btslpg_bn_moments(const BtsTensor *x, BtsTensor *mean, BtsTensor *var, void *workspace, size_t workspace_bytes, void *stream);
BTSLPG_API int btslpg_bn_fold(const BtsTensor *mean, const BtsTensor *var, const BtsTensor *gamma, const BtsTensor *beta,
                              BtsTensor *running_mean, BtsTensor *running_var, float momentum, float eps, int64_t count, BtsTensor *scale,
                              BtsTensor *shift, BtsTensor *rstd, void *stream);
BTSLPG_API int btslpg_bn_act_backward_stats(const BtsTensor *g, const BtsTensor *g2, const BtsTensor *x, const BtsTensor *scale,
                                            const BtsTensor *shift, const BtsTensor *mean, const BtsTensor *rstd, int relu,
                                            BtsTensor *g_gamma, BtsTensor *g_beta, void *workspace, size_t workspace_bytes, void *stream);
BTSLPG_API int btslpg_bn_act_backward(const BtsTensor *g, const BtsTensor *g2, const BtsTensor *x, const BtsTensor *scale,
                                      const BtsTensor *shift, const BtsTensor *mean, const BtsTensor *rstd, const BtsTensor *g_gamma,
                                      const BtsTensor *g_beta, int relu, BtsTensor *dst, int accumulate, const BtsTensor *dst_init,
                                      void *stream);

/* ---------------------------------------------------------------------------------------------
 * Nearest-neighbour x2 up-sampling of an NHWC map -- replaces the
 * `layers.UpSampling2D(size=2, interpolation='nearest')` in front of every upconv of the decoder
 * (bts_decoder.py:31, :38, :97): out[b, y, x, :] = in[b, y/2, x/2, :].
 *   in  (B,h,w,C) ; out (B,2h,2w,C) ; contiguous, one dtype (float32 / bfloat16).  16-byte aligned tensors whose
 *   pixels are a multiple of 16 bytes take the vectorised kernels, anything else a scalar one.
 * Backward: g_in[b,y,x,:] = sum of the four g_out pixels it was copied to, added in a fixed order.
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API int btslpg_upsample2x_forward(const BtsTensor *in, BtsTensor *out, void *stream);
BTSLPG_API int btslpg_upsample2x_backward(const BtsTensor *g_out, BtsTensor *g_in, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Per-channel affine + activation copy between channel slices of NHWC tensors -- the glue of
 * the DenseASPP block (bts_decoder.py:46-54, :61-76): with one (B,h,w,896) buffer that the blocks append to,
 * the reference's Concatenate + BatchNormalization + ReLU in front of every 1x1 conv (three passes over a growing
 * map) become one:  dst[b,y,x,c] = act(src[b,y,x,c] * scale[c] + shift[c]).
 *   src, dst  (B,H,W,C), channel stride 1, uniformly strided pixels (e.g. buffer[..., c0:c0+C]); may alias (in place)
 *   scale, shift  float32 [C], nullable together (an inference-mode BatchNormalization folded to an affine)
 *   act  0 none, 1 ELU(alpha=1), 2 ReLU
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API int btslpg_affine_act(const BtsTensor *src, const BtsTensor *scale, const BtsTensor *shift, int act,
                                 BtsTensor *dst, void *stream);
/* The same pass with ONE side in sub-grid ("space to batch") form: split_side 1 = src, 2 = dst (0: btslpg_affine_act).  That side is a
 * (B*s*s, H/s, W/s, C) tensor whose image (b*s + i)*s + j is the sub-grid [i::s, j::s] of image b of the other side's (B,H,W,C)
 * pixels.  The DenseASPP convolutions of rate 18 / 24 (bts_decoder.py:53) run as 2 x 2 sub-grids of rate 9 / 12 (the library's fast
 * kernels stop below rate 18); their input is produced and their output consumed by this pass anyway, so the re-ordering is free. */
BTSLPG_API int btslpg_affine_act_split(const BtsTensor *src, const BtsTensor *scale, const BtsTensor *shift, int act, BtsTensor *dst,
                                       int split_side, int s, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Backward of the decoder's last convolution -- bts_decoder.py:102
 *     Conv2D(1, kernel_size=3, strides=1, padding='same', use_bias=False)(iconv1)
 * both gradients in ONE pass (library convolutions need 3.7 ms for it at B = 32, 480x640: a tensor-core weight
 * gradient with a 288-element output plus two layout conversions; the traffic floor is 0.4 ms):
 *   x        (B,H,W,C)  the layer input, C = 16 or 32 (F/16), contiguous NHWC
 *   kernel   float32, 9*C elements: the Keras HWIO kernel (3,3,C,1) as it lies in memory ([tap][c])
 *   g_out    (B,H,W[,1]) gradient of the layer's (pre-activation) output
 *   act_in   0: x is the layer input; 1: x is the PRE-activation of iconv1 (bts_decoder.py:100) and the layer input is
 *            elu(x), as in btslpg_depthconv_forward(act_in = 1): g_kernel is taken against elu(x) and g_x is the gradient
 *            with respect to x itself (times elu'(x)) -- the framework's separate ELU forward and backward passes disappear
 *   g_x      (B,H,W,C)  d loss / d x, nullable
 *   g_kernel float32 [9*C] d loss / d kernel in the same layout, nullable; reduced deterministically through
 *            `workspace` (btslpg_depthconv_backward_workspace_bytes(C) bytes, first 256 zeroed once)
 * Exact float32 arithmetic.
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API size_t btslpg_depthconv_backward_workspace_bytes(int channels);
BTSLPG_API int btslpg_depthconv_backward(const BtsTensor *x, const BtsTensor *kernel, const BtsTensor *g_out, int act_in, BtsTensor *g_x,
                                         BtsTensor *g_kernel, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Forward of the same layer with its neighbours folded in -- bts_decoder.py:100-103
 *     iconv1           = Conv2D(F/16, 3, activation='elu')(concat1)          only the ACTIVATION (act_in = 1)
 *     depth_est_scaled = Conv2D(1, 3, padding='same', use_bias=False, activation='sigmoid')(iconv1)
 *     depth_est        = depth_est_scaled * max_depth                        (act_out = 1, out_scale = max_depth)
 * in ONE pass over the raw output of iconv1's convolution (the library path: ELU pass, NHWC->NCHW conversion,
 * convolution, sigmoid pass -- about 1.1 ms at B = 32, 480x640 against a 0.1 ms traffic floor):
 *   x        (B,H,W,C)  C = 16 or 32, contiguous NHWC; the PRE-activation map when act_in = 1
 *   kernel   float32, 9*C elements, Keras HWIO (3,3,C,1) as it lies in memory ([tap][c])
 *   act_in   0 none, 1 ELU(alpha = 1) applied to x inside the sum (padding='same' pads the activated map with zeros)
 *   act_out  0: y = the convolution (the logit; training / fused loss), 1: y = sigmoid(conv) * out_scale
 *   y        (B,H,W[,1]) contiguous, same dtype as x
 * float32-accurate (3xTF32 split on the tensor cores, float32 accumulation; csrc/depthconv_kernels.cuh has the error
 * bounds), fixed order, bit-reproducible.
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API int btslpg_depthconv_forward(const BtsTensor *x, const BtsTensor *kernel, int act_in, int act_out, float out_scale,
                                        BtsTensor *y, void *stream);

/* ---------------------------------------------------------------------------------------------
 * iconv1 as a tcgen05 implicit GEMM that reads the SOURCES of concat1 (SURVEY 8(a) a10 + 8(f) N1) -- replaces, in inference,
 * bts_decoder.py:98-100:
 *     upconv1 = Conv2D(F/16, 3, activation='elu')(upsample1)                        only the ACTIVATION (pass the linear output)
 *     concat1 = Concatenate(axis=3)([upconv1, depth_2x2_scaled, depth_4x4_scaled, depth_8x8_scaled])
 *     iconv1  = Conv2D(F/16, 3, strides=1, padding='same', use_bias=False, activation='elu')(concat1)
 * concat1 is never materialised: the kernel stages elu(a) and the three planes into shared memory as the A operand of
 * tcgen05.mma (TMEM accumulators), nine taps as nine address offsets into the same staged rows.
 *   a        (B,H,W,NF) the LINEAR output of upconv1's convolution, NF = F/16 = 16 or 32; or, a_subpixel != 0,
 *            (B,H/2,W/2,4*NF) in the sub-pixel layout of btslpg_concat_forward
 *   planes   3 single-channel maps (B,H,W[,1]): depth_2x2_scaled, depth_4x4_scaled, depth_8x8_scaled (concat order)
 *   kernel   float32, the Keras HWIO kernel (3,3,NF+3,NF) as it lies in memory
 *   act_out  0: out = the convolution (what btslpg_depthconv_forward(act_in = 1) consumes), 1: out = elu(convolution)
 *   out      (B,H,W,NF)
 * float32 tensors, contiguous, 32-byte aligned.  Arithmetic: TF32 operands (rounded to nearest), float32 accumulation --
 * the precision of the library convolution it replaces (cuDNN with TF32 enabled, the framework default); tolerance against
 * a float64 evaluation: 3e-3 of the largest output magnitude.  Fixed summation order: bit-reproducible.
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API int btslpg_iconv1_forward(const BtsTensor *a, int a_subpixel, const BtsTensor *const *planes, const BtsTensor *kernel,
                                     int act_out, BtsTensor *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * The optimizer step of the data-parallel training loop as ONE pass over flat float32 buffers -- replaces
 *   custom_optimizers.py:47-59  AdamW._resource_apply_dense: `var -= lr * (l1*sign(var) + l2*var)` BEFORE the Adam update
 *   tf.keras.optimizers.Adam    alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t) ; m += (g - m)(1 - beta1) ;
 *                               v += (g*g - v)(1 - beta2) ; var -= alpha * m / (sqrt(v) + epsilon)      (t = iterations + 1)
 *   bts_train.py:125-131        lr(step) = (lr_start - lr_end) * (1 - min(step, total_steps)/total_steps)^power + lr_end,
 *                               evaluated at the 0-based global step (custom_callbacks.py:46-50); lr_start already carries
 *                               the reference's "x num_replicas" (bts_train.py:125)
 *   MirroredStrategy's gradient average: grad_scale = 1/N multiplies the all-reduced (summed) gradient on the way in.
 *   param, grad, m, v   float32, same number of elements, contiguous, 16-byte aligned (slices of flat buckets)
 *   state   >= 4 32-bit words of device memory: [0] int32 number of completed updates (the kernel reads the step from
 *           here, so a CUDA graph replay advances without the host), [1] float32 lr of the last update (informational)
 *   advance != 0: this call is the last chunk of the step -- state[0] += 1 afterwards (a second 1-thread launch)
 *   cfg->zero_grad != 0: grad is overwritten with zeros as it is consumed (no separate zeroing pass before the next step)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    float lr_start, lr_end;   /* lr_end == lr_start or total_steps <= 0: constant learning rate */
    int64_t total_steps;
    float power;              /* 0.9 in the reference */
    float beta1, beta2;       /* Keras defaults 0.9, 0.999 */
    float epsilon;            /* --adam_eps, 1e-3 (bts_train.py:86) */
    float l1, l2;             /* decoupled decay of this group (decoder variables: 0, 0 -- bts.py:107 decays the encoder only) */
    float grad_scale;
    int32_t zero_grad;
} BtsAdamConfig;
BTSLPG_API int btslpg_adam_step(BtsTensor *param, BtsTensor *grad, BtsTensor *m, BtsTensor *v, BtsTensor *state,
                                const BtsAdamConfig *cfg, int advance, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Introspection used by bench.py ("gpu_launches") and the tests: number of kernel launches issued
 * through this library (process-wide) since the last reset, and the name of the
 * kernel variant the last call dispatched to (e.g. "lpg_fwd_vec<f32,r8,px1,ds4>").
 * ------------------------------------------------------------------------------------------- */
BTSLPG_API uint64_t btslpg_launch_count(void);
BTSLPG_API void btslpg_reset_launch_count(void);
BTSLPG_API const char *btslpg_last_kernel(void);

/* Device-wide barrier (cudaDeviceSynchronize on `device_id`).  NOT used by any torch-hosted path: it exists for hosts that
 * cannot hand the library their own stream (the experimental tf.py_function binding, tf_adapter.py), which must order the
 * legacy-stream launches against the framework's non-blocking streams. */
BTSLPG_API int btslpg_device_synchronize(int device_id);

/* Tuning knobs for experiments (threads per block of the vectorised kernels; 0 = default).
 * btslpg_set_tuning keys: 0 forward block threads, 1 backward block threads, 2 float32 r=8 patch rows
 * per lane (2/4/8), 3 float32 r=4 coarse pixels per thread (1/2), 9 last-convolution forward (0 tensor-core
 * phase 1 with the 3xTF32 split, 1 FP32 pipe), 10 concat forward (0 shared-memory-staged kernel, 1 the
 * chunked kernel where every 16-byte output chunk has one source: an experiment, measured slower).  Results do not depend on them beyond float32 rounding,
 * except that the r=8 backward sum is associated per lane group (still deterministic). */
BTSLPG_API void btslpg_set_block_threads(int fwd_threads, int bwd_threads);
BTSLPG_API void btslpg_set_tuning(int key, int value);

#ifdef __cplusplus
}
#endif
#endif /* BTSLPG_H_ */
