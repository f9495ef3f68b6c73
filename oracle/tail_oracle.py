"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64 unless told otherwise) of the decoder
tail rows of SURVEY 8(f): N2 `si_log_loss` (/root/reference/bts.py:27-41) behind the final activation
(/root/reference/bts_decoder.py:102-103) and N4 the eval metrics
(/root/reference/custom_eval_metrics.py:21-88).

Pinned by tests/golden/tail_*.npz, which tests/golden/make_golden.py produces by executing the
UNMODIFIED reference files over oracle/tf_shim (same status as the LPG oracle: pinned to the
reference's own source run over a stand-in runtime; TensorFlow's kernel rounding is not pinned).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np

EPS = 1e-7                                                   # K.epsilon()
GT_TH = {"nyu": 0.1, "kitti": 1.0, "matterport": 0.1}        # bts.py:28
METRIC_NAMES = ("silog", "abs_rel", "log10", "rmse", "sq_rel", "rmse_log", "d1", "d2", "d3")   # custom_eval_metrics.py:88


def depth_est(logit, max_depth, dtype=np.float64):
    """bts_decoder.py:102-103: sigmoid activation of the last Conv2D, then the `depth_est` Lambda."""
    z = np.asarray(logit, dtype)
    return (dtype(1) / (dtype(1) + np.exp(-z))) * dtype(max_depth)


def si_log_loss(y_true, y_pred, threshold, dtype=np.float64):
    """bts.py:31-38.  Returns (loss, stats) with stats = (n, mean d, mean d^2 - 0.85 mean(d)^2)."""
    yt, yp = np.asarray(y_true, dtype).ravel(), np.asarray(y_pred, dtype).ravel()
    mask = yt > dtype(threshold)                                            # :32
    d = np.log(yt[mask] + dtype(EPS)) - np.log(yp[mask] + dtype(EPS))       # :34-37
    with np.errstate(invalid="ignore", divide="ignore"):
        m1 = d.mean() if d.size else dtype(np.nan)
        m2 = (d * d).mean() if d.size else dtype(np.nan)
        var = m2 - dtype(0.85) * m1 * m1
        loss = np.sqrt(var) * dtype(10.0)                                   # :38
    return loss, (d.size, m1, var)


def si_log_loss_grad(y_true, y_pred, threshold, g_loss=1.0, max_depth=None):
    """Analytic gradient of the loss (float64): with respect to y_pred, or -- when max_depth is given and
    y_pred = sigmoid(z)*max_depth -- with respect to the logit z (chain rule through bts_decoder.py:102-103)."""
    yt, yp = np.asarray(y_true, np.float64), np.asarray(y_pred, np.float64)
    mask = yt > threshold
    _, (n, m1, var) = si_log_loss(yt, yp, threshold)
    g = np.zeros_like(yp)
    d = np.log(yt[mask] + EPS) - np.log(yp[mask] + EPS)
    gd = g_loss * 10.0 * (d - 0.85 * m1) / (n * np.sqrt(var))
    g[mask] = -gd / (yp[mask] + EPS)
    if max_depth is not None:
        g = g * yp * (1.0 - yp / max_depth)
    return g


def pre_eval(y_true, y_pred, min_depth_eval, max_depth_eval, dtype=np.float64):
    """custom_eval_metrics.py:24-42 (the crop helper :27-37 is dead code in the reference)."""
    yt, yp = np.asarray(y_true, dtype).ravel(), np.asarray(y_pred, dtype).ravel()
    mask = (yt < dtype(max_depth_eval)) & (yt > dtype(min_depth_eval))      # :39
    yp = np.where(np.isfinite(yp), yp, dtype(max_depth_eval))               # :40
    yp = np.clip(yp, dtype(min_depth_eval), dtype(max_depth_eval))          # :41
    return yt[mask], yp[mask]                                               # :42


def eval_metrics(y_true, y_pred, min_depth_eval, max_depth_eval, dtype=np.float64):
    """custom_eval_metrics.py:44-88 -> dict name -> value, plus n_valid."""
    gt, pred = pre_eval(y_true, y_pred, min_depth_eval, max_depth_eval, dtype)
    ratio = np.maximum(gt / pred, pred / gt)
    d = np.log(gt) - np.log(pred)
    out = {
        "d1": (ratio < 1.25).astype(dtype).mean(),                         # :47
        "d2": (ratio < 1.25 ** 2).astype(dtype).mean(),                    # :51
        "d3": (ratio < 1.25 ** 3).astype(dtype).mean(),                    # :55
        "rmse": np.sqrt(((gt - pred) ** 2).mean()),                        # :60
        "rmse_log": np.sqrt((d ** 2).mean()),                              # :64-65
        "abs_rel": (np.abs(gt - pred) / gt).mean(),                        # :70
        "sq_rel": (((gt - pred) ** 2) / gt).mean(),                        # :74
        "silog": np.sqrt((d ** 2).mean() - d.mean() ** 2) * 100,           # :79-80
        "log10": np.abs(d).mean() / np.log(10.0),                          # :85-86
        "n_valid": gt.size,
    }
    return out


def concat_elu(a, planes=(), b=None, act=False, dtype=np.float64, pad=0, scale=None, shift=None):
    """bts_decoder.py:98-99 (activation='elu' of upconv1, then Concatenate(axis=3)) and :42 (act=False):
    channel order = [a, b, *planes] (+ `pad` zero channels).  Keras elu: x > 0 ? x : exp(x) - 1.
    scale/shift: an inference-mode BatchNormalization after the activation (bts_decoder.py:33-34, :40-41)."""
    a = np.asarray(a, dtype)
    first = np.where(a > 0, a, np.expm1(np.minimum(a, 0))) if act else a
    if scale is not None:
        first = first * np.asarray(scale, dtype) + np.asarray(shift, dtype)
    parts = [first]
    if b is not None:
        parts.append(np.asarray(b, dtype))
    parts += [np.asarray(p, dtype).reshape(a.shape[:3] + (1,)) for p in planes]
    if pad:
        parts.append(np.zeros(a.shape[:3] + (pad,), dtype))
    return np.concatenate(parts, axis=3)


def concat_elu_grad(g_out, a, ca, cb, n_planes, act=False):
    """Gradients of concat_elu with respect to (a, b, planes) for upstream g_out (float64)."""
    g = np.asarray(g_out, np.float64)
    a = np.asarray(a, np.float64)
    g_a = g[..., :ca] * (np.where(a > 0, 1.0, np.exp(np.minimum(a, 0))) if act else 1.0)
    g_b = g[..., ca:ca + cb] if cb else None
    g_p = [g[..., ca + cb + k:ca + cb + k + 1] for k in range(n_planes)]
    return g_a, g_b, g_p


def upsample2x(x):
    """layers.UpSampling2D(size=2, interpolation='nearest') (bts_decoder.py:31, :38, :97) on NHWC:
    out[b, y, x] = in[b, y // 2, x // 2] (Keras implements it as repeat_elements along H then W)."""
    return np.repeat(np.repeat(np.asarray(x), 2, axis=1), 2, axis=2)


def upsample2x_grad(g_out):
    g = np.asarray(g_out, np.float64)
    B, H, W, C = g.shape
    return g.reshape(B, H // 2, 2, W // 2, 2, C).sum(axis=(2, 4))


def affine_act(x, scale=None, shift=None, act=0):
    """DenseASPP glue (bts_decoder.py:47-49, :51-52): inference BatchNormalization as an affine, then 0 none / 1 ELU / 2 ReLU."""
    x = np.asarray(x, np.float64)
    if scale is not None:
        x = x * np.asarray(scale, np.float64) + np.asarray(shift, np.float64)
    if act == 1:
        return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
    if act == 2:
        return np.maximum(x, 0)
    return x


def depthconv_forward(x, w9c):
    """bts_decoder.py:102 Conv2D(1, 3, padding='same', use_bias=False): y[p] = sum_{t,c} x[p + t][c] * w[t][c], float64.
    w9c: the HWIO kernel (3,3,C,1) flattened to (9, C)."""
    x = np.asarray(x, np.float64)
    B, H, W, C = x.shape
    w = np.asarray(w9c, np.float64).reshape(3, 3, C)
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    y = np.zeros((B, H, W))
    for ky in range(3):
        for kx in range(3):
            y += (xp[:, ky:ky + H, kx:kx + W, :] * w[ky, kx]).sum(-1)
    return y[..., None]


def depth_tail_forward(x_raw, w9c, act_in=True, max_depth=None):
    """bts_decoder.py:100-103 downstream of iconv1's convolution, float64:
        iconv1 = elu(x_raw)                                  (:100 activation='elu'; Keras elu = x > 0 ? x : expm1(x))
        logit  = Conv2D(1, 3, padding='same', use_bias=False)(iconv1)         (:102, zero padding of the ACTIVATED map)
        depth  = sigmoid(logit) * max_depth                  (:102 activation='sigmoid', :103) -- when max_depth is given
    """
    x = np.asarray(x_raw, np.float64)
    if act_in:
        x = np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
    y = depthconv_forward(x, w9c)
    if max_depth is not None:
        y = max_depth / (1.0 + np.exp(-y))
    return y


def depth_tail_backward(x_raw, w9c, g_out):
    """Gradients of depth_tail_forward(x_raw, w9c, act_in=True, max_depth=None) (bts_decoder.py:100-102), float64:
    (d loss / d x_raw, d loss / d kernel (9, C)) for g_out = d loss / d logit.  elu'(x) = 1 for x > 0, exp(x) otherwise."""
    x = np.asarray(x_raw, np.float64)
    xe = np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
    g_xe, g_w = depthconv_backward(xe, w9c, g_out)
    return g_xe * np.where(x > 0, 1.0, np.exp(np.minimum(x, 0))), g_w


def depthconv_backward(x, w9c, g_out):
    """Gradients of depthconv_forward: (g_x (B,H,W,C), g_w (9, C)), float64."""
    x = np.asarray(x, np.float64)
    B, H, W, C = x.shape
    w = np.asarray(w9c, np.float64).reshape(3, 3, C)
    g = np.asarray(g_out, np.float64).reshape(B, H, W)
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    gxp = np.zeros_like(xp)
    gw = np.zeros((3, 3, C))
    for ky in range(3):
        for kx in range(3):
            gw[ky, kx] = (g[..., None] * xp[:, ky:ky + H, kx:kx + W, :]).sum(axis=(0, 1, 2))
            gxp[:, ky:ky + H, kx:kx + W, :] += g[..., None] * w[ky, kx]
    return gxp[:, 1:-1, 1:-1, :], gw.reshape(9, C)


def iconv1_forward(a_raw, planes, hwio, act_out=False):
    """bts_decoder.py:98-100 in float64: concat1 = [elu(a_raw), d2, d4, d8] (:98-99), iconv1 = Conv2D(NF, 3, padding='same',
    use_bias=False)(concat1) with the Keras HWIO kernel (3,3,NF+3,NF), then ELU if act_out (:100)."""
    cat = concat_elu(a_raw, planes, act=True)                          # (B,H,W,NF+3)
    B, H, W, C = cat.shape
    w = np.asarray(hwio, np.float64).reshape(3, 3, C, -1)
    xp = np.pad(cat, ((0, 0), (1, 1), (1, 1), (0, 0)))
    y = np.zeros((B, H, W, w.shape[-1]))
    for ky in range(3):
        for kx in range(3):
            y += xp[:, ky:ky + H, kx:kx + W, :] @ w[ky, kx]
    if act_out:
        y = np.where(y > 0, y, np.expm1(np.minimum(y, 0)))
    return y


def conv_block_glue(raw, skip, planes, gamma, beta, eps, pad=0):
    """bts_decoder.py:32-42 in TRAINING mode, float64: elu -> BatchNormalization with the batch's own (biased) statistics ->
    Concatenate([., skip, *planes]) (+ zero pad channels).  Returns (concat, mean, biased variance)."""
    x = np.asarray(raw, np.float64)
    e = np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
    mean = e.mean(axis=(0, 1, 2))
    var = e.var(axis=(0, 1, 2))
    up = (e - mean) / np.sqrt(var + eps) * np.asarray(gamma, np.float64) + np.asarray(beta, np.float64)
    parts = [up, np.asarray(skip, np.float64)] + [np.asarray(p, np.float64).reshape(x.shape[:3] + (1,)) for p in planes]
    if pad:
        parts.append(np.zeros(x.shape[:3] + (pad,)))
    return np.concatenate(parts, axis=3), mean, var


def conv3x3_wgrad(x, g, tf32_operands=False):
    """d kernel (HWIO) of Conv2D(3x3, strides 1, padding='same', use_bias=False) -- bts_decoder.py:98, :100 -- in float64:
    dW[ky][kx][ci][co] = sum_{b,y,x} x[b, y+ky-1, x+kx-1, ci] * g[b, y, x, co], x zero outside the image.
    tf32_operands: cut both operands to TF32 first (the tensor core ignores the low 13 mantissa bits of float32 inputs)."""
    x = np.asarray(x, np.float32)
    g = np.asarray(g, np.float32)
    if tf32_operands:
        x = (x.view(np.int32) & np.int32(-8192)).view(np.float32)
        g = (g.view(np.int32) & np.int32(-8192)).view(np.float32)
    x, g = x.astype(np.float64), g.astype(np.float64)
    B, H, W, _ = x.shape
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    out = np.empty((3, 3, x.shape[3], g.shape[3]))
    for ky in range(3):
        for kx in range(3):
            out[ky, kx] = np.einsum("bhwi,bhwo->io", xp[:, ky:ky + H, kx:kx + W, :], g)
    return out


def bn_relu_backward(g, x, gamma, beta, eps, g2=None):
    """Backward of y = relu(BatchNormalization(x)) with the batch's own statistics (bts_decoder.py:47-48, :51-52 with is_training), float64:
    returns (d x, d gamma, d beta).  g2: a gradient that reaches the normalised value directly (no ReLU)."""
    x = np.asarray(x, np.float64)
    g = np.asarray(g, np.float64)
    ax = tuple(range(x.ndim - 1))
    n = x.size // x.shape[-1]
    mean, var = x.mean(axis=ax), x.var(axis=ax)
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean) * rstd
    z = xhat * gamma + beta
    gm = np.where(z > 0, g, 0.0) + (0.0 if g2 is None else np.asarray(g2, np.float64))
    d_beta, d_gamma = gm.sum(axis=ax), (gm * xhat).sum(axis=ax)
    d_x = gamma * rstd * (gm - d_beta / n - xhat * d_gamma / n)
    return d_x, d_gamma, d_beta
