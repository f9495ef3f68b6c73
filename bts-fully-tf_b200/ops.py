"""Operator layer above the C ABI: thin, allocation + autograd glue only.

Every function here ends in a call into libbtslpg.so on the tensor's CUDA device and the
caller's current stream; nothing is computed in Python or PyTorch.  Layout follows the
reference (Keras channels_last): coefficients (B,h,w,3), depth maps (B,H,W,1).

Reference lines replaced:
  lpg_forward / LpgFunction     custom_layers.py:47-56 (+ the slices bts_decoder.py:81,88)
  reduce_lpg / ReduceLpgFunction bts_decoder.py:79-81, 86-88, 93-94
  lpg_forward_multi / lpg_backward_multi   the three layers of one decoder in one launch
  depth_silog / si_log_loss      bts_decoder.py:102-103 + bts.py:27-41
  eval_metrics                   custom_eval_metrics.py:24-88
  concat_nhwc                    bts_decoder.py:98-99 (ELU of upconv1 + concat1) and :42
  upsample2x_nhwc                bts_decoder.py:31, :38, :97 UpSampling2D(size=2, 'nearest')
  affine_act                     bts_decoder.py:46-76 DenseASPP glue: BN affine + ReLU over channel slices
  depth_conv                     bts_decoder.py:102 last Conv2D(1, 3x3): fused data + weight gradient
  iconv1_forward                 bts_decoder.py:98-100 ELU + concat1 + iconv1's convolution as one tcgen05 implicit GEMM (inference)
  adam_step                      custom_optimizers.py:47-59 + Keras Adam + bts_train.py:125-131, one pass over flat buffers
  eval_metrics_png16             custom_eval_metrics.py:24-88 + bts_predict.py:140-141 (metrics and the uint16 depth image)
"""
import ctypes

import torch

from . import _cabi
from ._cabi import as_ref, check, current_stream_ptr, load, ptr_or_null


def _out_shape(coef, upratio):
    B, h, w, _ = coef.shape
    return B, h * upratio, w * upratio


def lpg_forward(coef, upratio, ds_stride=0, out_full=None, out_ds=None):
    """depth = LPG_r(coef) and, if ds_stride, depth[:, ::ds_stride, ::ds_stride].

    out_full / out_ds may be preallocated (e.g. a channel plane of a concat buffer); strides are
    honoured.  Returns (out_full, out_ds-or-None)."""
    lib = load()
    B, H, W = _out_shape(coef, upratio)
    if out_full is None:
        out_full = torch.empty((B, H, W, 1), dtype=coef.dtype, device=coef.device)
    if ds_stride and out_ds is None:
        out_ds = torch.empty((B, H // ds_stride, W // ds_stride, 1), dtype=coef.dtype, device=coef.device)
    rc, rf, rd = as_ref(coef), as_ref(out_full), as_ref(out_ds if ds_stride else None)
    check(lib.btslpg_forward(rc.ptr, int(upratio), rf.ptr, ptr_or_null(rd), int(ds_stride), current_stream_ptr(coef.device)))
    return out_full, (out_ds if ds_stride else None)


def _unit_stride_map(g):
    """Gradients that arrive as channel slices of an NHWC concat gradient (bts_decoder.py:42,99) have a
    column stride of the concat's channel count; one packing copy puts them on the vectorised kernels."""
    if g is not None and g.dim() >= 3 and g.shape[2] > 1 and g.stride(2) != 1:
        return g.contiguous()
    return g


def lpg_backward(coef, g_full, g_ds, upratio, ds_stride=0, g_coef=None):
    """d loss / d coef from d loss / d depth (g_full) and d loss / d depth_ds (g_ds); either may be None."""
    lib = load()
    if g_coef is None:
        g_coef = torch.empty_like(coef, memory_format=torch.contiguous_format)
    rc, rf, rd, rg = as_ref(coef), as_ref(g_full), as_ref(g_ds), as_ref(g_coef)
    check(lib.btslpg_backward(rc.ptr, ptr_or_null(rf), ptr_or_null(rd), int(upratio), int(ds_stride if g_ds is not None else 0),
                              rg.ptr, current_stream_ptr(coef.device)))
    return g_coef


def lpg_forward_multi(layers):
    """layers: list of dicts(coef, upratio, ds_stride, out_full, out_ds) -> one launch when all
    layers qualify for the vectorised kernels (else one launch per layer)."""
    lib = load()
    n = len(layers)
    args = (_cabi.BtsLpgForwardArgs * n)()
    keep = []
    for k, L in enumerate(layers):
        rc, rf = as_ref(L["coef"]), as_ref(L["out_full"])
        rd = as_ref(L.get("out_ds")) if L.get("ds_stride") else None
        keep += [rc, rf, rd]
        args[k].coef, args[k].upratio, args[k].ds_stride = rc.ptr, int(L["upratio"]), int(L.get("ds_stride") or 0)
        args[k].out_full = rf.ptr
        args[k].out_ds = rd.ptr if rd is not None else None
    check(lib.btslpg_forward_multi(args, n, current_stream_ptr(layers[0]["coef"].device)))


def lpg_backward_multi(layers):
    """layers: list of dicts(coef, g_full, g_ds, upratio, ds_stride, g_coef)."""
    lib = load()
    n = len(layers)
    args = (_cabi.BtsLpgBackwardArgs * n)()
    keep = []
    for k, L in enumerate(layers):
        rc, rg = as_ref(L["coef"]), as_ref(L["g_coef"])
        rf, rd = as_ref(L.get("g_full")), as_ref(L.get("g_ds"))
        keep += [rc, rg, rf, rd]
        args[k].coef, args[k].g_coef = rc.ptr, rg.ptr
        args[k].g_full = rf.ptr if rf is not None else None
        args[k].g_ds = rd.ptr if rd is not None else None
        args[k].upratio, args[k].ds_stride = int(L["upratio"]), int(L.get("ds_stride") or 0) if rd is not None else 0
    check(lib.btslpg_backward_multi(args, n, current_stream_ptr(layers[0]["coef"].device)))


class LpgFunction(torch.autograd.Function):
    """Differentiable LPG layer: (coef) -> (depth, depth_ds)."""

    @staticmethod
    def forward(ctx, coef, upratio, ds_stride):
        coef_c = coef.contiguous()
        full, ds = lpg_forward(coef_c, upratio, ds_stride)
        ctx.save_for_backward(coef_c)
        ctx.upratio, ctx.ds_stride = upratio, ds_stride
        ctx.set_materialize_grads(False)
        if ds is None:
            ds = full.new_empty(0)
            ctx.mark_non_differentiable(ds)
        return full, ds

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_full, g_ds):
        (coef,) = ctx.saved_tensors
        if not ctx.ds_stride:
            g_ds = None
        if g_full is None and g_ds is None:
            return torch.zeros_like(coef), None, None
        return lpg_backward(coef, _unit_stride_map(g_full), _unit_stride_map(g_ds), ctx.upratio, ctx.ds_stride), None, None


def local_planar_guidance(coef, upratio, ds_stride=0):
    """Functional form with autograd.  Returns depth, or (depth, depth_ds) when ds_stride > 0."""
    full, ds = LpgFunction.apply(coef, int(upratio), int(ds_stride))
    return (full, ds) if ds_stride else full


# ---------------------------------------------------------------------------------------------
# fused reduction head + LPG
# ---------------------------------------------------------------------------------------------
_workspaces = {}


def _workspace(device, nbytes):
    """Per (device, stream) scratch for the deterministic g_kernel reduction; its 256-byte header
    is zeroed once here and left zero by every launch."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _kernel2d(kernel):
    """Keras Conv2D kernel (1,1,C,3) HWIO or [C][3] -> contiguous float32 [C][3] view."""
    k = kernel.reshape(kernel.shape[-2], 3) if kernel.dim() == 4 else kernel
    if k.dtype != torch.float32 or not k.is_contiguous():
        k = k.float().contiguous()
    return k


def reduce_lpg_forward(feat, kernel, upratio, ds_stride=0, out_full=None, out_ds=None, coef_out=None):
    """coef = sigmoid(feat @ kernel); depth = LPG_r(coef); depth_ds = depth[:, ::d, ::d] in one kernel.
    Returns (coef, depth, depth_ds-or-None)."""
    lib = load()
    B, h, w, _ = feat.shape
    H, W = h * upratio, w * upratio
    if coef_out is None:
        coef_out = torch.empty((B, h, w, 3), dtype=feat.dtype, device=feat.device)
    if out_full is None:
        out_full = torch.empty((B, H, W, 1), dtype=feat.dtype, device=feat.device)
    if ds_stride and out_ds is None:
        out_ds = torch.empty((B, H // ds_stride, W // ds_stride, 1), dtype=feat.dtype, device=feat.device)
    k2 = _kernel2d(kernel)
    rf, rk, rc, ro = as_ref(feat), as_ref(k2), as_ref(coef_out), as_ref(out_full)
    rd = as_ref(out_ds) if ds_stride else None
    check(lib.btslpg_reduce_forward(rf.ptr, rk.ptr, int(upratio), rc.ptr, ro.ptr, ptr_or_null(rd), int(ds_stride),
                                    current_stream_ptr(feat.device)))
    return coef_out, out_full, (out_ds if ds_stride else None)


def reduce_lpg_backward(feat, kernel, coef, g_full, g_ds, upratio, ds_stride=0, need_g_feat=True, need_g_kernel=True,
                        g_kernel_out=None, need_g_coef=False):
    """Gradients of the fused op.  g_kernel_out may be a [C][3] float32 view into a flat gradient
    bucket (data-parallel training hands that bucket to the NCCL all-reduce).
    Returns (g_feat-or-None, g_kernel-or-None, g_coef-or-None)."""
    lib = load()
    C = feat.shape[-1]
    k2 = _kernel2d(kernel)
    g_feat = torch.empty_like(feat, memory_format=torch.contiguous_format) if need_g_feat else None
    g_kernel = None
    if need_g_kernel:
        g_kernel = g_kernel_out if g_kernel_out is not None else torch.empty((C, 3), dtype=torch.float32, device=feat.device)
    # channel counts / layouts outside the fused variants run the generic kernels, which need the LPG
    # coefficient gradient as scratch (the library never allocates)
    fused_ok = C in (32, 64, 128) and feat.is_contiguous() and coef.is_contiguous()
    g_coef = torch.empty_like(coef, memory_format=torch.contiguous_format) if (need_g_coef or not fused_ok) else None
    npix = feat.numel() // C
    nbytes = lib.btslpg_reduce_backward_workspace_bytes(npix, C)
    ws = _workspace(feat.device, nbytes)
    refs = [as_ref(feat), as_ref(k2), as_ref(coef), as_ref(g_full), as_ref(g_ds if ds_stride else None),
            as_ref(g_feat), as_ref(g_kernel), as_ref(g_coef)]
    check(lib.btslpg_reduce_backward(refs[0].ptr, refs[1].ptr, refs[2].ptr, ptr_or_null(refs[3]), ptr_or_null(refs[4]),
                                     int(upratio), int(ds_stride if refs[4] is not None else 0),
                                     ptr_or_null(refs[5]), ptr_or_null(refs[6]), ptr_or_null(refs[7]),
                                     ctypes.c_void_p(ws.data_ptr()), ws.numel(), current_stream_ptr(feat.device)))
    return g_feat, g_kernel, g_coef


class ReduceLpgFunction(torch.autograd.Function):
    """Differentiable fused head: (feat, kernel) -> (reduction, depth, depth_ds)."""

    @staticmethod
    def forward(ctx, feat, kernel, upratio, ds_stride, g_kernel_out=None, on_written=None):
        feat_c = feat.contiguous()
        coef, full, ds = reduce_lpg_forward(feat_c, kernel, upratio, ds_stride)
        ctx.save_for_backward(feat_c, kernel, coef)
        ctx.upratio, ctx.ds_stride = upratio, ds_stride
        # optional [C][3] float32 view into a flat gradient bucket: backward writes g_kernel there
        # directly (the bucket is what the all-reduce sends) instead of returning it to autograd
        ctx.g_kernel_out = g_kernel_out
        ctx.on_written = on_written
        ctx.set_materialize_grads(False)
        if ds is None:
            ds = full.new_empty(0)
            ctx.mark_non_differentiable(ds)
        # `reduction` is returned for inspection (layer name reduction_NxN); gradients flowing into it
        # directly are not part of the reference graph and are not supported.
        ctx.mark_non_differentiable(coef)
        return coef, full, ds

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _g_coef, g_full, g_ds):
        feat, kernel, coef = ctx.saved_tensors
        if not ctx.ds_stride:
            g_ds = None
        need_f, need_k = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if (g_full is None and g_ds is None) or not (need_f or need_k):
            return (torch.zeros_like(feat) if need_f else None), (torch.zeros_like(kernel) if need_k else None), None, None, None, None
        direct = ctx.g_kernel_out if need_k else None
        g_full, g_ds = _unit_stride_map(g_full), _unit_stride_map(g_ds)
        g_feat, g_kernel, _ = reduce_lpg_backward(feat, kernel, coef, g_full, g_ds, ctx.upratio, ctx.ds_stride,
                                                  need_g_feat=need_f, need_g_kernel=need_k,
                                                  g_kernel_out=None if direct is None else direct.view(-1, 3))
        if direct is not None:
            g_kernel = None                      # already in the bucket; nothing for autograd to accumulate
            if ctx.on_written is not None:
                ctx.on_written()
        elif g_kernel is not None:
            g_kernel = g_kernel.reshape(kernel.shape).to(kernel.dtype)
        return g_feat, g_kernel, None, None, None, None


def reduce_lpg(feat, kernel, upratio, ds_stride=0, g_kernel_out=None, on_written=None):
    """Functional fused head with autograd.  Returns (reduction, depth[, depth_ds]).
    g_kernel_out: optional float32 view (same numel as kernel) that receives d loss / d kernel in backward (overwritten,
    not accumulated); on_written: called once that backward kernel is enqueued."""
    coef, full, ds = ReduceLpgFunction.apply(feat, kernel, int(upratio), int(ds_stride), g_kernel_out, on_written)
    return (coef, full, ds) if ds_stride else (coef, full)


# ---------------------------------------------------------------------------------------------
# decoder tail: sigmoid * max_depth + si_log_loss, eval metrics
# ---------------------------------------------------------------------------------------------
def tail_workspace(device):
    """A fresh workspace for one silog forward/backward pair or one metrics call (header zeroed).  The
    forward leaves its statistics in it for the backward, so it is NOT shared between calls."""
    return torch.zeros(int(load().btslpg_tail_workspace_bytes()), dtype=torch.uint8, device=device)


def silog_forward(logit, y_true, max_depth, gt_threshold, depth_est=None, workspace=None):
    """depth_est = sigmoid(logit)*max_depth (logit given) and loss = si_log_loss(y_true, depth_est) (y_true given),
    one kernel.  Returns (depth_est, loss-or-None, workspace-or-None)."""
    lib = load()
    src = logit if logit is not None else depth_est
    if depth_est is None:
        depth_est = torch.empty_like(logit, memory_format=torch.contiguous_format)
    loss = None
    if y_true is not None:
        loss = torch.empty((), dtype=torch.float32, device=src.device)
        if workspace is None:
            workspace = tail_workspace(src.device)
    rz, rt, ry, rl = as_ref(logit), as_ref(y_true), as_ref(depth_est), as_ref(loss)
    check(lib.btslpg_silog_forward(ptr_or_null(rz), ptr_or_null(rt), float(max_depth), float(gt_threshold), ry.ptr, ptr_or_null(rl),
                                   ctypes.c_void_p(workspace.data_ptr() if workspace is not None else 0),
                                   workspace.numel() if workspace is not None else 0, current_stream_ptr(src.device)))
    return depth_est, loss, workspace


def silog_backward(depth_est, y_true, max_depth, gt_threshold, workspace, g_loss=None, wrt_logit=True, g_out=None):
    """d loss / d logit (wrt_logit) or d loss / d depth_est, scaled by the device scalar g_loss (None = 1)."""
    lib = load()
    if g_out is None:
        g_out = torch.empty_like(depth_est, memory_format=torch.contiguous_format)
    if g_loss is not None:
        g_loss = g_loss.reshape(1).float()
    ry, rt, rg, ro = as_ref(depth_est), as_ref(y_true), as_ref(g_loss), as_ref(g_out)
    check(lib.btslpg_silog_backward(ry.ptr, rt.ptr, float(max_depth), float(gt_threshold), ptr_or_null(rg),
                                    ctypes.c_void_p(workspace.data_ptr()), workspace.numel(), 1 if wrt_logit else 0, ro.ptr,
                                    current_stream_ptr(depth_est.device)))
    return g_out


class DepthSilogFunction(torch.autograd.Function):
    """(logit, y_true) -> (depth_est, loss): the last activation, the depth_est Lambda and the loss fused.
    depth_est is returned for inspection / metrics; only `loss` carries gradient."""

    @staticmethod
    def forward(ctx, logit, y_true, max_depth, gt_threshold, workspace=None):
        depth_est, loss, ws = silog_forward(logit.contiguous(), y_true.contiguous(), max_depth, gt_threshold, workspace=workspace)
        ctx.save_for_backward(depth_est, y_true.contiguous(), ws)
        ctx.max_depth, ctx.gt_threshold = max_depth, gt_threshold
        ctx.mark_non_differentiable(depth_est)
        return depth_est, loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _g_depth, g_loss):
        depth_est, y_true, ws = ctx.saved_tensors
        return silog_backward(depth_est, y_true, ctx.max_depth, ctx.gt_threshold, ws, g_loss, wrt_logit=True), None, None, None, None


class SilogLossFunction(torch.autograd.Function):
    """(y_true, y_pred) -> loss: bts.py:27-41 at the reference's own function boundary."""

    @staticmethod
    def forward(ctx, y_true, y_pred, gt_threshold):
        yt, yp = y_true.contiguous(), y_pred.contiguous()
        _, loss, ws = silog_forward(None, yt, 1.0, gt_threshold, depth_est=yp)
        ctx.save_for_backward(yp, yt, ws)
        ctx.gt_threshold = gt_threshold
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_loss):
        yp, yt, ws = ctx.saved_tensors
        return None, silog_backward(yp, yt, 1.0, ctx.gt_threshold, ws, g_loss, wrt_logit=False), None


def depth_silog(logit, y_true, max_depth, gt_threshold, workspace=None):
    """Fused decoder tail with autograd: returns (depth_est, loss).  `workspace` (tail_workspace(device)): a caller-owned
    scratch reused from step to step (one forward/backward pair in flight at a time) instead of a fresh zeroed one per call."""
    return DepthSilogFunction.apply(logit, y_true, float(max_depth), float(gt_threshold), workspace)


def si_log_loss(y_true, y_pred, gt_threshold):
    """si_log_loss(y_true, y_pred) of bts.py:31-38 with autograd (gradient with respect to y_pred)."""
    return SilogLossFunction.apply(y_true, y_pred, float(gt_threshold))


METRIC_NAMES = ("silog", "abs_rel", "log10", "rmse", "sq_rel", "rmse_log", "d1", "d2", "d3")   # custom_eval_metrics.py:88


def eval_metrics(y_true, y_pred, min_depth_eval, max_depth_eval, out=None, workspace=None):
    """All nine eval metrics of custom_eval_metrics.py in one pass.  Returns a float32 device tensor
    [silog, abs_rel, log10, rmse, sq_rel, rmse_log, d1, d2, d3, n_valid]."""
    lib = load()
    if out is None:
        out = torch.empty(10, dtype=torch.float32, device=y_true.device)
    if workspace is None:
        workspace = tail_workspace(y_true.device)
    rt, rp, ro = as_ref(y_true.contiguous()), as_ref(y_pred.contiguous()), as_ref(out)
    check(lib.btslpg_eval_metrics(rt.ptr, rp.ptr, float(min_depth_eval), float(max_depth_eval), ro.ptr,
                                  ctypes.c_void_p(workspace.data_ptr()), workspace.numel(), current_stream_ptr(y_true.device)))
    return out


def eval_metrics_png16(y_pred, png_max_depth, y_true=None, min_depth_eval=1e-3, max_depth_eval=80.0, out=None, png=None, workspace=None):
    """The metrics pass that also writes the uint16 depth image of bts_predict.py:140-141
    (`(pred * 65536 / max_depth).astype(np.uint16)`) from the same read of y_pred.  y_true None (bts_predict.py has no ground
    truth): the scaling pass alone.  Returns (png uint16 tensor shaped like y_pred, metrics-or-None)."""
    lib = load()
    y_pred = y_pred.contiguous()
    if png is None:
        png = torch.empty(y_pred.shape, dtype=torch.uint16, device=y_pred.device)
    rt = rm = None
    if y_true is not None:
        if out is None:
            out = torch.empty(10, dtype=torch.float32, device=y_pred.device)
        if workspace is None:
            workspace = tail_workspace(y_pred.device)
        rt, rm = as_ref(y_true.contiguous()), as_ref(out)
    rp, rg = as_ref(y_pred), as_ref(png)
    check(lib.btslpg_eval_metrics_png16(ptr_or_null(rt), rp.ptr, float(min_depth_eval), float(max_depth_eval), ptr_or_null(rm),
                                        float(png_max_depth), rg.ptr, ctypes.c_void_p(workspace.data_ptr() if workspace is not None else 0),
                                        workspace.numel() if workspace is not None else 0, current_stream_ptr(y_pred.device)))
    return png, (out if y_true is not None else None)


# ---------------------------------------------------------------------------------------------
# fused optimizer step over flat buffers (custom_optimizers.py:47-59 + Keras Adam + bts_train.py:125-131)
# ---------------------------------------------------------------------------------------------
def adam_config(lr_start, lr_end=None, total_steps=0, power=0.9, beta1=0.9, beta2=0.999, epsilon=1e-3, l1=0.0, l2=0.0,
                grad_scale=1.0, zero_grad=True):
    """BtsAdamConfig with the reference's defaults: Keras Adam betas, --adam_eps 1e-3 (bts_train.py:86), polynomial decay to
    0.1 * lr_start when no end rate is given (bts_train.py:126)."""
    if lr_end is None:
        lr_end = lr_start * 0.1 if total_steps > 0 else lr_start
    return _cabi.BtsAdamConfig(float(lr_start), float(lr_end), int(total_steps), float(power), float(beta1), float(beta2), float(epsilon),
                               float(l1), float(l2), float(grad_scale), 1 if zero_grad else 0)


def adam_state(device):
    """Device-resident optimizer state words: [0] int32 completed updates, [1] float32 last learning rate."""
    return torch.zeros(4, dtype=torch.float32, device=device)


def adam_step(param, grad, m, v, state, cfg, advance=True):
    """One fused AdamW update of the flat float32 buffers (param, m, v updated in place; grad consumed, zeroed when
    cfg.zero_grad).  `state` from adam_state(); the step counter is read and advanced on the device."""
    lib = load()
    rp, rg, rm, rv, rs = as_ref(param), as_ref(grad), as_ref(m), as_ref(v), as_ref(state)
    check(lib.btslpg_adam_step(rp.ptr, rg.ptr, rm.ptr, rv.ptr, rs.ptr, ctypes.byref(cfg), 1 if advance else 0, current_stream_ptr(param.device)))


# ---------------------------------------------------------------------------------------------
# fused activation + NHWC concat (concat1, conv_block concat)
# ---------------------------------------------------------------------------------------------
def _tensor_ptr_array(refs):
    arr = (_cabi._TP * max(len(refs), 1))()
    for k, r in enumerate(refs):
        arr[k] = r.ptr if r is not None else None
    return arr


def concat_forward(a, planes=(), b=None, act=False, out=None, pad=0, scale=None, shift=None, a_subpixel=False):
    """out[..., :CA] = affine(elu(a) if act else a) ; out[..., CA:CA+CB] = b ; one channel per plane ; `pad` zero
    channels.  affine (scale, shift: float32 [CA], optional) is an inference-mode BatchNormalization folded in.
    a_subpixel: `a` is (B,H/2,W/2,4*CA), the upconv evaluated on the low-res input (see subpixel_kernel).  One kernel."""
    lib = load()
    if a_subpixel:
        B, H, W, ca = a.shape[0], 2 * a.shape[1], 2 * a.shape[2], a.shape[3] // 4
    else:
        B, H, W, ca = a.shape
    cb = b.shape[-1] if b is not None else 0
    if out is None:
        out = torch.empty((B, H, W, ca + cb + len(planes) + pad), dtype=a.dtype, device=a.device)
    ra, rb, ro, rs, rt = as_ref(a), as_ref(b), as_ref(out), as_ref(scale), as_ref(shift)
    rp = [as_ref(p) for p in planes]
    check(lib.btslpg_concat_forward(ra.ptr, 1 if a_subpixel else 0, 1 if act else 0, ptr_or_null(rs), ptr_or_null(rt), ptr_or_null(rb),
                                    _tensor_ptr_array(rp), len(rp), int(pad), ro.ptr, current_stream_ptr(a.device)))
    return out


def concat_backward(g_out, y, act, ca, cb, n_planes, need_b=True, need_planes=None, pad=0, a_subpixel=False):
    """Split d concat into (g_a [* elu'(y)], g_b, [g_plane ...]); entries not needed come back as None.
    a_subpixel: g_a comes back as (B,H/2,W/2,4*CA), the layout of the low-res upconv's output."""
    lib = load()
    B, H, W, _ = g_out.shape
    need_planes = [True] * n_planes if need_planes is None else list(need_planes)
    g_a = torch.empty((B, H // 2, W // 2, 4 * ca) if a_subpixel else (B, H, W, ca), dtype=g_out.dtype, device=g_out.device)
    g_b = torch.empty((B, H, W, cb), dtype=g_out.dtype, device=g_out.device) if (cb and need_b) else None
    g_p = [torch.empty((B, H, W, 1), dtype=g_out.dtype, device=g_out.device) if need_planes[k] else None for k in range(n_planes)]
    if cb and not need_b:
        raise ValueError("concat_backward: the kernel needs g_b's geometry; pass need_b=True when CB > 0")
    rg, ry, ra, rb = as_ref(g_out), as_ref(y if act else None), as_ref(g_a), as_ref(g_b)
    rp = [as_ref(p) for p in g_p]
    check(lib.btslpg_concat_backward(rg.ptr, ptr_or_null(ry), 1 if act else 0, ra.ptr, 1 if a_subpixel else 0, ptr_or_null(rb),
                                     _tensor_ptr_array(rp), n_planes, int(pad), current_stream_ptr(g_out.device)))
    return g_a, g_b, g_p


class ConcatFunction(torch.autograd.Function):
    """(a, b-or-None, planes...) -> NHWC concat with the activation of `a` fused and `pad` zero channels appended."""

    @staticmethod
    def forward(ctx, a, b, act, pad, a_subpixel, *planes):
        a_c = a.contiguous()
        b_c = b.contiguous() if b is not None else None
        out = concat_forward(a_c, [p.contiguous() for p in planes], b_c, act, pad=pad, a_subpixel=a_subpixel)
        ctx.act, ctx.pad, ctx.sub = act, pad, a_subpixel
        ctx.ca, ctx.cb, ctx.np = a_c.shape[-1] // (4 if a_subpixel else 1), (b_c.shape[-1] if b_c is not None else 0), len(planes)
        if act:
            ctx.save_for_backward(out)          # elu' is taken from the output: nothing else is kept alive
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        y = ctx.saved_tensors[0] if ctx.act else None
        need_planes = [ctx.needs_input_grad[5 + k] for k in range(ctx.np)]
        g_a, g_b, g_p = concat_backward(g_out.contiguous(), y, ctx.act, ctx.ca, ctx.cb, ctx.np, need_b=True, need_planes=need_planes, pad=ctx.pad,
                                        a_subpixel=ctx.sub)
        return (g_a, g_b, None, None, None) + tuple(g_p)


def concat_nhwc(a, planes=(), b=None, act=False, pad=0, a_subpixel=False):
    """Fused `Concatenate(axis=3)([act(a), b, *planes])` (+ `pad` zero channels) with autograd (bts_decoder.py:98-99, :42)."""
    return ConcatFunction.apply(a, b, bool(act), int(pad), bool(a_subpixel), *planes)


# ---------------------------------------------------------------------------------------------
# training-mode conv block glue: ELU + BatchNormalization (batch statistics) + concat (bts_decoder.py:30-44)
# ---------------------------------------------------------------------------------------------
def bn_glue_supported(ca, ct, dtype):
    """The fused training glue takes float32, a power-of-two channel count in [4, 1024] and a concat width that is a multiple of 4."""
    return dtype == torch.float32 and 4 <= ca <= 1024 and (ca & (ca - 1)) == 0 and ct % 4 == 0


def bn_elu_stats(raw, gamma, beta, running_mean, running_var, momentum, eps, pack=None, act=True):
    """Batch statistics of elu(raw) per channel -> pack [8][C] (scale, shift, mean, std, 1/gamma, beta, -, -); updates the moving
    averages in place when given.  One read of `raw`, deterministic."""
    lib = load()
    C = raw.shape[-1]
    if pack is None:
        pack = torch.empty((8, C), dtype=torch.float32, device=raw.device)
    ws = _workspace(raw.device, int(lib.btslpg_bn_workspace_bytes(C)))
    rr, rg, rb, rp = as_ref(raw), as_ref(gamma), as_ref(beta), as_ref(pack)
    rm, rv = as_ref(running_mean), as_ref(running_var)
    check(lib.btslpg_bn_elu_stats(rr.ptr, 1 if act else 0, rg.ptr, rb.ptr, ptr_or_null(rm), ptr_or_null(rv), float(momentum), float(eps),
                                  rp.ptr, ctypes.c_void_p(ws.data_ptr()), ws.numel(), current_stream_ptr(raw.device)))
    return pack


class ConvBlockGlueFunction(torch.autograd.Function):
    """(raw upconv output, skip, gamma, beta, planes...) -> concat [BN(elu(raw)), skip, planes..., zero pad] in training mode, as one
    statistics pass + one fused concat pass forward and the same backward (csrc/bnstat_kernels.cuh, csrc/concat_kernels.cuh)."""

    @staticmethod
    def forward(ctx, raw, skip, gamma, beta, running_mean, running_var, momentum, eps, pad, *planes):
        raw_c, skip_c = raw.contiguous(), skip.contiguous()
        pack = bn_elu_stats(raw_c, gamma.detach().float().contiguous(), beta.detach().float().contiguous(), running_mean, running_var, momentum, eps)
        out = concat_forward(raw_c, [p.contiguous() for p in planes], skip_c, act=True, pad=pad, scale=pack[0], shift=pack[1])
        ctx.save_for_backward(out, pack)
        ctx.ca, ctx.cb, ctx.np, ctx.pad = raw_c.shape[-1], skip_c.shape[-1], len(planes), pad
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        out, pack = ctx.saved_tensors
        lib = load()
        g_out = g_out.contiguous()
        B, H, W, _ = g_out.shape
        g_gamma = torch.empty(ctx.ca, dtype=torch.float32, device=g_out.device)
        g_beta = torch.empty(ctx.ca, dtype=torch.float32, device=g_out.device)
        ws = _workspace(g_out.device, int(lib.btslpg_bn_workspace_bytes(ctx.ca)))
        rg, ry, rp, rgg, rgb = as_ref(g_out), as_ref(out), as_ref(pack), as_ref(g_gamma), as_ref(g_beta)
        check(lib.btslpg_bn_elu_backward_stats(rg.ptr, ry.ptr, int(ctx.ca), rp.ptr, rgg.ptr, rgb.ptr, ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                                               current_stream_ptr(g_out.device)))
        g_a = torch.empty((B, H, W, ctx.ca), dtype=g_out.dtype, device=g_out.device)
        g_b = torch.empty((B, H, W, ctx.cb), dtype=g_out.dtype, device=g_out.device)
        need_planes = [ctx.needs_input_grad[9 + k] for k in range(ctx.np)]
        g_p = [torch.empty((B, H, W, 1), dtype=g_out.dtype, device=g_out.device) if need_planes[k] else None for k in range(ctx.np)]
        ra, rb = as_ref(g_a), as_ref(g_b)
        rpl = [as_ref(p) for p in g_p]
        check(lib.btslpg_concat_backward_bn(rg.ptr, ry.ptr, 1, rp.ptr, ra.ptr, 0, rb.ptr, _tensor_ptr_array(rpl), ctx.np, int(ctx.pad),
                                            current_stream_ptr(g_out.device)))
        return (g_a, g_b, g_gamma, g_beta, None, None, None, None, None) + tuple(g_p)


def conv_block_glue(raw, skip, planes, bn, pad=0):
    """bts_decoder.py:32-42 in TRAINING mode for a torch BatchNorm2d `bn` (its weight / bias / running statistics are used and
    updated): concat [bn(elu(raw)), skip, *planes] (+ `pad` zero channels), NHWC."""
    return ConvBlockGlueFunction.apply(raw, skip, bn.weight, bn.bias, bn.running_mean, bn.running_var, float(bn.momentum), float(bn.eps),
                                       int(pad), *planes)


_subpixel_R = {}


def subpixel_kernel(weight):
    """Kernel of the 3x3 upconv evaluated on the LOW-RES input: OIHW (Cout,Cin,3,3) -> (4*Cout,Cin,3,3), output channels
    ordered (row parity, column parity, o).  UpSampling2D(2,'nearest') + Conv2D(3x3,'same') (bts_decoder.py:97-98) equals this
    convolution followed by a pixel shuffle: an even output row sees low-res rows (y-1: w0), (y: w1+w2), an odd one
    (y: w0+w1), (y+1: w2); same for columns.  Differentiable (a fixed linear map of the weights)."""
    key = (weight.device, weight.dtype)
    R = _subpixel_R.get(key)
    if R is None:       # built once per device (a host-to-device copy: not allowed while a CUDA graph is being captured)
        R = torch.tensor([[[1, 0, 0], [0, 1, 1], [0, 0, 0]], [[0, 0, 0], [1, 1, 0], [0, 0, 1]]], dtype=weight.dtype).to(weight.device)
        _subpixel_R[key] = R
    cout, cin = weight.shape[:2]
    return torch.einsum("aik,bjl,ockl->abocij", R, R, weight).reshape(4 * cout, cin, 3, 3)


def pad_to(channels, multiple=4):
    """Zero channels to append so that a concat is a multiple of `multiple` channels wide."""
    return (-channels) % multiple


# ---------------------------------------------------------------------------------------------
# nearest x2 up-sampling (UpSampling2D in front of every upconv)
# ---------------------------------------------------------------------------------------------
def upsample2x_forward(x, out=None):
    """x (B,h,w,C) NHWC contiguous -> (B,2h,2w,C), out[b,y,x] = in[b,y//2,x//2]."""
    lib = load()
    B, h, w, C = x.shape
    if out is None:
        out = torch.empty((B, 2 * h, 2 * w, C), dtype=x.dtype, device=x.device)
    ri, ro = as_ref(x), as_ref(out)
    check(lib.btslpg_upsample2x_forward(ri.ptr, ro.ptr, current_stream_ptr(x.device)))
    return out


def upsample2x_backward(g_out, g_in=None):
    lib = load()
    B, H, W, C = g_out.shape
    if g_in is None:
        g_in = torch.empty((B, H // 2, W // 2, C), dtype=g_out.dtype, device=g_out.device)
    rg, ri = as_ref(g_out), as_ref(g_in)
    check(lib.btslpg_upsample2x_backward(rg.ptr, ri.ptr, current_stream_ptr(g_out.device)))
    return g_in


class Upsample2xFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return upsample2x_forward(x.contiguous())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        return upsample2x_backward(g_out.contiguous())


def upsample2x_nhwc(x):
    """`UpSampling2D(size=2, interpolation='nearest')` on an NHWC tensor, with autograd."""
    return Upsample2xFunction.apply(x)


# ---------------------------------------------------------------------------------------------
# strided per-channel affine + activation copy (DenseASPP glue, inference)
# ---------------------------------------------------------------------------------------------
ACT_NONE, ACT_ELU, ACT_RELU = 0, 1, 2


def affine_act(src, dst=None, scale=None, shift=None, act=ACT_NONE, src_split=0, dst_split=0):
    """dst[..., c] = act(src[..., c] * scale[c] + shift[c]); src / dst are (B,H,W,C) views with channel stride 1 and
    uniformly strided pixels (channel slices of wider NHWC buffers), may alias.  No autograd (inference glue).
    src_split / dst_split = s > 1: that side is in sub-grid form, (B*s*s, H/s, W/s, C) with image (b*s + i)*s + j = [i::s, j::s] of
    image b (what the split dilated convolutions of the DenseASPP read and write; decoder._dilation_split)."""
    lib = load()
    s = int(src_split or dst_split or 1)
    if src_split and dst_split:
        raise ValueError("affine_act: only one side can be in sub-grid form")
    if dst is None:
        B, H, W, C = src.shape
        shape = (B * s * s, H // s, W // s, C) if dst_split else (B // (s * s), H * s, W * s, C) if src_split else src.shape
        dst = torch.empty(shape, dtype=src.dtype, device=src.device)
    rs, rd, rsc, rsh = as_ref(src), as_ref(dst), as_ref(scale), as_ref(shift)
    check(lib.btslpg_affine_act_split(rs.ptr, ptr_or_null(rsc), ptr_or_null(rsh), int(act), rd.ptr, 1 if src_split else 2 if dst_split else 0, s,
                                      current_stream_ptr(src.device)))
    return dst


# ---------------------------------------------------------------------------------------------
# training-mode BatchNormalization (+ ReLU) over channel slices (DenseASPP glue, bts_decoder.py:46-54, :61-76); csrc/bnrelu_kernels.cuh
# ---------------------------------------------------------------------------------------------
def bn_slices_supported(nf, dtype):
    """DenseASPP widths the training glue takes: float32, nf a multiple of 8 (half-width pieces of whole 16-byte vectors), slices of
    at most 1024 channels (the widest is 2 * nf)."""
    return dtype == torch.float32 and nf % 8 == 0 and 8 <= nf <= 512


def _bn_ws(device, channels):
    lib = load()
    ws = _workspace(device, int(lib.btslpg_bn_workspace_bytes(int(channels))))
    return ctypes.c_void_p(ws.data_ptr()), ws.numel()


def bn_moments(x, mean, var):
    """Per-channel mean and biased variance of the (B,H,W,C) channel slice `x` into the float32 [C] vectors `mean`, `var`."""
    lib = load()
    wp, wn = _bn_ws(x.device, x.shape[-1])
    rx, rm, rv = as_ref(x), as_ref(mean), as_ref(var)
    check(lib.btslpg_bn_moments(rx.ptr, rm.ptr, rv.ptr, wp, wn, current_stream_ptr(x.device)))


def bn_fold(mean, var, gamma, beta, running_mean, running_var, momentum, eps, count):
    """Batch moments + gamma / beta -> (scale, shift, mean, rstd), the vectors the forward (affine_act) and the backward take; updates
    the moving averages in place when given."""
    lib = load()
    scale, shift, rstd = (torch.empty_like(mean) for _ in range(3))
    refs = [as_ref(t) for t in (mean, var, gamma, beta, running_mean, running_var, scale, shift, rstd)]
    check(lib.btslpg_bn_fold(refs[0].ptr, refs[1].ptr, refs[2].ptr, refs[3].ptr, ptr_or_null(refs[4]), ptr_or_null(refs[5]), float(momentum),
                             float(eps), int(count), refs[6].ptr, refs[7].ptr, refs[8].ptr, current_stream_ptr(mean.device)))
    return scale, shift, mean, rstd


def bn_act_backward(g, x, vecs, g_gamma, g_beta, dst, relu=True, accumulate=False, g2=None, dst_init=None):
    """Backward of y = act(x * scale + shift) with batch statistics, over one channel slice: fills g_gamma / g_beta [C] and writes or
    adds d loss / d x into `dst` (dst_init: dst = dst_init + d loss / d x).  vecs = (scale, shift, mean, rstd) slices; g2: gradient
    reaching the normalised value directly."""
    lib = load()
    wp, wn = _bn_ws(g.device, g.shape[-1])
    rg, rh, rx, rd, ri = as_ref(g), as_ref(g2), as_ref(x), as_ref(dst), as_ref(dst_init)
    rv = [as_ref(v) for v in vecs]
    rgg, rgb = as_ref(g_gamma), as_ref(g_beta)
    st = current_stream_ptr(g.device)
    check(lib.btslpg_bn_act_backward_stats(rg.ptr, ptr_or_null(rh), rx.ptr, rv[0].ptr, rv[1].ptr, rv[2].ptr, rv[3].ptr, 1 if relu else 0,
                                           rgg.ptr, rgb.ptr, wp, wn, st))
    check(lib.btslpg_bn_act_backward(rg.ptr, ptr_or_null(rh), rx.ptr, rv[0].ptr, rv[1].ptr, rv[2].ptr, rv[3].ptr, rgg.ptr, rgb.ptr,
                                     1 if relu else 0, rd.ptr, 1 if accumulate else 0, ptr_or_null(ri), st))
    return dst


# ---------------------------------------------------------------------------------------------
# last convolution Conv2D(1, 3x3, 'same'): hand-written forward (with the neighbouring ELU / sigmoid folded in) and backward
# ---------------------------------------------------------------------------------------------
def depthconv_backward(x, kernel9c, g_out, need_g_x=True, need_g_kernel=True, act_in=False):
    """x (B,H,W,C) NHWC, kernel9c float32 [9*C] ([tap][c] == Keras HWIO (3,3,C,1)), g_out (B,H,W,1).
    act_in: x is the pre-activation and the convolution saw elu(x) (see depthconv_forward); g_x is then d loss / d x.
    Returns (g_x-or-None, g_kernel [9*C] float32-or-None), both from one pass."""
    lib = load()
    C = x.shape[-1]
    g_x = torch.empty_like(x, memory_format=torch.contiguous_format) if need_g_x else None
    g_k = torch.empty(9 * C, dtype=torch.float32, device=x.device) if need_g_kernel else None
    ws = _workspace(x.device, int(lib.btslpg_depthconv_backward_workspace_bytes(C))) if need_g_kernel else None
    rx, rk, rg, rgx, rgk = as_ref(x), as_ref(kernel9c), as_ref(g_out), as_ref(g_x), as_ref(g_k)
    check(lib.btslpg_depthconv_backward(rx.ptr, rk.ptr, rg.ptr, int(bool(act_in)), ptr_or_null(rgx), ptr_or_null(rgk),
                                        ctypes.c_void_p(ws.data_ptr() if ws is not None else 0), ws.numel() if ws is not None else 0,
                                        current_stream_ptr(x.device)))
    return g_x, g_k


def kernel9c(weight):
    """torch OIHW (1,C,3,3) conv weight -> float32 [9*C] in Keras HWIO (3,3,C,1) memory order ([ky][kx][c])."""
    return weight.detach().permute(2, 3, 1, 0).reshape(-1).float().contiguous()


def depthconv_forward(x, kernel9c, act_in=False, sigmoid_scale=None, out=None):
    """bts_decoder.py:100-103 in one pass: y = conv3x3_same(elu(x) if act_in else x) (one output channel), then
    sigmoid(y) * sigmoid_scale when a scale is given.  x (B,H,W,C) contiguous NHWC, C in (16, 32); kernel9c float32
    [9*C] ([tap][c] == Keras HWIO (3,3,C,1)).  Returns (B,H,W,1).  No autograd (see depth_conv)."""
    lib = load()
    y = out if out is not None else torch.empty(x.shape[:3] + (1,), dtype=x.dtype, device=x.device)
    rx, rk, ry = as_ref(x), as_ref(kernel9c), as_ref(y)
    check(lib.btslpg_depthconv_forward(rx.ptr, rk.ptr, int(bool(act_in)), int(sigmoid_scale is not None),
                                       float(sigmoid_scale if sigmoid_scale is not None else 1.0), ry.ptr, current_stream_ptr(x.device)))
    return y


class DepthConvFunction(torch.autograd.Function):
    """x NHWC (B,H,W,C), weight torch OIHW (1,C,3,3) -> (B,H,W,1): hand-written forward (float32-accurate) and one
    hand-written pass for both gradients.  act_in: x is the PRE-activation of iconv1 and the ELU (bts_decoder.py:100) runs
    inside both kernels (no ELU tensor, no separate ELU forward / backward pass)."""

    @staticmethod
    def forward(ctx, x_nhwc, weight, act_in=False):
        x_nhwc = x_nhwc.contiguous()
        ctx.save_for_backward(x_nhwc, weight)
        ctx.act_in = bool(act_in)
        return depthconv_forward(x_nhwc, kernel9c(weight), act_in=act_in)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        x, weight = ctx.saved_tensors
        g_x, g_k = depthconv_backward(x, kernel9c(weight), g_out.contiguous(), ctx.needs_input_grad[0], ctx.needs_input_grad[1], act_in=ctx.act_in)
        g_w = None
        if g_k is not None:
            C = x.shape[-1]
            g_w = g_k.view(3, 3, C, 1).permute(3, 2, 0, 1).to(weight.dtype)                      # back to OIHW
        return g_x, g_w, None


def depth_conv(x_nhwc, weight, act_in=False):
    """The decoder's last convolution with autograd; C in (16, 32).  act_in=True: conv(elu(x)) with the ELU folded in."""
    return DepthConvFunction.apply(x_nhwc, weight, act_in)


# ---------------------------------------------------------------------------------------------
# iconv1 as a tcgen05 implicit GEMM over the concat's sources (inference)
# ---------------------------------------------------------------------------------------------
def conv3x3_wgrad(x, g, out=None):
    """d kernel (HWIO, float32 [3,3,Cin,Cout]) of a stride-1 'same' 3x3 convolution from its NHWC input `x` and the gradient `g` of
    its output, on the tcgen05 tensor cores (TF32 operands, float32 accumulation; csrc/wgrad_kernels.cuh)."""
    lib = load()
    cin, cout = x.shape[-1], g.shape[-1]
    if out is None:
        out = torch.empty((3, 3, cin, cout), dtype=torch.float32, device=x.device)
    ws = _workspace(x.device, int(lib.btslpg_conv3x3_wgrad_workspace_bytes(int(cin), int(cout))))
    rx, rg, ro = as_ref(x), as_ref(g), as_ref(out)
    check(lib.btslpg_conv3x3_wgrad(rx.ptr, rg.ptr, ro.ptr, ctypes.c_void_p(ws.data_ptr()), ws.numel(), current_stream_ptr(x.device)))
    return out


def kernel_hwio(weight):
    """torch OIHW (O,I,3,3) conv weight -> float32 Keras HWIO (3,3,I,O) contiguous."""
    return weight.detach().permute(2, 3, 1, 0).float().contiguous()


def iconv1_forward(a, planes, kernel_hwio_, a_subpixel=False, act_out=False, out=None):
    """bts_decoder.py:98-100 without concat1: out = conv3x3_same([elu(a), d2, d4, d8]) (then ELU if act_out).
    a: upconv1's linear output (B,H,W,NF), NF in (16, 32), or its sub-pixel form (B,H/2,W/2,4*NF); planes: the three LPG maps
    (B,H,W,1); kernel_hwio_: float32 (3,3,NF+3,NF).  TF32 tensor-core arithmetic (tcgen05), float32 accumulation.  No autograd."""
    lib = load()
    if a_subpixel:
        B, H, W, nf = a.shape[0], 2 * a.shape[1], 2 * a.shape[2], a.shape[3] // 4
    else:
        B, H, W, nf = a.shape
    if out is None:
        out = torch.empty((B, H, W, nf), dtype=a.dtype, device=a.device)
    ra, rk, ro = as_ref(a), as_ref(kernel_hwio_), as_ref(out)
    rp = [as_ref(p.contiguous()) for p in planes]
    check(lib.btslpg_iconv1_forward(ra.ptr, 1 if a_subpixel else 0, _tensor_ptr_array(rp), rk.ptr, 1 if act_out else 0, ro.ptr,
                                    current_stream_ptr(a.device)))
    return out


def launch_count():
    return int(load().btslpg_launch_count())


def reset_launch_count():
    load().btslpg_reset_launch_count()


def last_kernel():
    return load().btslpg_last_kernel().decode()


def set_block_threads(fwd=0, bwd=0):
    load().btslpg_set_block_threads(int(fwd), int(bwd))


def set_tuning(key, value):
    """Experiment knobs of the library (see include/btslpg.h: btslpg_set_tuning)."""
    load().btslpg_set_tuning(int(key), int(value))
