"""TensorFlow binding of the same C ABI -- the drop-in for the reference's own files.

TensorFlow is not installable in the build image (no wheel, no network), so this module is
import-guarded and untested here; it is the concrete form of the stub shown in INTEGRATION.md.
With TF present, `LocalPlanarGuidance` below IS a `tf.keras.layers.Layer` with the reference's
name, constructor, `build`, `call` and `get_config` (custom_layers.py:25-61), so
`from custom_layers import LocalPlanarGuidance` in bts_decoder.py:24 can be pointed at it
unchanged.  Tensors cross into libbtslpg.so as DLPack capsules (zero copy): TF allocates inputs and
outputs, the library only launches kernels on TF's buffers.

Status: EXPERIMENTAL -- written against the TF 2.x Python API, never executed (no TensorFlow here).

Stream note: TF runs its GPU ops on its own compute stream, created with CU_STREAM_NON_BLOCKING and not exposed to
Python, so a launch on the legacy default stream is NOT ordered against TF's work.  Every entry point below therefore
brackets its launch with a device-wide barrier (`btslpg_device_synchronize`): one before the launch, after the output
buffers have been allocated (TF's producers of the inputs and its `tf.zeros` fills are complete), and one after it (the
results are written before TF's consumers are enqueued).  That is correct but serialises host and device at every call,
and `tf.py_function` runs its body under the GIL; the production binding is a `tf.load_op_library` custom op whose
`Compute()` passes `ctx->eigen_gpu_device().stream()` as the ABI's stream argument -- it needs
TensorFlow's headers to build, which this image does not have.
"""
import ctypes

from . import _cabi

try:  # pragma: no cover - TensorFlow is absent in the build image
    import tensorflow as tf
    HAVE_TF = True
except Exception:  # ImportError or a broken install
    tf = None
    HAVE_TF = False


def _require_tf():
    if not HAVE_TF:
        raise ImportError("TensorFlow is not installed; use bts_fully_tf_b200.LocalPlanarGuidance (torch host) instead")


def _capsule(t):
    return tf.experimental.dlpack.to_dlpack(t)


def _sync(ref):  # pragma: no cover
    """Device-wide barrier on the device of a described tensor (see the stream note in the module docstring)."""
    _cabi.check(_cabi.load().btslpg_device_synchronize(int(ref.struct.device.device_id)))


def _gpu_py_function(fn, inputs, tout):  # pragma: no cover
    """tf.py_function pinned to the device of its first input, so that the body receives device tensors (the library
    rejects host tensors: there is no CPU path)."""
    dev = getattr(inputs[0], "device", "") or None
    if dev:
        with tf.device(dev):
            return tf.py_function(fn, inputs, tout)
    return tf.py_function(fn, inputs, tout)


def _forward_eager(coef, upratio, ds_stride):  # pragma: no cover
    lib = _cabi.load()
    B, h, w, _ = coef.shape
    H, W = h * upratio, w * upratio
    with tf.device(coef.device):
        full = tf.zeros((B, H, W, 1), coef.dtype)
        ds = tf.zeros((B, H // ds_stride, W // ds_stride, 1), coef.dtype) if ds_stride else None
    rc, rf = _cabi.from_dlpack_capsule(_capsule(coef)), _cabi.from_dlpack_capsule(_capsule(full))
    rd = _cabi.from_dlpack_capsule(_capsule(ds)) if ds_stride else None
    _sync(rc)
    _cabi.check(lib.btslpg_forward(rc.ptr, int(upratio), rf.ptr, _cabi.ptr_or_null(rd), int(ds_stride), ctypes.c_void_p(0)))
    _sync(rc)
    return full, ds


def _backward_eager(coef, g_full, g_ds, upratio, ds_stride):  # pragma: no cover
    lib = _cabi.load()
    with tf.device(coef.device):
        g_coef = tf.zeros(coef.shape, coef.dtype)
    refs = [_cabi.from_dlpack_capsule(_capsule(t)) if t is not None else None for t in (coef, g_full, g_ds, g_coef)]
    _sync(refs[0])
    _cabi.check(lib.btslpg_backward(refs[0].ptr, _cabi.ptr_or_null(refs[1]), _cabi.ptr_or_null(refs[2]), int(upratio),
                                    int(ds_stride if g_ds is not None else 0), refs[3].ptr, ctypes.c_void_p(0)))
    _sync(refs[0])
    return g_coef


def local_planar_guidance(inputs, upratio):  # pragma: no cover
    """Differentiable TF op: (B,h,w,3) -> (B,h*r,w*r,1); works eagerly and inside tf.function."""
    _require_tf()

    @tf.custom_gradient
    def op(x):
        def fwd(x_):
            return _forward_eager(x_, upratio, 0)[0]

        y = _gpu_py_function(fwd, [x], x.dtype)
        y.set_shape([x.shape[0], x.shape[1] * upratio, x.shape[2] * upratio, 1])

        def grad(dy):
            g = _gpu_py_function(lambda x_, dy_: _backward_eager(x_, dy_, None, upratio, 0), [x, dy], x.dtype)
            g.set_shape(x.shape)
            return g

        return y, grad

    return op(inputs)


if HAVE_TF:  # pragma: no cover
    class LocalPlanarGuidance(tf.keras.layers.Layer):
        """Drop-in for reference custom_layers.py:25-61 backed by libbtslpg.so."""

        def __init__(self, upratio, **kwargs):
            super().__init__(**kwargs)
            self.upratio = upratio

        def build(self, input_shape):
            assert len(input_shape) > 2
            # the (1,H,W,3) pixel_dir_unit constant of the reference is not materialised: its r*r distinct
            # vectors live in the kernels' __constant__ memory
            return super().build(input_shape)

        def call(self, inputs):
            return local_planar_guidance(inputs, self.upratio)

        def get_config(self):
            base_config = super().get_config()
            config = {"upratio": self.upratio}
            return dict(list(base_config.items()) + list(config.items()))


# ---------------------------------------------------------------------------------------------
# decoder tail and concat (INTEGRATION.md sections 3a, 3b): same C ABI, TF-allocated buffers
# ---------------------------------------------------------------------------------------------
GT_TH = {"nyu": 0.1, "kitti": 1.0, "matterport": 0.1}      # bts.py:28


def _ref(t):  # pragma: no cover
    return _cabi.from_dlpack_capsule(_capsule(t)) if t is not None else None


def _tail_workspace(device):  # pragma: no cover
    with tf.device(device):
        return tf.zeros([int(_cabi.load().btslpg_tail_workspace_bytes())], tf.uint8)


def si_log_loss_wrapper(dataset):  # pragma: no cover
    """Drop-in for reference bts.py:27-41: same name, argument and assertion; one kernel forward, one backward."""
    _require_tf()
    assert dataset in GT_TH

    @tf.custom_gradient
    def si_log_loss(y_true, y_pred):
        lib = _cabi.load()
        state = {}

        def fwd(yt, yp):
            ws = _tail_workspace(yp.device)
            with tf.device(yp.device):
                loss = tf.zeros([1], tf.float32)
            state["ws"] = ws
            rt, rp, rl, rw = _ref(yt), _ref(yp), _ref(loss), _ref(ws)
            _sync(rp)
            _cabi.check(lib.btslpg_silog_forward(None, rt.ptr, 1.0, GT_TH[dataset], rp.ptr, rl.ptr,
                                                 ctypes.c_void_p(rw.struct.data), int(ws.shape[0]), ctypes.c_void_p(0)))
            _sync(rp)
            return loss[0]

        loss = _gpu_py_function(fwd, [y_true, y_pred], tf.float32)
        loss.set_shape([])

        def grad(g_loss):
            def bwd(yt, yp, gl):
                with tf.device(yp.device):
                    g = tf.zeros(yp.shape, yp.dtype)
                    gl1 = tf.reshape(tf.cast(gl, tf.float32), [1])
                ws = state["ws"]
                rt, rp, rg, rgl, rw = _ref(yt), _ref(yp), _ref(g), _ref(gl1), _ref(ws)
                _sync(rp)
                _cabi.check(lib.btslpg_silog_backward(rp.ptr, rt.ptr, 1.0, GT_TH[dataset], rgl.ptr, ctypes.c_void_p(rw.struct.data),
                                                      int(ws.shape[0]), 0, rg.ptr, ctypes.c_void_p(0)))
                _sync(rp)
                return g
            g = _gpu_py_function(bwd, [y_true, y_pred, g_loss], y_pred.dtype)
            g.set_shape(y_pred.shape)
            return None, g

        return loss, grad

    return si_log_loss


def metrics_list_factory(args):  # pragma: no cover
    """Drop-in for reference custom_eval_metrics.py:21-88: nine named callables, ONE fused pass per (y_true, y_pred)."""
    _require_tf()
    names = ("silog", "abs_rel", "log10", "rmse", "sq_rel", "rmse_log", "d1", "d2", "d3")
    cache = {}        # holds the LAST (y_true, y_pred) pair itself: identity of live objects, never a recycled id()

    def all_metrics(y_true, y_pred):
        if not (cache.get("yt") is y_true and cache.get("yp") is y_pred):
            cache.clear()
            cache["yt"], cache["yp"] = y_true, y_pred

            def run(yt, yp):
                lib = _cabi.load()
                ws = _tail_workspace(yp.device)
                with tf.device(yp.device):
                    out = tf.zeros([10], tf.float32)
                rt, rp, ro, rw = _ref(yt), _ref(yp), _ref(out), _ref(ws)
                _sync(rt)
                _cabi.check(lib.btslpg_eval_metrics(rt.ptr, rp.ptr, float(args.min_depth_eval), float(args.max_depth_eval), ro.ptr,
                                                    ctypes.c_void_p(rw.struct.data), int(ws.shape[0]), ctypes.c_void_p(0)))
                _sync(rt)
                return out
            v = _gpu_py_function(run, [y_true, y_pred], tf.float32)
            v.set_shape([10])
            cache["v"] = v
        return cache["v"]

    def make(i, name):
        def metric(y_true, y_pred):
            return all_metrics(y_true, y_pred)[i]
        metric.__name__ = name
        return metric

    return [make(i, n) for i, n in enumerate(names)]


def concat1(upconv1_linear, d2, d4, d8):  # pragma: no cover
    """bts_decoder.py:98-99 fused: ELU of upconv1 (built with activation=None) + Concatenate([upconv1, d2, d4, d8])."""
    _require_tf()

    @tf.custom_gradient
    def op(a, p0, p1, p2):
        lib = _cabi.load()

        def fwd(a_, p0_, p1_, p2_):
            with tf.device(a_.device):
                out = tf.zeros(a_.shape[:3] + [a_.shape[3] + 3], a_.dtype)
            ra, ro = _ref(a_), _ref(out)
            rp = [_ref(t) for t in (p0_, p1_, p2_)]
            arr = (_cabi._TP * 3)(*[r.ptr for r in rp])
            _sync(ra)
            _cabi.check(lib.btslpg_concat_forward(ra.ptr, 0, 1, None, None, None, arr, 3, 0, ro.ptr, ctypes.c_void_p(0)))
            _sync(ra)
            return out

        y = _gpu_py_function(fwd, [a, p0, p1, p2], a.dtype)
        y.set_shape(a.shape[:3] + [a.shape[3] + 3])

        def grad(g_out):
            def bwd(g_, y_):
                with tf.device(g_.device):
                    g_a = tf.zeros(a.shape, a.dtype)
                    g_p = [tf.zeros(p0.shape, a.dtype) for _ in range(3)]
                rg, ry, rga = _ref(g_), _ref(y_), _ref(g_a)
                rp = [_ref(t) for t in g_p]
                arr = (_cabi._TP * 3)(*[r.ptr for r in rp])
                _sync(rg)
                _cabi.check(lib.btslpg_concat_backward(rg.ptr, ry.ptr, 1, rga.ptr, 0, None, arr, 3, 0, ctypes.c_void_p(0)))
                _sync(rg)
                return [g_a] + g_p
            outs = _gpu_py_function(bwd, [g_out, y], [a.dtype] * 4)
            outs[0].set_shape(a.shape)
            for o in outs[1:]:
                o.set_shape(p0.shape)
            return tuple(outs)

        return y, grad

    return op(upconv1_linear, d2, d4, d8)


def depth_tail(iconv1_linear, kernel, max_depth=None):  # pragma: no cover
    """bts_decoder.py:100-103 fused.  `iconv1_linear`: the output of iconv1's convolution built with activation=None;
    `kernel`: the (3,3,C,1) kernel variable of the last Conv2D (C = 16 or 32).  max_depth given (inference): returns
    depth_est = sigmoid(conv(elu(iconv1_linear))) * max_depth from ONE kernel.  max_depth None (training): returns the
    logit with a custom gradient -- ONE backward kernel for d loss / d iconv1_linear (ELU' included) and d loss / d kernel;
    feed it to si_log_loss_wrapper's fused form together with max_depth."""
    _require_tf()
    lib = _cabi.load()

    def fwd(x_, k_):
        with tf.device(x_.device):
            y_ = tf.zeros(x_.shape[:3] + [1], x_.dtype)
        rx, rk, ry = _ref(x_), _ref(tf.reshape(k_, [-1])), _ref(y_)
        _sync(rx)
        _cabi.check(lib.btslpg_depthconv_forward(rx.ptr, rk.ptr, 1, 1 if max_depth is not None else 0,
                                                 float(max_depth if max_depth is not None else 1.0), ry.ptr, ctypes.c_void_p(0)))
        _sync(rx)
        return y_

    if max_depth is not None:
        y = _gpu_py_function(fwd, [iconv1_linear, kernel], iconv1_linear.dtype)
        y.set_shape(iconv1_linear.shape[:3] + [1])
        return y

    @tf.custom_gradient
    def op(x, k):
        y = _gpu_py_function(fwd, [x, k], x.dtype)
        y.set_shape(x.shape[:3] + [1])

        def grad(g_out):
            def bwd(x_, k_, g_):
                C = int(x_.shape[3])
                with tf.device(x_.device):
                    g_x = tf.zeros(x_.shape, x_.dtype)
                    g_k = tf.zeros([9 * C], tf.float32)
                    ws = tf.zeros([int(lib.btslpg_depthconv_backward_workspace_bytes(C))], tf.uint8)
                rx, rk, rg, rgx, rgk, rw = _ref(x_), _ref(tf.reshape(k_, [-1])), _ref(g_), _ref(g_x), _ref(g_k), _ref(ws)
                _sync(rx)
                _cabi.check(lib.btslpg_depthconv_backward(rx.ptr, rk.ptr, rg.ptr, 1, rgx.ptr, rgk.ptr, ctypes.c_void_p(rw.struct.data),
                                                          int(ws.shape[0]), ctypes.c_void_p(0)))
                _sync(rx)
                return g_x, tf.reshape(g_k, k_.shape)
            g_x, g_k = _gpu_py_function(bwd, [x, k, g_out], [x.dtype, tf.float32])
            g_x.set_shape(x.shape)
            g_k.set_shape(k.shape)
            return g_x, g_k

        return y, grad

    return op(iconv1_linear, kernel)


def reduction_lpg(feat, kernel, upratio, ds_stride=0):  # pragma: no cover
    """bts_decoder.py:79-81 / 86-88 / 93-94 fused: the reduction head Conv2D(3, 1, activation='sigmoid', use_bias=False), its
    LocalPlanarGuidance(upratio) and the strided down-sample Lambda as ONE forward and ONE backward kernel.
    feat (B,h,w,C); kernel: the head's (1,1,C,3) kernel variable.  Returns (reduction (B,h,w,3), depth (B,H,W,1),
    depth_ds (B,H/d,W/d,1) or None); gradients flow to feat and kernel from depth and depth_ds (reduction is the
    saved tensor, as in layers.ReductionLPG)."""
    _require_tf()
    lib = _cabi.load()
    r, d = int(upratio), int(ds_stride)

    @tf.custom_gradient
    def op(x, k):
        def fwd(x_, k_):
            B, h, w = (int(v) for v in x_.shape[:3])
            with tf.device(x_.device):
                coef = tf.zeros([B, h, w, 3], x_.dtype)
                full = tf.zeros([B, h * r, w * r, 1], x_.dtype)
                ds = tf.zeros([B, h * r // d, w * r // d, 1], x_.dtype) if d else tf.zeros([0], x_.dtype)
            rx, rk, rc, ro = _ref(x_), _ref(tf.reshape(tf.cast(k_, tf.float32), [-1, 3])), _ref(coef), _ref(full)
            rd = _ref(ds) if d else None
            _sync(rx)
            _cabi.check(lib.btslpg_reduce_forward(rx.ptr, rk.ptr, r, rc.ptr, ro.ptr, rd.ptr if d else None, d, ctypes.c_void_p(0)))
            _sync(rx)
            return coef, full, ds
        coef, full, ds = _gpu_py_function(fwd, [x, k], [x.dtype] * 3)
        coef.set_shape(x.shape[:3] + [3])
        full.set_shape([x.shape[0], x.shape[1] * r, x.shape[2] * r, 1])
        if d:
            ds.set_shape([x.shape[0], x.shape[1] * r // d, x.shape[2] * r // d, 1])

        def grad(g_coef_unused, g_full, g_ds):
            def bwd(x_, k_, c_, gf_, gd_):
                C = int(x_.shape[3])
                npix = int(x_.shape[0]) * int(x_.shape[1]) * int(x_.shape[2])
                with tf.device(x_.device):
                    g_x = tf.zeros(x_.shape, x_.dtype)
                    g_k = tf.zeros([C, 3], tf.float32)
                    ws = tf.zeros([int(lib.btslpg_reduce_backward_workspace_bytes(npix, C))], tf.uint8)
                refs = [_ref(x_), _ref(tf.reshape(tf.cast(k_, tf.float32), [-1, 3])), _ref(c_), _ref(gf_), _ref(gd_) if d else None,
                        _ref(g_x), _ref(g_k), _ref(ws)]
                _sync(refs[0])
                _cabi.check(lib.btslpg_reduce_backward(refs[0].ptr, refs[1].ptr, refs[2].ptr, refs[3].ptr, refs[4].ptr if d else None, r, d,
                                                       refs[5].ptr, refs[6].ptr, None, ctypes.c_void_p(refs[7].struct.data), int(ws.shape[0]),
                                                       ctypes.c_void_p(0)))
                _sync(refs[0])
                return g_x, tf.reshape(g_k, k_.shape)
            g_x, g_k = _gpu_py_function(bwd, [x, k, coef, g_full, g_ds], [x.dtype, tf.float32])
            g_x.set_shape(x.shape)
            g_k.set_shape(k.shape)
            return g_x, g_k

        return (coef, full, ds), grad

    coef, full, ds = op(feat, kernel)
    return coef, full, (ds if d else None)


def conv3x3_tensor_core_wgrad(inputs, kernel):  # pragma: no cover
    """Conv2D(kernel_size=3, strides=1, padding='same', use_bias=False) -- upconv1 / iconv1 / conv block 2 of
    bts_decoder.py:38-44, :96-101 built with activation=None -- whose kernel gradient comes from the tcgen05 kernel
    (btslpg_conv3x3_wgrad) while the forward and the input gradient stay TensorFlow's.
    inputs (B,H,W,Cin), Cin a multiple of 4 up to 256; kernel: the layer's (3,3,Cin,Cout) variable, Cout a multiple of 4 up to 128."""
    _require_tf()
    lib = _cabi.load()

    @tf.custom_gradient
    def op(x, k):
        y = tf.nn.conv2d(x, k, strides=1, padding="SAME")

        def grad(g_out):
            g_x = tf.compat.v1.nn.conv2d_backprop_input(tf.shape(x), k, g_out, strides=[1, 1, 1, 1], padding="SAME")

            def wgrad(x_, g_):
                cin, cout = int(x_.shape[3]), int(g_.shape[3])
                with tf.device(x_.device):
                    g_k = tf.zeros([3, 3, cin, cout], tf.float32)
                    ws = tf.zeros([int(lib.btslpg_conv3x3_wgrad_workspace_bytes(cin, cout))], tf.uint8)
                rx, rg, rk, rw = _ref(x_), _ref(g_), _ref(g_k), _ref(ws)
                _sync(rx)
                _cabi.check(lib.btslpg_conv3x3_wgrad(rx.ptr, rg.ptr, rk.ptr, ctypes.c_void_p(rw.struct.data), int(ws.shape[0]), ctypes.c_void_p(0)))
                _sync(rx)
                return g_k
            g_k = _gpu_py_function(wgrad, [x, g_out], tf.float32)
            g_k.set_shape(k.shape)
            return g_x, g_k

        return y, grad

    return op(inputs, kernel)
