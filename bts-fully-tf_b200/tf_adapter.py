"""TensorFlow binding of the same C ABI -- the drop-in for the reference's own files.

TensorFlow is not installable in the build image (no wheel, no network), so this module is
import-guarded and untested here; it is the concrete form of the stub shown in INTEGRATION.md.
With TF present, `LocalPlanarGuidance` below IS a `tf.keras.layers.Layer` with the reference's
name, constructor, `build`, `call` and `get_config` (custom_layers.py:25-61), so
`from custom_layers import LocalPlanarGuidance` in bts_decoder.py:24 can be pointed at it
unchanged.  Tensors cross into libbtslpg.so as DLPack capsules (zero copy): TF allocates inputs and
outputs, the library only launches kernels on TF's buffers.

Stream note: TF runs its GPU ops on its own compute stream, which is not exposed to Python.  The
binding therefore launches on the legacy default stream (stream = NULL), which implicitly
synchronises with TF's blocking streams; `tf.experimental.dlpack.to_dlpack` itself waits for the
producer op, as DLPack requires.
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


def _forward_eager(coef, upratio, ds_stride):  # pragma: no cover
    lib = _cabi.load()
    B, h, w, _ = coef.shape
    H, W = h * upratio, w * upratio
    with tf.device(coef.device):
        full = tf.zeros((B, H, W, 1), coef.dtype)
        ds = tf.zeros((B, H // ds_stride, W // ds_stride, 1), coef.dtype) if ds_stride else None
    rc, rf = _cabi.from_dlpack_capsule(_capsule(coef)), _cabi.from_dlpack_capsule(_capsule(full))
    rd = _cabi.from_dlpack_capsule(_capsule(ds)) if ds_stride else None
    _cabi.check(lib.btslpg_forward(rc.ptr, int(upratio), rf.ptr, _cabi.ptr_or_null(rd), int(ds_stride), ctypes.c_void_p(0)))
    return full, ds


def _backward_eager(coef, g_full, g_ds, upratio, ds_stride):  # pragma: no cover
    lib = _cabi.load()
    with tf.device(coef.device):
        g_coef = tf.zeros(coef.shape, coef.dtype)
    refs = [_cabi.from_dlpack_capsule(_capsule(t)) if t is not None else None for t in (coef, g_full, g_ds, g_coef)]
    _cabi.check(lib.btslpg_backward(refs[0].ptr, _cabi.ptr_or_null(refs[1]), _cabi.ptr_or_null(refs[2]), int(upratio),
                                    int(ds_stride if g_ds is not None else 0), refs[3].ptr, ctypes.c_void_p(0)))
    return g_coef


def local_planar_guidance(inputs, upratio):  # pragma: no cover
    """Differentiable TF op: (B,h,w,3) -> (B,h*r,w*r,1); works eagerly and inside tf.function."""
    _require_tf()

    @tf.custom_gradient
    def op(x):
        def fwd(x_):
            return _forward_eager(x_, upratio, 0)[0]

        y = tf.py_function(fwd, [x], x.dtype)
        y.set_shape([x.shape[0], x.shape[1] * upratio, x.shape[2] * upratio, 1])

        def grad(dy):
            g = tf.py_function(lambda x_, dy_: _backward_eager(x_, dy_, None, upratio, 0), [x, dy], x.dtype)
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
