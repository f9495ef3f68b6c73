"""TEST INFRASTRUCTURE ONLY -- a torch-CPU stand-in for the handful of TensorFlow
symbols the reference's hot-path files touch.

Why this exists: TensorFlow is not installed in the build image and cannot be
installed (no wheel, no network), and the reference ships no tests or golden
vectors.  To pin the oracle to the *reference's own source* rather than to a
re-typed copy of it, ``tests/golden/make_golden.py`` puts this directory on
``sys.path`` and then imports the UNMODIFIED ``/root/reference/custom_layers.py``
and ``/root/reference/bts_decoder.py``.  Every line of the reference's layer code
then executes as written; only the primitive ops underneath (sin, cos, concat,
repeat_elements, sum, l2_normalize, meshgrid, Conv2D ...) are supplied here, by
torch-CPU ops that follow TensorFlow's documented semantics.

What that pins and what it does not: it pins the reference's algorithm (op
order, channel order, axis pairing, epsilon placement, strided slices, concat
order).  It does NOT pin TensorFlow's own kernels' last-ulp rounding (Eigen
sin/cos/rsqrt); DESIGN.md states this as "parity pinned to the reference source
over a stand-in runtime; TF kernel rounding unpinned".

Nothing in the product package imports this module.
"""
import numpy as _np
import torch as _torch

from . import keras  # noqa: F401

float32 = _torch.float32
float64 = _torch.float64


def _t(x):
    if isinstance(x, _torch.Tensor):
        return x
    return _torch.from_numpy(_np.ascontiguousarray(x))


def meshgrid(*args, indexing="xy"):
    """tf.meshgrid: default indexing='xy' (same as numpy)."""
    arrs = [a.numpy() if isinstance(a, _torch.Tensor) else _np.asarray(a) for a in args]
    return [_t(g.copy()) for g in _np.meshgrid(*arrs, indexing=indexing)]


def function(*a, **k):  # tf.function() decorator: eager is fine for an oracle
    def deco(f):
        return f
    if len(a) == 1 and callable(a[0]) and not k:
        return a[0]
    return deco


def boolean_mask(tensor, mask):
    return tensor[mask]


# ---- symbols used by /root/reference/custom_eval_metrics.py (TensorFlow documentation semantics) ----
def logical_and(a, b):
    return _torch.logical_and(a, b)


def where(condition, x, y):
    """tf.where(cond, x, y) with broadcasting; python scalars take the other operand's dtype."""
    if not isinstance(x, _torch.Tensor):
        x = _torch.tensor(x, dtype=y.dtype)
    if not isinstance(y, _torch.Tensor):
        y = _torch.tensor(y, dtype=x.dtype)
    return _torch.where(condition, x, y)


def clip_by_value(t, clip_value_min, clip_value_max):
    return _torch.clamp(t, min=clip_value_min, max=clip_value_max)


def reduce_mean(x):
    return _torch.mean(x)


def cast(x, dtype):
    return x.to(dtype)


def maximum(a, b):
    return _torch.maximum(a, b)


def sqrt(x):
    return _torch.sqrt(x)


def abs(x):  # noqa: A001 - mirrors the tf name
    return _torch.abs(x)


def zeros_like(x):
    return _torch.zeros_like(x)


class math:  # noqa: N801 - tf.math namespace
    @staticmethod
    def is_finite(x):
        return _torch.isfinite(x)

    @staticmethod
    def log(x):
        if not isinstance(x, _torch.Tensor):
            x = _torch.tensor(x, dtype=_torch.float64)
        return _torch.log(x)
