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
