"""Keras-shaped layer surface of the hot path (the drop-in boundary of SURVEY 8(b)).

`LocalPlanarGuidance(upratio, name=...)` keeps the reference's class name, constructor, lazy
`build(input_shape)`, `call(inputs)`, `__call__` and `get_config()` (custom_layers.py:25-61); the
body is one launch of the sm_100a kernel behind the C ABI instead of ~20 tf ops.  It is a
`torch.nn.Module` so that it slots into the torch/cuDNN glue that hosts the rest of the decoder
on this box; the TensorFlow-side binding of the same C ABI lives in tf_adapter.py.

`ReductionLPG` is the fused form of `reduction_NxN` + LPG + down-sampling Lambda
(bts_decoder.py:79-81, 86-88, 93-94) and owns the Conv2D kernel in Keras' HWIO layout.
"""
import math

import torch

from . import ops


class LocalPlanarGuidance(torch.nn.Module):
    """depth[b, y, x] = n4 / (n . pixel_dir_unit[y % r, x % r] + K.epsilon()),  (n, n4) decoded from
    inputs[b, y//r, x//r] = [phi/(2 pi), theta/(pi/3), dist]          (reference custom_layers.py:47-56)

    inputs (B, h, w, 3) NHWC -> (B, h*r, w*r, 1).  `ds_stride` (not in the reference class) additionally
    returns depth[:, ::ds_stride, ::ds_stride] from the same launch -- the Lambda of bts_decoder.py:81,88.
    """

    def __init__(self, upratio, ds_stride=0, **kwargs):
        super().__init__()
        # Keras Layer kwargs that the reference forwards to layers.Layer (custom_layers.py:26-27)
        self.layer_name = kwargs.pop("name", None) or "local_planar_guidance"
        self.trainable = kwargs.pop("trainable", True)
        self.layer_dtype = kwargs.pop("dtype", "float32")
        if kwargs:
            raise TypeError("Keyword argument not understood: %s" % sorted(kwargs))
        self.upratio = int(upratio)
        self.ds_stride = int(ds_stride)
        self.built = False
        self.input_hw = None

    @property
    def name(self):
        return self.layer_name

    def build(self, input_shape):
        assert len(input_shape) > 2                       # custom_layers.py:31
        h, w = input_shape[1], input_shape[2]
        if h is None or w is None:
            raise ValueError("LocalPlanarGuidance needs static height and width (custom_layers.py:32)")
        # the reference materialises a (1, h*r, w*r, 3) constant here; the kernels hold its r*r distinct
        # vectors in __constant__ memory (csrc/lpg_dir_tables.h), so there is nothing to allocate
        self.input_hw = (int(h), int(w))
        self.built = True

    def compute_output_shape(self, input_shape):
        return (input_shape[0], input_shape[1] * self.upratio, input_shape[2] * self.upratio, 1)

    def call(self, inputs):
        return ops.local_planar_guidance(inputs, self.upratio, self.ds_stride)

    def forward(self, inputs):
        if not self.built:
            self.build(tuple(inputs.shape))
        elif tuple(inputs.shape[1:3]) != self.input_hw:
            raise ValueError("%s was built for coarse size %s, got %s" % (self.name, self.input_hw, tuple(inputs.shape[1:3])))
        return self.call(inputs)

    def get_config(self):                                  # custom_layers.py:58-61
        base_config = {"name": self.name, "trainable": self.trainable, "dtype": self.layer_dtype}
        config = {"upratio": self.upratio}
        if self.ds_stride:
            config["ds_stride"] = self.ds_stride
        return dict(list(base_config.items()) + list(config.items()))

    @classmethod
    def from_config(cls, config):
        return cls(**config)

    def extra_repr(self):
        return "upratio=%d, ds_stride=%d, name=%r" % (self.upratio, self.ds_stride, self.name)


class ReductionLPG(torch.nn.Module):
    """reduction_NxN (Conv2D(3, 1x1, sigmoid, use_bias=False)) + LocalPlanarGuidance(upratio) + the
    strided-slice Lambda, as ONE kernel forward and ONE kernel backward.

    forward(feat NHWC (B,h,w,C)) -> (reduction (B,h,w,3), depth (B,H,W,1)[, depth_ds])
    `kernel` is the Keras Conv2D kernel, HWIO (1,1,C,3), glorot_uniform like the Keras default.
    """

    def __init__(self, in_channels, upratio, ds_stride=0, name=None):
        super().__init__()
        self.upratio, self.ds_stride = int(upratio), int(ds_stride)
        self.layer_name = name or "reduction_%dx%d" % (upratio, upratio)
        limit = math.sqrt(6.0 / (in_channels + 3))        # glorot_uniform, fan_in = C, fan_out = 3
        self.kernel = torch.nn.Parameter((torch.rand(1, 1, in_channels, 3) * 2 - 1) * limit)
        self._grad_view = None
        self._grad_written = None

    def bind_gradient_view(self, view, on_written=None):
        """Let backward write d loss / d kernel straight into `view` (a float32 slice of a flat gradient
        bucket, see parallel.GradientBucket / trainer.FlatState) instead of handing it to autograd for accumulation.
        The kernel OVERWRITES the slice (it does not accumulate: zero-initialised buckets and one backward per step).
        `on_written()` is called right after the backward kernel has been enqueued -- the bucket's cue that this
        parameter's gradient is complete (autograd's own hooks never fire for it, since it sees no gradient)."""
        if view is not None and (view.dtype != torch.float32 or view.numel() != self.kernel.numel() or not view.is_contiguous()):
            raise ValueError("gradient view must be a contiguous float32 tensor with %d elements" % self.kernel.numel())
        self._grad_view = view
        self._grad_written = on_written if view is not None else None

    @property
    def name(self):
        return self.layer_name

    def forward(self, feat):
        return ops.reduce_lpg(feat, self.kernel, self.upratio, self.ds_stride, self._grad_view, self._grad_written)
