"""TEST INFRASTRUCTURE ONLY -- op-by-op torch-CPU restatement of
/root/reference/custom_layers.py:25-61 and the reduction heads / slices of
/root/reference/bts_decoder.py:79-94.

It materialises the same intermediates the tf.keras graph does (constant
(1,H,W,3) direction tensor, two repeat_elements expansions, product, sum, add,
divide) and takes its backward from autograd, exactly as TF autodiff would.  It is
(a) one of the two restatements the tests cross-check, and (b) the "port" that
bench.py times on the host cores as the CPU baseline, TensorFlow being absent.
"""
from math import pi

import numpy as np
import torch

K_EPSILON = 1e-7


class LocalPlanarGuidanceLiteral:
    def __init__(self, upratio, name=None):           # custom_layers.py:26-28
        self.upratio = upratio
        self.name = name
        self.pixel_dir_unit = None

    def build(self, input_shape):                      # custom_layers.py:30-45
        assert len(input_shape) > 2
        r = self.upratio
        height, width = input_shape[1] * r, input_shape[2] * r
        v, u = np.meshgrid(np.linspace(0, width - 1, width, dtype=np.float32),
                           np.linspace(0, height - 1, height, dtype=np.float32))
        v = torch.from_numpy(v)[None]
        v = (v % r - (r - 1) / 2) / float(r)
        u = torch.from_numpy(u)[None]
        u = (u % r - (r - 1) / 2) / float(r)
        x = torch.stack([u, v, torch.ones_like(u)], dim=-1)
        # K.l2_normalize: x * rsqrt(max(sum(x^2), 1e-12))
        self.pixel_dir_unit = x * torch.rsqrt(torch.clamp_min((x * x).sum(dim=3, keepdim=True), 1e-12))

    def __call__(self, inputs):                        # custom_layers.py:47-56
        if self.pixel_dir_unit is None:
            self.build(tuple(inputs.shape))
        r = self.upratio
        phi, theta, raw_dist = inputs[:, :, :, 0:1] * 2 * pi, inputs[:, :, :, 1:2] * pi / 3, inputs[:, :, :, 2:3]
        plane_coeffs = torch.cat([torch.sin(theta) * torch.cos(phi), torch.sin(theta) * torch.sin(phi),
                                  torch.cos(theta), raw_dist], dim=-1)
        plane_exp_height = torch.repeat_interleave(plane_coeffs, r, dim=1)
        plane_exp = torch.repeat_interleave(plane_exp_height, r, dim=2)
        dirs = self.pixel_dir_unit.to(plane_exp.dtype)
        denominator = (dirs * plane_exp[..., 0:3]).sum(dim=-1, keepdim=True) + K_EPSILON
        return plane_exp[..., 3:] / denominator


def reduction_head(feat, kernel):
    """bts_decoder.py:79,86,93: Conv2D(3, 1x1, sigmoid, use_bias=False); kernel [C][3] (HWIO squeezed)."""
    return torch.sigmoid(feat @ kernel)


def downsample(x, d):
    """bts_decoder.py:81,88: x[:, ::d, ::d, ...]."""
    return x[:, ::d, ::d, ...]


def lpg_fwd_bwd(coef, g_full, r, g_ds=None, d=0):
    """Forward + autograd backward of one LPG layer (and its strided slice) on torch-CPU.
    Returns (out, out_ds or None, g_coef)."""
    coef = coef.detach().clone().requires_grad_(True)
    layer = LocalPlanarGuidanceLiteral(r)
    out = layer(coef)
    outs, grads = [out], [g_full.reshape(out.shape)]
    ds = None
    if g_ds is not None:
        ds = downsample(out, d)
        outs.append(ds)
        grads.append(g_ds.reshape(ds.shape))
    torch.autograd.backward(outs, grads)
    return out.detach(), (ds.detach() if ds is not None else None), coef.grad
