#!/usr/bin/env python
"""Experiment: upsample x2 (nearest) + conv3x3  ==  conv3x3 on the low-res input with 4*Cout channels + pixel shuffle.
Checks the identity numerically and times both forms (forward and forward+backward) on cuDNN, channels_last, TF32."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bts_fully_tf_b200 import ops  # noqa: E402

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")


def effective_kernel(w):
    """w (Cout, Cin, 3, 3) -> (4*Cout, Cin, 3, 3): output parity (a, b) of the up-sampled conv as a 3x3 conv on the low-res input."""
    Cout, Cin = w.shape[:2]
    # row combination matrices: parity a -> (3 low-res taps) x (3 original taps)
    R = torch.zeros(2, 3, 3, device=w.device, dtype=w.dtype)
    R[0, 0, 0] = 1; R[0, 1, 1] = 1; R[0, 1, 2] = 1          # even output row: low-res rows (y-1: w0), (y: w1 + w2)
    R[1, 1, 0] = 1; R[1, 1, 1] = 1; R[1, 2, 2] = 1          # odd output row:  (y: w0 + w1), (y+1: w2)
    # w_eff[a, b, o, c, i, j] = sum_{k, l} R[a, i, k] * R[b, j, l] * w[o, c, k, l]
    weff = torch.einsum("aik,bjl,ockl->abocij", R, R, w)
    return weff.reshape(4 * Cout, Cin, 3, 3)


def timed(fn, reps=10):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
for name, (B, h, w_, cin, cout) in {"upconv1_nyu_b32": (32, 240, 320, 64, 32), "block2_upconv_nyu_b32": (32, 120, 160, 128, 64),
                                    "block3_upconv_nyu_b32": (32, 60, 80, 128, 128), "upconv1_kitti_b16": (16, 176, 608, 64, 32)}.items():
    x = torch.randn(B, cin, h, w_, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wt = (torch.randn(cout, cin, 3, 3, device=dev) * 0.05).requires_grad_(True)

    def ref():
        up = ops.upsample2x_nhwc(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        return F.conv2d(up, wt, padding=1)

    def sub():
        y4 = F.conv2d(x, effective_kernel(wt), padding=1)            # (B, 4*Cout, h, w), channel = (a, b, o)
        return y4

    def sub_shuffled():
        y4 = sub()
        B_, _, hh, ww = y4.shape
        return y4.view(B_, 2, 2, cout, hh, ww).permute(0, 3, 4, 1, 5, 2).reshape(B_, cout, 2 * hh, 2 * ww)

    with torch.no_grad():
        a, b_ = ref(), sub_shuffled()
        err = float((a - b_).abs().max() / a.abs().max())
    g = torch.randn(B, cout, 2 * h, 2 * w_, device=dev).contiguous(memory_format=torch.channels_last)
    g4 = torch.randn(B, 4 * cout, h, w_, device=dev).contiguous(memory_format=torch.channels_last)

    def ref_fb():
        x.grad = wt.grad = None
        ref().backward(g)

    def sub_fb():
        x.grad = wt.grad = None
        sub().backward(g4)

    with torch.no_grad():
        t_ref_f, t_sub_f = timed(ref), timed(sub)
    t_ref_fb, t_sub_fb = timed(ref_fb, 5), timed(sub_fb, 5)
    out[name] = {"rel_err": err, "fwd_ms": {"upsample+conv": round(t_ref_f, 3), "lowres_conv_4x": round(t_sub_f, 3)},
                 "fwd_bwd_ms": {"upsample+conv": round(t_ref_fb, 3), "lowres_conv_4x (no shuffle)": round(t_sub_fb, 3)}}
    del x, wt, g, g4
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
