#!/usr/bin/env python
"""Experiment: DenseASPP dilated convolutions of rate 18 / 24 (bts_decoder.py:53) as s x s interleaved sub-grids with rate d / s
(space-to-batch, the exact same arithmetic), against the library's own rate-d kernels: forward and both gradients, NHWC tensors."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_iconv import timed  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True


def s2b(x_nhwc, s):
    B, H, W, C = x_nhwc.shape
    Hp, Wp = -(-H // s) * s, -(-W // s) * s
    if (Hp, Wp) != (H, W):
        x_nhwc = F.pad(x_nhwc, (0, 0, 0, Wp - W, 0, Hp - H))
    return x_nhwc.reshape(B, Hp // s, s, Wp // s, s, C).permute(0, 2, 4, 1, 3, 5).reshape(B * s * s, Hp // s, Wp // s, C)


def b2s(y, s, B, H, W):
    _, h, w, C = y.shape
    return y.reshape(B, s, s, h, w, C).permute(0, 3, 1, 4, 2, 5).reshape(B, h * s, w * s, C)[:, :H, :W]


def conv_nhwc(x, w, d):
    return F.conv2d(x.permute(0, 3, 1, 2), w, None, 1, d, d).permute(0, 2, 3, 1)


out = []
for name, B, H, W, Cin, Cout in (("cfg4", 32, 44, 152, 256, 128), ("cfg3", 32, 60, 80, 256, 128), ("cfg5", 32, 52, 68, 128, 64)):
    for d, s in ((18, 3), (24, 2), (18, 2), (24, 3), (24, 4), (12, 2)):
        x = torch.randn(B, H, W, Cin, device=dev)
        w = (torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        g = torch.randn(B, H, W, Cout, device=dev)
        xr = x.clone().requires_grad_(True)

        def direct(k):
            w.grad = None
            xr.grad = None
            conv_nhwc(xr, w, d).backward(g)

        def split(k):
            w.grad = None
            xr.grad = None
            b2s(conv_nhwc(s2b(xr, s), w, d // s), s, B, H, W).backward(g)

        def direct_f(k):
            with torch.no_grad():
                conv_nhwc(x, w, d)

        def split_f(k):
            with torch.no_grad():
                b2s(conv_nhwc(s2b(x, s), w, d // s), s, B, H, W)

        direct(0)
        ga, gx = w.grad.clone(), xr.grad.clone()
        split(0)
        rel_w = float((w.grad - ga).abs().max() / ga.abs().max())
        rel_x = float((xr.grad - gx).abs().max() / gx.abs().max())
        with torch.no_grad():
            rel_y = float((conv_nhwc(x, w, d) - b2s(conv_nhwc(s2b(x, s), w, d // s), s, B, H, W)).abs().max())
        out.append({"case": name, "rate": d, "s": s, "direct_fwd_bwd_us": round(timed(direct, 1, reps=3), 1), "split_fwd_bwd_us": round(timed(split, 1, reps=3), 1),
                    "direct_fwd_us": round(timed(direct_f, 1, reps=3), 1), "split_fwd_us": round(timed(split_f, 1, reps=3), 1),
                    "rel_w": rel_w, "rel_x": rel_x, "abs_y": rel_y})
print(json.dumps(out))
