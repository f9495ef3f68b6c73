#!/usr/bin/env python
"""Experiment: DenseASPP 1x1 convolution + folded BatchNormalization + ReLU in inference (bts_decoder.py:49-52) as
(A) library convolution + one in-place ops.affine_act pass (the shipped form) against (B) the library's fused
convolution + bias + ReLU entry with the scale folded into the kernel."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_iconv import timed  # noqa: E402
from bts_fully_tf_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
out = []
B, H, W, Cout = 32, 60, 80, 256
for Cin in (256, 384, 512, 640, 768):
    x = [torch.relu(torch.randn(B, H, W, Cin, device=dev)).permute(0, 3, 1, 2) for _ in range(2)]
    w = (torch.randn(Cout, Cin, 1, 1, device=dev) * 0.05).contiguous(memory_format=torch.channels_last)
    scale, shift = torch.rand(Cout, device=dev) + 0.5, torch.randn(Cout, device=dev)
    wf = (w * scale[:, None, None, None]).contiguous(memory_format=torch.channels_last)

    def a(k):
        y = torch.nn.functional.conv2d(x[k], w).contiguous(memory_format=torch.channels_last)
        yn = y.permute(0, 2, 3, 1)
        ops.affine_act(yn, dst=yn, scale=scale, shift=shift, act=ops.ACT_RELU)
        return y

    def b(k):
        return torch.cudnn_convolution_relu(x[k], wf, shift, [1, 1], [0, 0], [1, 1], 1)

    with torch.no_grad():
        ya, yb = a(0), b(0)
        rel = float((ya - yb).abs().max() / ya.abs().max())
        ta = timed(lambda k: a(k), 2, reps=4)
        tb = timed(lambda k: b(k), 2, reps=4)
    out.append({"Cin": Cin, "conv_plus_affine_act_us": round(ta, 1), "fused_bias_relu_us": round(tb, 1), "rel_diff": rel})
print(json.dumps(out))
