#!/usr/bin/env python
"""Experiment: the library's weight gradient of the DenseASPP dilated convolutions (bts_decoder.py:53, rates 3..24) as it is called
(padding = rate) against the same gradient on a pre-padded input with padding = 0, and the input gradient, per rate."""
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
out = []
for name, B, H, W, Cin, Cout in (("cfg4", 32, 44, 152, 256, 128), ("cfg5", 32, 52, 68, 128, 64)):
    for d in (3, 6, 12, 18, 24):
        x = [torch.randn(B, Cin, H, W, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(2)]
        g = [torch.randn(B, Cout, H, W, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(2)]
        w = torch.randn(Cout, Cin, 3, 3, device=dev).contiguous(memory_format=torch.channels_last)
        xp = [F.pad(t, (d, d, d, d)).contiguous(memory_format=torch.channels_last) for t in x]

        def a(k):
            torch.ops.aten.convolution_backward(g[k], x[k], w, None, [1, 1], [d, d], [d, d], False, [0, 0], 1, [False, True, False])

        def b(k):
            torch.ops.aten.convolution_backward(g[k], xp[k], w, None, [1, 1], [0, 0], [d, d], False, [0, 0], 1, [False, True, False])

        def c(k):
            torch.ops.aten.convolution_backward(g[k], x[k], w, None, [1, 1], [d, d], [d, d], False, [0, 0], 1, [True, False, False])

        def f(k):
            F.conv2d(x[k], w, None, 1, d, d)

        ra = torch.ops.aten.convolution_backward(g[0], x[0], w, None, [1, 1], [d, d], [d, d], False, [0, 0], 1, [False, True, False])[1]
        rb = torch.ops.aten.convolution_backward(g[0], xp[0], w, None, [1, 1], [0, 0], [d, d], False, [0, 0], 1, [False, True, False])[1]
        out.append({"case": name, "rate": d, "wgrad_us": round(timed(a, 2, reps=3), 1), "wgrad_prepadded_us": round(timed(b, 2, reps=3), 1),
                    "dgrad_us": round(timed(c, 2, reps=3), 1), "fprop_us": round(timed(f, 2, reps=3), 1),
                    "rel_diff": float((ra - rb).abs().max() / ra.abs().max())})
print(json.dumps(out))
