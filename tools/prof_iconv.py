#!/usr/bin/env python
"""A few launches of the tcgen05 iconv1 kernel at B = 8, 480x640 (for ncu)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bts_fully_tf_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, H, W, NF = 8, 480, 640, int(sys.argv[1]) if len(sys.argv) > 1 else 32
a4 = torch.randn(B, H // 2, W // 2, 4 * NF, device=dev)
planes = [torch.rand(B, H, W, 1, device=dev) * 10 for _ in range(3)]
hwio = torch.randn(3, 3, NF + 3, NF, device=dev) * 0.05
out = torch.empty(B, H, W, NF, device=dev)
for _ in range(4):
    ops.iconv1_forward(a4, planes, hwio, a_subpixel=True, out=out)
torch.cuda.synchronize()
print("ok", ops.last_kernel())
