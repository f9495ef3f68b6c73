#!/usr/bin/env python
"""A few launches of the last convolution's backward kernel for ncu: B = 32, 480x640, C from argv (default 16)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bts_fully_tf_b200 import ops  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
x = torch.randn(32, 480, 640, C, device=dev)
w = torch.randn(1, C, 3, 3, device=dev) * 0.1
g = torch.randn(32, 480, 640, 1, device=dev)
k9 = ops.kernel9c(w)
for _ in range(4):
    ops.depthconv_backward(x, k9, g, act_in=True)
torch.cuda.synchronize()
print("ok")
