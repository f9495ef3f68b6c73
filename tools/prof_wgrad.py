#!/usr/bin/env python
"""A few launches of the tcgen05 weight-gradient kernel for ncu (tools/gpu_ncu_wgrad.sh): B = 32, 416x544, Cin -> Cout from argv
(default 32 16: upconv1 of config 5)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bts_fully_tf_b200 import ops  # noqa: E402

cin, cout = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 16)
dev = torch.device("cuda:0")
x = torch.randn(32, 416, 544, cin, device=dev)
g = torch.randn(32, 416, 544, cout, device=dev)
out = torch.empty(3, 3, cin, cout, device=dev)
for _ in range(4):
    ops.conv3x3_wgrad(x, g, out=out)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
