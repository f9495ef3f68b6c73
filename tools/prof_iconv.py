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
if "--roles" in sys.argv:
    import ctypes
    from bts_fully_tf_b200 import _cabi
    lib = _cabi.load()
    lib.btslpg_debug_iconv1_profile.argtypes = [ctypes.c_void_p]
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.btslpg_debug_iconv1_profile(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.iconv1_forward(a4, planes, hwio, a_subpixel=True, out=out)
    e1.record()
    torch.cuda.synchronize()
    lib.btslpg_debug_iconv1_profile(ctypes.c_void_p(0))
    v = buf.cpu().tolist()
    n = 148
    names = ["mma_wait_full", "mma_wait_acc_empty", "mma_total", "prod_wait_empty", "prod_issue", "prod_wait_landed", "prod_activate",
             "prod_total", "epi_wait_acc_full", "epi_total"]
    print("kernel ms", e0.elapsed_time(e1))
    for k, name in enumerate(names):
        print("%-20s %10.0f cycles per CTA" % (name, v[k] / n))
