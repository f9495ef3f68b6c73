#!/usr/bin/env python
"""cuDNN cost of the full-resolution tail convolutions (forward, dgrad, wgrad separately) next to their HBM floors."""
import json
import torch
import torch.nn.functional as F
from torch.profiler import ProfilerActivity, profile

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
out = {}
for tag, (B, H, W) in {"nyu_b32": (32, 480, 640), "kitti_b16": (16, 352, 1216)}.items():
    for name, cin, cout in (("depth_conv", 32, 1), ("iconv1", 36, 32), ("upconv1", 64, 32)):
        x = torch.randn(B, cin, H, W, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        w = torch.randn(cout, cin, 3, 3, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        g = torch.randn(B, cout, H, W, device=dev).contiguous(memory_format=torch.channels_last)
        for _ in range(3):
            y = F.conv2d(x, w, padding=1)
            y.backward(g)
            x.grad = w.grad = None
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            y = F.conv2d(x, w, padding=1)
            y.backward(g)
            torch.cuda.synchronize()
        rows = sorted(((e.device_time_total, e.key[:70]) for e in prof.key_averages() if e.device_time_total), reverse=True)[:6]
        px = B * H * W
        out["%s/%s" % (tag, name)] = {"kernels_us": [(round(t, 1), k) for t, k in rows],
                                      "floor_us": {"fwd": round((cin + cout) * 4 * px / 6.5e6, 1), "dgrad": round((cin + cout) * 4 * px / 6.5e6, 1),
                                                   "wgrad": round((cin + cout) * 4 * px / 6.5e6, 1)}}
        if name == "depth_conv":
            import os, sys
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            from bts_fully_tf_b200 import ops
            xn = x.detach().permute(0, 2, 3, 1).contiguous()
            gn = g.permute(0, 2, 3, 1).contiguous()
            k9c = w.detach().permute(2, 3, 1, 0).reshape(-1).float().contiguous()
            ops.depthconv_backward(xn, k9c, gn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.depthconv_backward(xn, k9c, gn)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 100
            nb = (2 * cin + 1) * 4 * px
            out["%s/%s" % (tag, name)]["ours_fused_backward"] = {"us": round(us, 1), "GBps": round(nb / us * 1e-3, 1), "kernel": ops.last_kernel()}
            del xn, gn
        del x, w, g, y
        torch.cuda.empty_cache()
print(json.dumps(out))
