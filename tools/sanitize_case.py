#!/usr/bin/env python
"""Smallest shapes that reach every kernel family with shared-memory pipelines, mbarriers or last-CTA reductions, for
compute-sanitizer (one tool per gpurun call):  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py [--no-tcgen05]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bts_fully_tf_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
B, H, W = 1, 32, 64
done = []
# fused heads: TMA ring (cp.async.bulk + mbarrier) forward / backward, lane-split r = 8 backward, two-level g_kernel reduction
for r, d, C in ((8, 4, 64), (8, 4, 128), (4, 2, 64), (2, 0, 32)):
    h, w = H // r, W // r
    feat = torch.nn.functional.elu(torch.randn(B * 4, h, w, C, device=dev, generator=g))
    kern = torch.randn(C, 3, device=dev, generator=g) * 0.2
    coef, full, ds = ops.reduce_lpg_forward(feat, kern, r, d)
    done.append(ops.last_kernel())
    g_full = torch.randn(B * 4, H, W, 1, device=dev, generator=g)
    g_ds = torch.randn(B * 4, H // d, W // d, 1, device=dev, generator=g) if d else None
    ops.reduce_lpg_backward(feat, kern, coef, g_full, g_ds, r, d)
    done.append(ops.last_kernel())
# stand-alone LPG (split-patch backward exchanges partial sums through shared memory)
for r, d in ((8, 4), (4, 2), (2, 0)):
    coef = torch.rand(B, H // r, W // r, 3, device=dev, generator=g)
    full, ds = ops.lpg_forward(coef, r, d)
    ops.lpg_backward(coef, torch.randn_like(full), torch.randn_like(ds) if d else None, r, d)
    done.append(ops.last_kernel())
# concat (cp.async staging), up-sampling, last convolution forward (cp.async ring + mma.sync) / backward (last-CTA reduction)
a = torch.randn(B, H, W, 32, device=dev, generator=g)
planes = [torch.rand(B, H, W, 1, device=dev, generator=g) for _ in range(3)]
cat = ops.concat_forward(a, planes, act=True, pad=1)
done.append(ops.last_kernel())
ops.concat_backward(torch.randn_like(cat), cat, True, 32, 0, 3, pad=1)
done.append(ops.last_kernel())
ops.upsample2x_backward(ops.upsample2x_forward(a[:, ::2, ::2].contiguous()))
w9c = torch.randn(9 * 32, device=dev, generator=g) * 0.1
ops.depthconv_forward(a, w9c, act_in=True, sigmoid_scale=10.0)
done.append(ops.last_kernel())
ops.depthconv_backward(a, w9c, torch.randn(B, H, W, 1, device=dev, generator=g), act_in=True)
done.append(ops.last_kernel())
# tail: grid reductions through the workspace
logit = torch.randn(B, H, W, 1, device=dev, generator=g)
y_true = torch.rand(B, H, W, 1, device=dev, generator=g) * 10
depth, loss, ws = ops.silog_forward(logit, y_true, 10.0, 0.1)
ops.silog_backward(depth, y_true, 10.0, 0.1, ws)
ops.eval_metrics_png16(depth, 10.0, y_true=y_true, max_depth_eval=10.0)
done.append(ops.last_kernel())
# fused optimizer
p, gr, m, v = (torch.randn(4099, device=dev, generator=g)[:4096] for _ in range(4))
ops.adam_step(p, gr, m.abs(), v.abs(), ops.adam_state(dev), ops.adam_config(1e-3, total_steps=10))
done.append(ops.last_kernel())
# tcgen05 implicit GEMM (TMEM, tcgen05.mma / ld / st / commit, cp.async + mbarrier rings)
if "--no-tcgen05" not in sys.argv:
    for NF in (32, 16):
        a4 = torch.randn(B, H // 2, W // 2, 4 * NF, device=dev, generator=g)
        hwio = torch.randn(3, 3, NF + 3, NF, device=dev, generator=g) * 0.05
        ops.iconv1_forward(a4, planes, hwio, a_subpixel=True)
        done.append(ops.last_kernel())
torch.cuda.synchronize()
print("SANITIZE_CASE_OK", len(done), "kernel families:", ", ".join(done))
