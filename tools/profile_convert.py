#!/usr/bin/env python
"""Which convolutions of a training step make cuDNN launch layout-conversion kernels (convertTensor / nchwToNhwc): torch.profiler with
shapes, kernels attributed to their aten parent.  python tools/profile_convert.py [--config 5] [--per-gpu-batch 32]"""
import argparse
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_decoder  # noqa: E402
from bts_fully_tf_b200 import trainer  # noqa: E402
from bts_fully_tf_b200.decoder import BtsDecoder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=5)
    ap.add_argument("--per-gpu-batch", type=int, default=32)
    ap.add_argument("--inference", action="store_true", help="one eval-mode forward under no_grad instead of a training step")
    a = ap.parse_args()
    cfg = bench_decoder.CONFIGS[a.config]
    chans, F = bench_decoder.TAPS[cfg["encoder"]]
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    b, H, W = a.per_gpu_batch, cfg["H"], cfg["W"]
    torch.manual_seed(0)
    dec = BtsDecoder(chans, cfg["max_depth"], num_filters=F).to(dev)
    feats = [torch.relu(torch.randn(b, H // s, W // s, c, device=dev)) for s, c in zip((32, 2, 4, 8, 16), chans)]
    gt = torch.rand(b, H, W, 1, device=dev) * cfg["max_depth"]
    if a.inference:
        dec.eval()

        def step():
            with torch.no_grad():
                dec(feats)
    else:
        eng = trainer.DataParallelStep(dec, feats, gt, dataset=cfg["dataset"], use_graph=False)
        step = eng.step
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU], record_shapes=True) as prof:
        step()
        torch.cuda.synchronize()
    # attribute each kernel to the innermost CPU op that encloses its launch
    events = prof.events()
    cpu_ops = [e for e in events if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("aten::")]
    out = collections.defaultdict(lambda: [0, 0.0])
    for e in cpu_ops:
        if "conv" not in e.name:
            continue
        for k in e.kernels:
            key = (e.name, str(e.input_shapes)[:160], k.name[:60])
            out[key][0] += 1
            out[key][1] += k.duration
    rows = sorted(((v[1], v[0], k) for k, v in out.items()), reverse=True)
    print(json.dumps([{"us": round(t, 1), "n": n, "op": k[0], "shapes": k[1], "kernel": k[2]} for t, n, k in rows[:60]]))


if __name__ == "__main__":
    main()
