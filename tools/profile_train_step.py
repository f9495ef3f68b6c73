#!/usr/bin/env python
"""Where a data-parallel training step of the decoder spends its time at a given per-GPU batch (one GPU, eager, torch.profiler):
kernel count and the top kernels by device time.   python tools/profile_train_step.py [--config 5] [--per-gpu-batch 4]"""
import argparse
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
    ap.add_argument("--per-gpu-batch", type=int, default=4)
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
    eng = trainer.DataParallelStep(dec, feats, gt, dataset=cfg["dataset"], use_graph=False)
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        eng.step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        if getattr(e, "device_type", None) is not None and e.self_device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA:
            rows.append((e.key, e.count, e.self_device_time_total))
    if not rows:
        rows = [(e.key, e.count, e.self_device_time_total) for e in prof.key_averages() if e.self_device_time_total > 0]
    rows.sort(key=lambda r: -r[2])
    total = sum(r[2] for r in rows)
    print(json.dumps({"config": a.config, "per_gpu_batch": b, "kernels_launched": int(sum(r[1] for r in rows)), "device_us": round(total, 1),
                      "top": [{"name": r[0][:90], "count": int(r[1]), "us": round(r[2], 1), "share": round(r[2] / total, 3)} for r in rows[:40]]}))


if __name__ == "__main__":
    main()
