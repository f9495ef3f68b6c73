#!/usr/bin/env python
"""Time the multi-layer kernels: one-shot CTAs vs persistent TMA-staged warps (tuning key 8)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bts_fully_tf_b200 import ops  # noqa: E402
from bts_fully_tf_b200.host_io import DeviceSet, algorithmic_bytes  # noqa: E402
from sweep import timed  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B, H, W = 32, 480, 640
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6533.8
    out = []
    for dtype, es, name in ((torch.float32, 4, "f32"), (torch.bfloat16, 2, "bf16")):
        gen = torch.Generator(device=dev).manual_seed(0)
        sets = [DeviceSet(B, H, W, dtype, dev, generator=gen) for _ in range(4)]
        fwd_b, bwd_b, _ = algorithmic_bytes(B, H, W, es)
        for impl in (0, 1):
            ops.set_tuning(8, impl)
            f = timed(lambda s: s.forward(True), sets)
            kf = ops.last_kernel()
            b = timed(lambda s: s.backward(True), sets)
            kb = ops.last_kernel()
            out.append(dict(dtype=name, impl=impl, fwd=kf, bwd=kb, fwd_us=round(f, 2), bwd_us=round(b, 2),
                            fwd_frac=round(fwd_b / f / 1e3 / peak, 3), bwd_frac=round(bwd_b / b / 1e3 / peak, 3),
                            step_frac=round((fwd_b + bwd_b) / (f + b) / 1e3 / peak, 3)))
        ops.set_tuning(8, 0)
        del sets
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
