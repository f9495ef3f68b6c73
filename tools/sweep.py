#!/usr/bin/env python
"""GPU tuning sweep: per-kernel time / algorithmic GB/s of every LPG launch for several block sizes.
Writes one JSON document to stdout.  Timing: CUDA-graph replays over rotating buffer sets (> L2),
CUDA events, 100 launches per point."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bts_fully_tf_b200 import ops  # noqa: E402
from bts_fully_tf_b200.host_io import DeviceSet, algorithmic_bytes  # noqa: E402


def timed(fn_per_set, sets, n=100, warm=10):
    graphs = []
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for s in sets:
            fn_per_set(s)
        per_graph = 5 * len(sets)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for k in range(per_graph):
                fn_per_set(sets[k % len(sets)])
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    reps = max(1, n // per_graph)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * per_graph) * 1e3   # us per launch, back to back inside one graph


def main():
    dev = torch.device("cuda:0")
    B, H, W = 32, 480, 640
    peak = 6533.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {"peak": peak, "points": []}
    for dtype, es, name in ((torch.float32, 4, "f32"), (torch.bfloat16, 2, "bf16")):
        gen = torch.Generator(device=dev).manual_seed(0)
        sets = [DeviceSet(B, H, W, dtype, dev, generator=gen) for _ in range(4)]
        fwd_b, bwd_b, per = algorithmic_bytes(B, H, W, es)
        configs = [(128, 0, 0), (64, 0, 0)]
        if name == "f32":
            configs += [(128, 8, 0), (128, 2, 0), (128, 0, 2)]      # r8 rows-per-lane 8 / 2 for both directions, r4 px 2
        for threads, r8rows, r4px in configs:
            ops.set_block_threads(threads, threads)
            ops.set_tuning(2, r8rows)
            ops.set_tuning(3, r4px)
            tag_cfg = "t%d" % threads + ("_r8rows%d" % r8rows if r8rows else "") + ("_r4px%d" % r4px if r4px else "")
            for idx, (r, fb, bb) in enumerate(per):
                def f(s, idx=idx):
                    L = s.layers[idx]
                    ops.lpg_forward(L["coef"], L["upratio"], L["ds_stride"], out_full=L["out_full"], out_ds=L["out_ds"])

                def b(s, idx=idx):
                    L = s.layers[idx]
                    ops.lpg_backward(L["coef"], L["g_full"], L["g_ds"], L["upratio"], L["ds_stride"], g_coef=L["g_coef"])
                for tag, fn, nb in (("fwd", f, fb), ("bwd", b, bb)):
                    us = timed(fn, sets)
                    out["points"].append(dict(dtype=name, kernel="%s_r%d" % (tag, r), threads=tag_cfg, variant=ops.last_kernel(), us=round(us, 2),
                                              GBps=round(nb / us / 1e3, 1), frac=round(nb / us / 1e3 / peak, 3)))
            for tag, fn, nb in (("fwd_multi", lambda s: s.forward(True), fwd_b), ("bwd_multi", lambda s: s.backward(True), bwd_b),
                                ("fwd_3launch", lambda s: s.forward(False), fwd_b), ("bwd_3launch", lambda s: s.backward(False), bwd_b)):
                us = timed(fn, sets)
                out["points"].append(dict(dtype=name, kernel=tag, threads=tag_cfg, us=round(us, 2),
                                          GBps=round(nb / us / 1e3, 1), frac=round(nb / us / 1e3 / peak, 3)))
        ops.set_block_threads(0, 0)
        ops.set_tuning(2, 0)
        ops.set_tuning(3, 0)
        del sets
        torch.cuda.empty_cache()
    # plain device copy of the same size as a reference point for this timing method
    a = torch.empty(64 * 1024 * 1024, device=dev)
    bufs = [(torch.empty_like(a), torch.empty_like(a)) for _ in range(3)]
    for x, y in bufs:
        y.copy_(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(30):
        x, y = bufs[k % 3]
        y.copy_(x)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    out["copy_256MB"] = dict(us=round(us, 1), GBps=round(2 * a.numel() * 4 / us / 1e3, 1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
