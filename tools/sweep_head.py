#!/usr/bin/env python
"""Time the fused reduction-head + LPG kernels (bts_decoder.py:79-94) at BASELINE config-2 shapes
(B=32, 480x640) for the densenet161 (C = 128/128/64) and resnet50 (C = 64/64/32) decoders.
Algorithmic bytes per BASELINE.md section 3.  One JSON document on stdout."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bts_fully_tf_b200 import ops  # noqa: E402


def timed(fn, nsets, n=40):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for k in range(nsets):
            fn(k)
        per_graph = 5 * nsets
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for k in range(per_graph):
                fn(k % nsets)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    reps = max(1, n // per_graph)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * per_graph) * 1e3


def collect(dtypes=("f32", "bf16"), encoders=("densenet161", "resnet50"), only_r=None, B=32, H=480, W=640, device=None, n=40):
    """Time every fused head forward / backward of the given decoders; returns {"peak": ..., "points": [...]}."""
    dev = device or torch.device("cuda:0")
    peak = 6533.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {"peak": peak, "points": []}
    nsets = 2
    all_enc = {"densenet161": (128, 128, 64), "resnet50": (64, 64, 32)}
    for dtype, es, name in ((torch.float32, 4, "f32"), (torch.bfloat16, 2, "bf16")):
        if name not in dtypes:
            continue
        for enc in encoders:
            chans = all_enc[enc]
            for (r, d), C in zip(((8, 4), (4, 2), (2, 0)), chans):
                if only_r is not None and r != only_r:
                    continue
                h, w = H // r, W // r
                gen = torch.Generator(device=dev).manual_seed(0)
                feats = [torch.nn.functional.elu(torch.randn(B, h, w, C, device=dev, generator=gen)).to(dtype) for _ in range(nsets)]
                kern = ((torch.rand(C, 3, device=dev, generator=gen) * 2 - 1) * (6.0 / (C + 3)) ** 0.5)
                g_full = [torch.randn(B, H, W, 1, device=dev, generator=gen).to(dtype) for _ in range(nsets)]
                g_ds = [torch.randn(B, H // d, W // d, 1, device=dev, generator=gen).to(dtype) for _ in range(nsets)] if d else None
                coef = [torch.empty(B, h, w, 3, device=dev, dtype=dtype) for _ in range(nsets)]
                full = [torch.empty(B, H, W, 1, device=dev, dtype=dtype) for _ in range(nsets)]
                ds = [torch.empty(B, H // d, W // d, 1, device=dev, dtype=dtype) for _ in range(nsets)] if d else None
                gk = torch.empty(C, 3, device=dev)

                def fwd(k):
                    ops.reduce_lpg_forward(feats[k], kern, r, d, out_full=full[k], out_ds=ds[k] if d else None, coef_out=coef[k])
                for k in range(nsets):
                    fwd(k)
                hw, HW = h * w, H * W
                dsz = HW // (d * d) if d else 0
                fb = es * B * (C * hw + 3 * hw + HW + dsz) + 4 * 3 * C
                bb = es * B * (HW + dsz + 3 * hw + 2 * C * hw) + 2 * 4 * 3 * C
                us = timed(fwd, nsets, n)
                out["points"].append(dict(dtype=name, enc=enc, kernel="head_fwd_r%d_C%d" % (r, C), variant=ops.last_kernel(), us=round(us, 2),
                                          MB=round(fb / 1e6, 1), GBps=round(fb / us / 1e3, 1), frac=round(fb / us / 1e3 / peak, 3)))

                def bwd(k):
                    ops.reduce_lpg_backward(feats[k], kern, coef[k], g_full[k], g_ds[k] if d else None, r, d, g_kernel_out=gk)
                # reduce_lpg_backward allocates g_feat each call; use a static buffer variant for graph capture
                gfeat = [torch.empty_like(feats[0]) for _ in range(nsets)]
                from bts_fully_tf_b200 import _cabi
                import ctypes
                lib = _cabi.load()
                nbytes = lib.btslpg_reduce_backward_workspace_bytes(B * hw, C)
                ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)

                def bwd_static(k):
                    refs = [_cabi.as_ref(t) for t in (feats[k], kern, coef[k], g_full[k], g_ds[k] if d else None, gfeat[k], gk)]
                    _cabi.check(lib.btslpg_reduce_backward(refs[0].ptr, refs[1].ptr, refs[2].ptr, refs[3].ptr, _cabi.ptr_or_null(refs[4]), r, d,
                                                           refs[5].ptr, refs[6].ptr, None, ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                                                           _cabi.current_stream_ptr(dev)))
                us = timed(bwd_static, nsets, n)
                out["points"].append(dict(dtype=name, enc=enc, kernel="head_bwd_r%d_C%d" % (r, C), variant=ops.last_kernel(), us=round(us, 2),
                                          MB=round(bb / 1e6, 1), GBps=round(bb / us / 1e3, 1), frac=round(bb / us / 1e3 / peak, 3)))
                del feats, g_full, g_ds, coef, full, ds, gfeat
                torch.cuda.empty_cache()
    return out


def main():
    # --tune KEY=VALUE (btslpg_set_tuning, e.g. 7=1: register-staged loads instead of the TMA ring), --only-r R
    only_r, batch = None, 32
    for i, arg in enumerate(sys.argv[1:]):
        if arg == "--batch":
            batch = int(sys.argv[i + 2])
        if arg == "--tune":
            k, v = sys.argv[i + 2].split("=")
            ops.set_tuning(int(k), int(v))
        if arg == "--only-r":
            only_r = int(sys.argv[i + 2])
    dtypes = ("f32",) if "--f32-only" in sys.argv else ("f32", "bf16")
    print(json.dumps(collect(dtypes=dtypes, only_r=only_r, B=batch)))


if __name__ == "__main__":
    main()
