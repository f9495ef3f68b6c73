#!/usr/bin/env python
"""iconv1 without concat1 (tcgen05 implicit GEMM, ops.iconv1_forward) against the path it replaces (fused ELU + concat kernel
followed by the cuDNN convolution, TF32) at B = 32, 480x640 (BASELINE config 2 shapes) for NF = 32 (densenet161) and 16
(resnet50).  Algorithmic bytes per pixel: NF*4 (upconv1) + 12 (planes) read, NF*4 written.  One JSON line."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bts_fully_tf_b200 import ops  # noqa: E402


def timed(fn, nsets, reps=12):
    for k in range(nsets):
        fn(k)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for k in range(nsets * 2):
                fn(k % nsets)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * nsets * 2) * 1e3


def collect(B=32, H=480, W=640, device=None):
    dev = device or torch.device("cuda", torch.cuda.current_device())
    peak = 6533.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    torch.backends.cudnn.benchmark = True
    out = {"workload": "iconv1 forward, B=%d at %dx%d, float32 tensors, TF32 tensor-core arithmetic" % (B, H, W), "peak_GBps": peak, "points": []}
    gen = torch.Generator(device=dev).manual_seed(0)
    for NF in (32, 16):
        nsets = 2
        a4 = [torch.randn(B, H // 2, W // 2, 4 * NF, device=dev, generator=gen) for _ in range(nsets)]
        planes = [[torch.rand(B, H, W, 1, device=dev, generator=gen) * 10 for _ in range(3)] for _ in range(nsets)]
        w = torch.randn(NF, NF + 3, 3, 3, device=dev, generator=gen) * 0.05
        hwio = ops.kernel_hwio(w)
        outs = [torch.empty(B, H, W, NF, device=dev) for _ in range(nsets)]
        pad = ops.pad_to(NF + 3)
        wp = F.pad(w, (0, 0, 0, 0, 0, pad)).contiguous(memory_format=torch.channels_last)
        nbytes = B * H * W * (NF * 4 + 12 + NF * 4)

        def fused(k):
            ops.iconv1_forward(a4[k], planes[k], hwio, a_subpixel=True, out=outs[k])

        def library(k):
            cat = ops.concat_forward(a4[k], planes[k], act=True, pad=pad, a_subpixel=True)
            return F.conv2d(cat.permute(0, 3, 1, 2), wp, padding=1)
        with torch.no_grad():
            us_f = timed(fused, nsets)
            kern = ops.last_kernel()
            us_l = timed(library, nsets)
            ref = library(0).permute(0, 2, 3, 1)
            fused(0)
            torch.cuda.synchronize()
            diff = float((outs[0] - ref).abs().max() / ref.abs().max())
        out["points"].append({"NF": NF, "kernel": kern, "us": round(us_f, 1), "algorithmic_bytes": nbytes, "GBps": round(nbytes / us_f / 1e3, 1),
                              "frac_of_peak": round(nbytes / us_f / 1e3 / peak, 4), "library_path_us": round(us_l, 1),
                              "library_path": "ops.concat_forward (ELU + concat1, pad to %d ch) + cuDNN conv (TF32, autotuned, channels_last)" % (NF + 3 + pad),
                              "speedup": round(us_l / us_f, 2), "max_rel_diff_vs_library": diff})
        del a4, planes, outs
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    print(json.dumps(collect()))
