#!/usr/bin/env python
"""Weight gradient of the full-resolution 3x3 convolutions (upconv1, iconv1; bts_decoder.py:98, :100): the tcgen05 kernel
(ops.conv3x3_wgrad) against the library's (aten convolution_backward, weight gradient only, TF32, channels_last) at the decoder's
shapes.  Algorithmic bytes per pixel: (Cin + Cout) * 4.  One JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_iconv import timed  # noqa: E402
from bts_fully_tf_b200 import ops  # noqa: E402

CASES = [  # name, B, H, W, Cin, Cout
    ("cfg5 upconv1 (resnet50 NYU 416x544)", 32, 416, 544, 32, 16),
    ("cfg5 iconv1", 32, 416, 544, 20, 16),
    ("cfg2/3 upconv1 (densenet161 NYU 480x640)", 32, 480, 640, 64, 32),
    ("cfg2/3 iconv1", 32, 480, 640, 36, 32),
    ("cfg4 upconv1 (densenet161 KITTI 352x1216)", 32, 352, 1216, 64, 32),
    ("cfg4 iconv1", 32, 352, 1216, 36, 32),
    ("cfg5 at per-GPU batch 4, upconv1", 4, 416, 544, 32, 16),
    ("cfg5 upconv1 as sub-pixel conv on H/2 (32 -> 4x16)", 32, 208, 272, 32, 64),
    ("cfg4 upconv1 as sub-pixel conv on H/2 (64 -> 4x32)", 32, 176, 608, 64, 128),
    ("cfg5 conv_block 2 upconv on H/2 (64 -> 32)", 32, 208, 272, 64, 32),
    ("cfg5 conv_block 2 iconv on H/2 (100 -> 32)", 32, 208, 272, 100, 32),
    ("cfg5 conv_block 3 upconv on H/4 (64 -> 64)", 32, 104, 136, 64, 64),
    ("cfg4 conv_block 2 upconv on H/2 (128 -> 64)", 32, 176, 608, 128, 64),
    ("cfg4 conv_block 2 iconv on H/2 (164 -> 64)", 32, 176, 608, 164, 64),
    ("cfg4 conv_block 3 upconv on H/4 (128 -> 128)", 32, 88, 304, 128, 128),
]


def collect(device=None, library=True):
    dev = device or torch.device("cuda", torch.cuda.current_device())
    peak = 6533.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    out = {"workload": "3x3 conv weight gradient, float32 tensors, TF32 tensor-core arithmetic", "peak_GBps": peak, "points": []}
    gen = torch.Generator(device=dev).manual_seed(0)
    for name, B, H, W, Cin, Cout in CASES:
        nsets = 2
        xs = [torch.randn(B, H, W, Cin, device=dev, generator=gen) for _ in range(nsets)]
        gs = [torch.randn(B, H, W, Cout, device=dev, generator=gen) for _ in range(nsets)]
        w = torch.randn(Cout, Cin, 3, 3, device=dev).contiguous(memory_format=torch.channels_last)
        res = torch.empty(3, 3, Cin, Cout, device=dev)
        nbytes = B * H * W * (Cin + Cout) * 4
        ours = timed(lambda k: ops.conv3x3_wgrad(xs[k], gs[k], out=res), nsets)
        pt = {"case": name, "B": B, "H": H, "W": W, "Cin": Cin, "Cout": Cout, "algorithmic_bytes": nbytes, "tcgen05_us": round(ours, 1),
              "GBps": round(nbytes / ours / 1e3, 1), "frac_of_peak": round(nbytes / ours / 1e3 / peak, 4),
              "tf32_tflops": round(2 * B * H * W * 9 * Cin * Cout / ours / 1e6, 1)}
        if library:
            def lib(k):
                torch.ops.aten.convolution_backward(gs[k].permute(0, 3, 1, 2), xs[k].permute(0, 3, 1, 2), w, None, [1, 1], [1, 1], [1, 1], False,
                                                    [0, 0], 1, [False, True, False])
            t = timed(lib, nsets, reps=3)
            pt["library_us"] = round(t, 1)
            pt["speedup"] = round(t / ours, 2)
        out["points"].append(pt)
        del xs, gs
    return out


if __name__ == "__main__":
    print(json.dumps(collect(library="--no-library" not in sys.argv)))
