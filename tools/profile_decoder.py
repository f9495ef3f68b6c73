#!/usr/bin/env python
"""Where a decoder step spends its GPU time: torch.profiler over one inference step (config 3 shapes, per-GPU
batch 32) and one training step (config 4), kernels grouped by name, plus CUDA-event timings of the
full-resolution tail convolutions (upconv1 incl. its nearest x2 up-sampling, iconv1, depth conv) in isolation."""
import json
import os
import sys

import torch
import torch.nn.functional as F
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bts_fully_tf_b200 import ops  # noqa: E402
from bts_fully_tf_b200.decoder import BtsDecoder  # noqa: E402

TAPS = ([2208, 96, 96, 192, 384], 512)


def top_kernels(prof, n=24):
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
        if t and e.device_type.name == "CUDA":
            rows.append((t, e.count, e.key[:90]))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    return tot, [dict(us=round(t, 1), share=round(t / tot, 3), calls=c, name=k) for t, c, k in rows[:n]]


def event_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    chans, Fd = TAPS
    out = {}
    for tag, (B, H, W, train) in {"infer_480x640_b32": (32, 480, 640, False), "train_352x1216_b16": (16, 352, 1216, True)}.items():
        torch.manual_seed(0)
        feats = [torch.relu(torch.randn(B, H // s, W // s, c, device=dev)) for s, c in zip((32, 2, 4, 8, 16), chans)]
        dec = BtsDecoder(chans, 10.0, num_filters=Fd).to(dev)
        dec.train(train)
        gt = torch.rand(B, H, W, 1, device=dev) * 10

        def step():
            if train:
                dec.zero_grad(set_to_none=True)
                _, loss = dec.forward_loss(feats, gt, "nyu")
                loss.backward()
            else:
                with torch.no_grad():
                    dec(feats)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step()
            torch.cuda.synchronize()
        tot, rows = top_kernels(prof)
        out[tag] = {"gpu_us_total": round(tot, 1), "top": rows}
        if not train:
            # the tail in isolation (inference)
            with torch.no_grad():
                iconv2 = torch.randn(B, Fd // 8, H // 2, W // 2, device=dev).contiguous(memory_format=torch.channels_last)
                cat1 = torch.randn(B, Fd // 16 + 3, H, W, device=dev).contiguous(memory_format=torch.channels_last)
                ic1 = torch.randn(B, Fd // 16, H, W, device=dev).contiguous(memory_format=torch.channels_last)
                out[tag]["tail_ms"] = {
                    "upsample_x2_torch": round(event_time(lambda: F.interpolate(iconv2, scale_factor=2, mode="nearest")), 3),
                    "upsample_x2_ours": round(event_time(lambda: ops.upsample2x_forward(iconv2.permute(0, 2, 3, 1))), 3),
                    "upconv1_with_torch_upsample": round(event_time(lambda: dec.upconv1(F.interpolate(iconv2, scale_factor=2, mode="nearest"))), 3),
                    "iconv1_conv_elu": round(event_time(lambda: F.elu(dec.iconv1(cat1))), 3),
                    "depth_conv_sigmoid": round(event_time(lambda: torch.sigmoid(dec.depth_conv(ic1)) * 10.0), 3),
                    "whole_step": round(event_time(step), 3),
                }
        del dec, feats
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
