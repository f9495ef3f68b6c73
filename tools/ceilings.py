#!/usr/bin/env python
"""HBM ceilings for different read/write mixes on this GPU, same timing method as bench.py (CUDA events,
multi-launch CUDA graph, buffers that exceed the 126 MB L2):  write-only (cudaMemsetAsync / fill), read-only
(a sum reduction), copy (read + write).  Context for roofline fractions of kernels whose traffic is not 50/50:
lpg_fwd_multi is 77 % writes, lpg_bwd_multi 97 % reads.  One JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def timed_graph(fn, per_graph=8, reps=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(0)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for k in range(per_graph):
                fn(k)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * per_graph) * 1e-3


def main():
    dev = torch.device("cuda:0")
    nbytes = 512 << 20
    nset = 4
    src = [torch.empty(nbytes // 4, dtype=torch.float32, device=dev).normal_() for _ in range(nset)]
    dst = [torch.empty(nbytes // 4, dtype=torch.float32, device=dev) for _ in range(nset)]
    acc = torch.zeros(nset, device=dev)
    out = {}
    t = timed_graph(lambda k: dst[k % nset].zero_())
    out["write_only_memset_GBps"] = round(nbytes / t * 1e-9, 1)
    t = timed_graph(lambda k: dst[k % nset].fill_(1.5))
    out["write_only_fill_GBps"] = round(nbytes / t * 1e-9, 1)
    t = timed_graph(lambda k: torch.sum(src[k % nset], dim=0, out=acc[k % nset]))
    out["read_only_sum_GBps"] = round(nbytes / t * 1e-9, 1)
    t = timed_graph(lambda k: dst[k % nset].copy_(src[k % nset]))
    out["copy_GBps_read_plus_write"] = round(2 * nbytes / t * 1e-9, 1)
    try:
        out["MEASURED_PEAKS_hbm_gbs"] = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        out["MEASURED_PEAKS_hbm_gbs"] = None
    out["buffer_MB"] = nbytes >> 20
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
