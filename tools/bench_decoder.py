#!/usr/bin/env python
"""Decoder-level context numbers (BASELINE.json configs 3-5): BTS decoder images/s with the fused LPG
heads, data-parallel over the GPUs of one node.  Encoder taps are synthetic tensors of the
reference's shapes (bts.py:72,80) -- the encoder is out of scope.

    python tools/bench_decoder.py [--config 3|4|5] [--steps K] [--warmup W] [--per-gpu-batch b]
    torchrun --nproc-per-node N tools/bench_decoder.py ...

Prints one JSON line per config on rank 0.  images/s is conv-bound (cuDNN): the LPG
kernels' share of the step is reported next to it, measured with CUDA events around the fused ops.
`--lpg literal` swaps the fused kernels for the op-by-op torch restatement of the reference layer
running on the GPU (what a framework port without custom kernels executes), for comparison.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bts_fully_tf_b200 import parallel  # noqa: E402
from bts_fully_tf_b200.decoder import BtsDecoder, si_log_loss  # noqa: E402

TAPS = {  # channels of [dense_features(/32), skip_2, skip_4, skip_8, skip_16], decoder filters
    "densenet161_bts": ([2208, 96, 96, 192, 384], 512),
    "resnet50": ([2048, 64, 64, 256, 512], 256),
}
CONFIGS = {
    3: dict(name="BTS-NYU DenseNet-161 decoder inference 480x640, global batch 64", encoder="densenet161_bts", H=480, W=640,
            global_batch=64, train=False, max_depth=10.0, dataset="nyu"),
    4: dict(name="BTS-KITTI Eigen 352x1216 decoder training step, global batch 32", encoder="densenet161_bts", H=352, W=1216,
            global_batch=32, train=True, max_depth=80.0, dataset="kitti"),
    5: dict(name="BTS ResNet-50 taps NYU 416x544 decoder training step, global batch 32", encoder="resnet50", H=416, W=544,
            global_batch=32, train=True, max_depth=10.0, dataset="nyu"),
}


class LiteralHeads:
    """Swap the fused heads of a BtsDecoder for the op-by-op restatement (torch ops on the GPU)."""

    @staticmethod
    def install(dec):
        from oracle import lpg_literal  # checker code, used here only as the comparison arm
        for head in (dec.reduction_8x8, dec.reduction_4x4, dec.reduction_2x2):
            layer = lpg_literal.LocalPlanarGuidanceLiteral(head.upratio)

            def fwd(feat, head=head, layer=layer):
                coef = torch.sigmoid(feat @ head.kernel.reshape(-1, 3))
                if layer.pixel_dir_unit is None:
                    layer.build(tuple(coef.shape))
                    layer.pixel_dir_unit = layer.pixel_dir_unit.to(coef.device)
                depth = layer(coef)
                if head.ds_stride:
                    return coef, depth, depth[:, ::head.ds_stride, ::head.ds_stride]
                return coef, depth
            head.forward = fwd


def run(cfg_id, a, rank, local_rank, world):
    cfg = CONFIGS[cfg_id]
    device = torch.device("cuda", local_rank)
    chans, F = TAPS[cfg["encoder"]]
    gb = cfg["global_batch"] if a.per_gpu_batch is None else a.per_gpu_batch * world
    lo, hi = parallel.shard_range(gb, world, rank)
    b = hi - lo
    H, W = cfg["H"], cfg["W"]
    torch.manual_seed(rank)
    feats = [torch.relu(torch.randn(b, H // s, W // s, c, device=device)) for s, c in zip((32, 2, 4, 8, 16), chans)]
    dec = BtsDecoder(chans, cfg["max_depth"], num_filters=F).to(device)
    if a.lpg == "literal":
        LiteralHeads.install(dec)
    dec.train(cfg["train"])
    gt = torch.rand(b, H, W, 1, device=device) * cfg["max_depth"]
    params = list(dec.parameters())
    bucket = opt = None
    if cfg["train"]:
        bucket = parallel.GradientBucket(params)
        if a.lpg == "fused":
            bucket.bind_heads(dec)
        opt = torch.optim.Adam(params, lr=parallel.scaled_learning_rate(1e-4, world), eps=1e-3)   # bts_train.py:86,125

    def step():
        if cfg["train"]:
            bucket.zero()
            if a.lpg == "fused":
                _, loss = dec.forward_loss(feats, gt, cfg["dataset"])      # fused sigmoid*max_depth + si_log_loss kernels
            else:
                loss = si_log_loss(gt, dec(feats), cfg["dataset"])
            loss.backward()
            bucket.all_reduce(average=True)            # the one exchange step: decoder gradients only
            bucket.wait()
            opt.step()
            return loss
        with torch.no_grad():
            return dec(feats)

    graph = None
    for _ in range(a.warmup):
        out = step()
    if not cfg["train"] and not a.no_graph:
        # inference: the whole decoder step as ONE CUDA graph (≈130 launches, launch-bound at small per-GPU batches);
        # every op of this repo is capture-safe (no synchronisation, no host-side data dependence)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    out = step()
            torch.cuda.current_stream().wait_stream(side)
            inner = step

            def step():                     # noqa: F811
                graph.replay()
                return out
            step()
        except Exception as exc:  # noqa: BLE001
            graph = None
            step = inner if "inner" in dir() else step
            print("CUDA graph capture failed, running eagerly: %s" % exc, file=sys.stderr)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    result = float(out.detach().float().mean())   # device->host read of the step's result

    # share of the step spent in the LPG heads (events around the three fused ops, forward only)
    lpg_ms = None
    if a.lpg == "fused":
        with torch.no_grad():
            red_in = [torch.relu(torch.randn(b, H // r, W // r, c, device=device)) for r, c in ((8, F // 4), (4, F // 4), (2, F // 8))]
            heads = (dec.reduction_8x8, dec.reduction_4x4, dec.reduction_2x2)
            for h_, x in zip(heads, red_in):
                h_(x)
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(10):
                for h_, x in zip(heads, red_in):
                    h_(x)
            s1.record()
            torch.cuda.synchronize()
            lpg_ms = s0.elapsed_time(s1) / 10
    if rank == 0:
        print(json.dumps({
            "config": cfg_id, "workload": cfg["name"], "metric": "decoder_images_per_s", "value": round(gb * a.steps / (ms * 1e-3), 2),
            "unit": "images/s", "n_gpus": world, "per_gpu_batch": b, "global_batch": gb, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": round(ms / a.steps, 3), "scaling": "strong" if a.per_gpu_batch is None else "weak",
            "lpg_path": a.lpg, "fused_heads_forward_ms": None if lpg_ms is None else round(lpg_ms, 3),
            "conv_math": "TF32 (cuDNN default)" if torch.backends.cudnn.allow_tf32 else "fp32", "cudnn_autotune": bool(torch.backends.cudnn.benchmark), "cuda_graph": graph is not None,
            "grad_bucket_bytes": None if bucket is None else bucket.nbytes(), "result_mean": result,
            "data": "synthetic encoder taps, random-init decoder"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, nargs="*", default=[3, 4, 5])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--per-gpu-batch", type=int, default=None, help="fix the per-GPU batch (weak scaling) instead of the global batch")
    ap.add_argument("--lpg", default="fused", choices=["fused", "literal"])
    ap.add_argument("--no-graph", action="store_true", help="inference configs: run eagerly instead of replaying one CUDA graph per step")
    ap.add_argument("--no-cudnn-autotune", action="store_true",
                    help="keep cuDNN's heuristic algorithm choice (default: autotune, as TensorFlow does with TF_CUDNN_USE_AUTOTUNE=1)")
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = not a.no_cudnn_autotune
    rank, local_rank, world = parallel.init_distributed()
    torch.cuda.set_device(local_rank)
    for c in a.config:
        run(c, a, rank, local_rank, world)
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    t0 = time.time()
    main()
