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


def _max_over_ranks(ms, device, world):
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _timed(step, steps, device, world):
    """barrier + sync, `steps` steps between two CUDA events on the current stream, barrier + sync; max over ranks (ms)."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    return _max_over_ranks(e0.elapsed_time(e1), device, world), out


def run_config(cfg_id, rank, local_rank, world, steps=5, warmup=3, per_gpu_batch=None, lpg="fused", no_graph=False, measure_comm=True):
    """One BASELINE config (3: inference, 4 / 5: data-parallel training step) -> result dict (identical on every rank)."""
    from bts_fully_tf_b200 import ops, trainer
    cfg = CONFIGS[cfg_id]
    device = torch.device("cuda", local_rank)
    chans, F = TAPS[cfg["encoder"]]
    gb = cfg["global_batch"] if per_gpu_batch is None else per_gpu_batch * world
    lo, hi = parallel.shard_range(gb, world, rank)
    b = hi - lo
    H, W = cfg["H"], cfg["W"]
    torch.manual_seed(0)                                   # identical initial weights on every rank (a replicated model)
    dec = BtsDecoder(chans, cfg["max_depth"], num_filters=F).to(device)
    torch.manual_seed(1000 + rank)                         # a different shard of synthetic data per rank
    feats = [torch.relu(torch.randn(b, H // s, W // s, c, device=device)) for s, c in zip((32, 2, 4, 8, 16), chans)]
    if lpg == "literal":
        LiteralHeads.install(dec)
    dec.train(cfg["train"])
    gt = torch.rand(b, H, W, 1, device=device) * cfg["max_depth"]
    res = {"config": cfg_id, "workload": cfg["name"], "metric": "decoder_images_per_s", "unit": "images/s", "n_gpus": world, "per_gpu_batch": b,
           "global_batch": gb, "steps": steps, "warmup": warmup, "scaling": "strong" if per_gpu_batch is None else "weak", "lpg_path": lpg,
           "conv_math": "TF32 (cuDNN default)" if torch.backends.cudnn.allow_tf32 else "fp32", "cudnn_autotune": bool(torch.backends.cudnn.benchmark),
           "data": "synthetic encoder taps, random-init decoder"}

    if cfg["train"] and lpg == "fused":
        # the data-parallel training step: forward, fused loss, backward, chunked all-reduce overlapped with backward on a side
        # stream, fused AdamW -- ONE CUDA graph per rank (trainer.DataParallelStep; bts_train.py:194-209, :125-131)
        eng = trainer.DataParallelStep(dec, feats, gt, dataset=cfg["dataset"], base_lr=1e-4, total_steps=100000, adam_eps=1e-3, use_graph=not no_graph)
        ops.reset_launch_count()
        eng._eager_step()                                                 # first eager step (autotuning) also counts this repo's launches
        own_launches = ops.launch_count()
        graph_err = None
        try:
            eng.warmup_and_capture(max(1, warmup - 1))
        except Exception as exc:  # noqa: BLE001
            graph_err = "%s: %s" % (type(exc).__name__, str(exc)[:160])
            eng.graph = None
        if world > 1:                                                     # all ranks replay graphs, or none does
            ok = torch.tensor([1 if eng.graph is not None else 0], device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0 and eng.graph is not None:
                eng.graph, graph_err = None, "capture failed on another rank"
        ms, loss = _timed(eng.step, steps, device, world)
        res.update({"value": round(gb * steps / (ms * 1e-3), 2), "ms_per_step": round(ms / steps, 3), "cuda_graph": eng.graph is not None,
                    "graph_error": graph_err, "launches_per_step": 1 if eng.graph is not None else None, "own_kernel_launches_per_step": own_launches,
                    "grad_bucket_bytes": eng.flat.nbytes(), "grad_chunks": [int(hi_ - lo_) * 4 for lo_, hi_, _, _ in eng.flat.chunks],
                    "allreduce_order": list(eng.comm.launch_order), "nccl_registered_bucket": bool(eng.registered),
                    "optimizer": "fused AdamW kernel (1/N, decay, Adam, grad zeroing in one pass; device-resident step / lr)",
                    "result_mean": float(loss), "completed_updates": eng.completed_updates()})
        if world > 1 and measure_comm:
            # (a) the whole bucket's all-reduce alone, (b) the same step with the exchange removed: step - (b) = exposed communication
            g = eng.flat.grad
            for _ in range(3):
                dist.all_reduce(g)
            ms_ar, _ = _timed(lambda: dist.all_reduce(g), 10, device, world)
            res["allreduce_ms"] = round(ms_ar / 10, 4)
            try:
                eng.comm.enabled = False
                if eng.graph is not None:
                    eng.capture()
                eng.step()
                ms_nc, _ = _timed(eng.step, steps, device, world)
                res["ms_per_step_without_allreduce"] = round(ms_nc / steps, 3)
                res["exposed_comm_ms"] = round(max(0.0, (ms - ms_nc) / steps), 4)
            except Exception as exc:  # noqa: BLE001
                res["exposed_comm_error"] = "%s: %s" % (type(exc).__name__, str(exc)[:160])
        eng.close()
        del eng
        return res

    bucket = opt = None
    if cfg["train"]:                                                      # literal-LPG comparison arm: plain eager step
        params = list(dec.parameters())
        bucket = parallel.GradientBucket(params)
        opt = torch.optim.Adam(params, lr=parallel.scaled_learning_rate(1e-4, world), eps=1e-3)

    def step():
        if cfg["train"]:
            bucket.zero()
            loss = si_log_loss(gt, dec(feats), cfg["dataset"])
            loss.backward()
            bucket.all_reduce(average=True)
            bucket.wait()
            opt.step()
            return loss
        with torch.no_grad():
            return dec(feats)

    graph = None
    for _ in range(warmup):
        out = step()
    ops.reset_launch_count()
    out = step()
    own_launches = ops.launch_count()
    if not cfg["train"] and not no_graph:
        # inference: the whole decoder step as ONE CUDA graph (launch-bound at small per-GPU batches);
        # every op of this repo is capture-safe (no synchronisation, no host-side data dependence)
        eager = step
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    out = eager()
            torch.cuda.current_stream().wait_stream(side)

            def step():                     # noqa: F811
                graph.replay()
                return out
            step()
        except Exception as exc:  # noqa: BLE001
            graph, step = None, eager
            print("CUDA graph capture failed, running eagerly: %s" % exc, file=sys.stderr)
    ms, out = _timed(step, steps, device, world)
    res.update({"value": round(gb * steps / (ms * 1e-3), 2), "ms_per_step": round(ms / steps, 3), "cuda_graph": graph is not None,
                "own_kernel_launches_per_step": own_launches, "result_mean": float(out.detach().float().mean())})
    return res


def run(cfg_id, a, rank, local_rank, world):
    res = run_config(cfg_id, rank, local_rank, world, a.steps, a.warmup, a.per_gpu_batch, a.lpg, a.no_graph)
    if rank == 0:
        print(json.dumps(res), flush=True)


def main():
    # `@file` arguments as in the reference's scripts (bts_train.py:45-53: one or more whitespace-separated options per line)
    ap = argparse.ArgumentParser(fromfile_prefix_chars="@")
    ap.convert_arg_line_to_args = lambda line: [a for a in line.split() if a.strip()]
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
