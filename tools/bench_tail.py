#!/usr/bin/env python
"""Micro-benchmark of the decoder-tail kernels against the HBM roofline.

    python tools/bench_tail.py [--batch 32 --height 480 --width 640] [--dtype f32|bf16] [--steps 50] [--warmup 5]

One JSON line: for each kernel the algorithmic bytes (silog forward 3 maps, silog backward 3 maps,
metrics 2 maps -- DESIGN.md section 4), the device time (CUDA events on the launch stream, 4 rotating
buffer sets so that every launch streams from HBM, not from the 126 MB L2), GB/s and the fraction of
the measured copy bandwidth; next to it the same step as the op-by-op torch restatement of the
reference ON THE GPU (what a framework port without custom kernels runs) and on the host cores.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bts_fully_tf_b200 import ops  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def literal_silog(logit, y_true, max_depth, th):
    """bts_decoder.py:102-103 + bts.py:31-38 op by op (torch), with autograd."""
    z = logit.detach().requires_grad_(True)
    y_pred = torch.sigmoid(z) * max_depth
    mask = y_true > th
    d = torch.log(y_true[mask] + 1e-7) - torch.log(y_pred[mask] + 1e-7)
    loss = torch.sqrt((d * d).mean() - 0.85 * d.mean() ** 2) * 10.0
    loss.backward()
    return loss, z.grad


def literal_metrics(y_true, y_pred, lo, hi):
    """custom_eval_metrics.py:44-88 op by op: pre_eval is re-run for each of the nine metrics, as in the reference."""
    def pre():
        mask = (y_true < hi) & (y_true > lo)
        p = torch.where(torch.isfinite(y_pred), y_pred, torch.full_like(y_pred, hi)).clamp(lo, hi)
        return y_true[mask], p[mask]
    out = []
    gt, pr = pre(); d = torch.log(gt) - torch.log(pr); out.append(torch.sqrt((d * d).mean() - d.mean() ** 2) * 100)   # noqa: E702
    gt, pr = pre(); out.append(((gt - pr).abs() / gt).mean())                                                         # noqa: E702
    gt, pr = pre(); out.append((torch.log(gt) - torch.log(pr)).abs().mean() / 2.302585092994046)                      # noqa: E702
    gt, pr = pre(); out.append(torch.sqrt(((gt - pr) ** 2).mean()))                                                   # noqa: E702
    gt, pr = pre(); out.append((((gt - pr) ** 2) / gt).mean())                                                        # noqa: E702
    gt, pr = pre(); d = torch.log(gt) - torch.log(pr); out.append(torch.sqrt((d * d).mean()))                         # noqa: E702
    for k in (1, 2, 3):
        gt, pr = pre()
        out.append((torch.maximum(gt / pr, pr / gt) < 1.25 ** k).float().mean())
    return torch.stack(out)


def time_gpu(fn, nsets, steps, warmup, graph=True):
    """us per call of fn(set index).  graph=True: the calls are captured into ONE CUDA graph that walks the
    buffer sets (launch overhead of the Python/ctypes wrapper stays outside the timed region, as in bench.py)."""
    for i in range(max(warmup, nsets)):
        fn(i % nsets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not graph:
        e0.record()
        for i in range(steps):
            fn(i % nsets)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps * 1e3
    per_graph = nsets * 4
    gr = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr, stream=side):
            for i in range(per_graph):
                fn(i % nsets)
    torch.cuda.current_stream().wait_stream(side)
    reps = max(1, steps // per_graph)
    gr.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * per_graph) * 1e3


def collect(argv=None):
    """Run the tail micro-benchmarks; returns the result dict (bench.py's extras.tail imports this)."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--steps", type=int, default=160)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--sets", type=int, default=4)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="plain launches (for ncu)")
    ap.add_argument("--skip-literal", action="store_true")
    ap.add_argument("--skip-concat", action="store_true")
    ap.add_argument("--skip-depthconv", action="store_true")
    ap.add_argument("--only-depthconv", action="store_true")
    a = ap.parse_args(argv)
    dev = torch.device("cuda", torch.cuda.current_device())
    dtype = torch.float32 if a.dtype == "f32" else torch.bfloat16
    es = 4 if a.dtype == "f32" else 2
    md, th, lo, hi = 10.0, 0.1, 1e-3, 10.0
    shape = (a.batch, a.height, a.width, 1)
    n = a.batch * a.height * a.width
    g = torch.Generator(device=dev).manual_seed(0)
    sets = []
    for _ in range(a.sets):
        logit = torch.randn(shape, generator=g, device=dev).to(dtype)
        y_true = (torch.rand(shape, generator=g, device=dev) * md).to(dtype)
        y_true[torch.rand(shape, generator=g, device=dev) < 0.3] = 0
        sets.append(dict(logit=logit, y_true=y_true, depth=torch.empty_like(logit), g=torch.empty_like(logit), ws=ops.tail_workspace(dev),
                         mws=ops.tail_workspace(dev), mout=torch.empty(10, device=dev)))
    pk, pk_src = peak()
    res = {}

    def fwd(i):
        s = sets[i]
        ops.silog_forward(s["logit"], s["y_true"], md, th, depth_est=s["depth"], workspace=s["ws"])

    def bwd(i):
        s = sets[i]
        ops.silog_backward(s["depth"], s["y_true"], md, th, s["ws"], None, True, g_out=s["g"])

    def met(i):
        s = sets[i]
        ops.eval_metrics(s["y_true"], s["depth"], lo, hi, out=s["mout"], workspace=s["mws"])

    for i in range(a.sets):
        fwd(i)
    ops.reset_launch_count()
    if a.only_depthconv:
        a.skip_concat = a.skip_cpu = True
    for name, fn, maps in (() if a.only_depthconv else (("silog_fwd", fwd, 3), ("silog_bwd", bwd, 3), ("eval_metrics", met, 2))):
        us = time_gpu(fn, a.sets, a.steps, a.warmup, graph=not a.no_graph)
        nbytes = maps * n * es
        res[name] = {"us": round(us, 2), "algorithmic_bytes": nbytes, "GBps": round(nbytes / us * 1e-3, 1), "frac_of_peak": round(nbytes / us * 1e-3 / pk, 4),
                     "kernel": ops.last_kernel()}
    # fused ELU + concat1 (bts_decoder.py:98-99): [upconv1 (32 ch), d2, d4, d8] -> 35 channels
    if not a.skip_concat:
        ca, npl, nset_c = 32, 3, 2
        csets = []
        for _ in range(nset_c):
            csets.append(dict(a=torch.randn(a.batch, a.height, a.width, ca, generator=g, device=dev).to(dtype),
                              planes=[torch.randn(shape, generator=g, device=dev).to(dtype) for _ in range(npl)],
                              out=torch.empty(a.batch, a.height, a.width, ca + npl, device=dev, dtype=dtype)))
        g_cat = torch.randn(a.batch, a.height, a.width, ca + npl, generator=g, device=dev).to(dtype)

        def cfwd(i):
            s = csets[i]
            ops.concat_forward(s["a"], s["planes"], act=True, out=s["out"])
        us = time_gpu(cfwd, nset_c, max(8, a.steps // 4), 2, graph=not a.no_graph)
        nbytes = (ca + npl + ca + npl) * n * es
        res["concat1_fwd"] = {"us": round(us, 2), "algorithmic_bytes": nbytes, "GBps": round(nbytes / us * 1e-3, 1), "frac_of_peak": round(nbytes / us * 1e-3 / pk, 4),
                              "kernel": ops.last_kernel()}
        # the backward allocates its outputs: time it with plain launches (1.4 ms kernels: launch overhead is < 1 %)
        us = time_gpu(lambda i: ops.concat_backward(g_cat, csets[i]["out"], True, ca, 0, npl), nset_c, max(8, a.steps // 8), 2, graph=False)
        nbytes = (2 * (ca + npl) + ca + npl) * n * es
        res["concat1_bwd"] = {"us": round(us, 2), "algorithmic_bytes": nbytes, "GBps": round(nbytes / us * 1e-3, 1), "frac_of_peak": round(nbytes / us * 1e-3 / pk, 4),
                              "kernel": ops.last_kernel()}
        if a.dtype == "f32" and not a.skip_literal:
            def lit(i):
                x = csets[i]["a"].detach().requires_grad_(True)
                y = torch.cat([torch.nn.functional.elu(x)] + csets[i]["planes"], 3)
                y.backward(g_cat)
            us_l = time_gpu(lit, nset_c, 6, 2, graph=False)
            res["concat1_torch_gpu_literal_fwd_bwd_us"] = round(us_l, 1)
            res["concat1_speedup"] = round(us_l / (res["concat1_fwd"]["us"] + res["concat1_bwd"]["us"]), 2)
        del csets, g_cat
        torch.cuda.empty_cache()
    # last convolution forward with iconv1's ELU and sigmoid * max_depth folded in (bts_decoder.py:100-103), C = F/16
    if not a.skip_depthconv:
        for cdc in (32, 16):
            dsets = [dict(x=torch.randn(a.batch, a.height, a.width, cdc, generator=g, device=dev).to(dtype),
                          y=torch.empty(shape, device=dev, dtype=dtype)) for _ in range(2 if cdc == 32 else 4)]
            w9c = (torch.rand(9 * cdc, generator=g, device=dev) - 0.5) * 0.28
            us = time_gpu(lambda i: ops.depthconv_forward(dsets[i]["x"], w9c, act_in=True, sigmoid_scale=md, out=dsets[i]["y"]), len(dsets),
                          max(8, a.steps // 4), 2, graph=not a.no_graph)
            nbytes = (cdc + 1) * n * es
            key = "depthconv_fwd_C%d" % cdc
            res[key] = {"us": round(us, 2), "algorithmic_bytes": nbytes, "GBps": round(nbytes / us * 1e-3, 1), "frac_of_peak": round(nbytes / us * 1e-3 / pk, 4),
                        "kernel": ops.last_kernel()}
            ops.set_tuning(9, 1)
            us1 = time_gpu(lambda i: ops.depthconv_forward(dsets[i]["x"], w9c, act_in=True, sigmoid_scale=md, out=dsets[i]["y"]), len(dsets),
                           max(8, a.steps // 4), 2, graph=not a.no_graph)
            res[key]["fp32pipe_us"] = round(us1, 2)
            ops.set_tuning(9, 0)
            if a.dtype == "f32" and not a.skip_literal:
                wt = w9c.view(3, 3, cdc, 1).permute(3, 2, 0, 1).contiguous()
                xs = [d["x"].permute(0, 3, 1, 2) for d in dsets]              # NCHW views of channels_last memory, as in the decoder

                def lib_path(i):
                    with torch.no_grad():
                        return torch.sigmoid(torch.nn.functional.conv2d(torch.nn.functional.elu(xs[i]), wt, padding=1)) * md
                torch.backends.cudnn.benchmark = True
                us_l = time_gpu(lib_path, len(dsets), 8, 3, graph=False)
                res[key]["library_path_us"] = round(us_l, 1)
                res[key]["speedup"] = round(us_l / us, 2)
                ref = lib_path(0).permute(0, 2, 3, 1)
                ops.depthconv_forward(dsets[0]["x"], w9c, act_in=True, sigmoid_scale=md, out=dsets[0]["y"])
                res[key]["max_abs_diff_vs_library_tf32"] = float((dsets[0]["y"] - ref).abs().max())
            # backward of the same layer: d loss / d x (with ELU') and d loss / d kernel from one pass (bts_decoder.py:102)
            gouts = [torch.randn(shape, generator=g, device=dev).to(dtype) for _ in dsets]
            us_b = time_gpu(lambda i: ops.depthconv_backward(dsets[i]["x"], w9c, gouts[i], act_in=True), len(dsets), max(8, a.steps // 8), 2, graph=False)
            nb = (2 * cdc + 1) * n * es
            res["depthconv_bwd_C%d" % cdc] = {"us": round(us_b, 2), "algorithmic_bytes": nb, "GBps": round(nb / us_b * 1e-3, 1),
                                              "frac_of_peak": round(nb / us_b * 1e-3 / pk, 4), "kernel": ops.last_kernel()}
            del dsets, gouts
            torch.cuda.empty_cache()
    launches = ops.launch_count()

    # the same work as torch ops on the GPU (a port without custom kernels)
    if a.dtype == "f32" and not a.skip_literal and not a.only_depthconv:
        lit_f = time_gpu(lambda i: literal_silog(sets[i]["logit"], sets[i]["y_true"], md, th), a.sets, max(3, a.steps // 5), 2, graph=False)
        lit_m = time_gpu(lambda i: literal_metrics(sets[i]["y_true"], sets[i]["depth"], lo, hi), a.sets, max(3, a.steps // 5), 2, graph=False)
        res["torch_gpu_literal"] = {"silog_fwd_bwd_us": round(lit_f, 1), "eval_metrics_us": round(lit_m, 1),
                                    "speedup_silog": round(lit_f / (res["silog_fwd"]["us"] + res["silog_bwd"]["us"]), 1),
                                    "speedup_metrics": round(lit_m / res["eval_metrics"]["us"], 1)}
        # parity of the timed configuration against the literal run (same device, float32)
        s = sets[0]
        loss_l, gz_l = literal_silog(s["logit"], s["y_true"], md, th)
        _, loss_k, ws = ops.silog_forward(s["logit"], s["y_true"], md, th, depth_est=s["depth"])
        gz_k = ops.silog_backward(s["depth"], s["y_true"], md, th, ws, None, True)
        res["check"] = {"loss_rel_diff": abs(float(loss_k) - float(loss_l.detach())) / float(loss_l.detach()),
                        "grad_max_rel_diff": float((gz_k - gz_l).abs().max() / gz_l.abs().max())}
    if not a.skip_cpu and a.dtype == "f32":
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sb = min(a.batch, 8)
        lz, yt = sets[0]["logit"][:sb].cpu(), sets[0]["y_true"][:sb].cpu()
        literal_silog(lz, yt, md, th)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            literal_silog(lz, yt, md, th)
        dt = (time.perf_counter() - t0) / reps
        res["cpu_baseline"] = {"kind": "port", "cores": cores, "sample": "%d of %d images, silog fwd+bwd, torch-CPU literal" % (sb, a.batch),
                               "GBps": round(6 * sb * a.height * a.width * 4 / dt * 1e-9, 2)}
    return {"bench": "decoder tail (SURVEY 8(f) N2, N4)", "workload": "batch %d at %dx%d, %s" % (a.batch, a.height, a.width, a.dtype),
            "peak_GBps": pk, "peak_source": pk_src, "sets": a.sets, "steps": a.steps, "cuda_graph": not a.no_graph, "gpu_launches": launches, **res}


def main():
    print(json.dumps(collect()))


if __name__ == "__main__":
    main()
