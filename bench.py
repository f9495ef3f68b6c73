#!/usr/bin/env python
"""bench.py -- LPG hot-path benchmark (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f32|bf16] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input: LocalPlanarGuidance forward
AND backward for the three decoder scales (r = 8, 4, 2, with the strided down-sampled copies of
bts_decoder.py:81,88) at batch 32, 480x640.  `value` is algorithmic GB/s (BASELINE.md section 3
bytes / device time) with inputs resident in HBM; `e2e` is the same metric with every input coming
from pinned host memory and every result returned to it inside the timed region; `roofline` is the
dominant kernel against the measured HBM copy bandwidth; `cpu_baseline` is the op-by-op restatement
of the reference layer on the host cores (TensorFlow is not installable here -- this is a port,
labelled as such).

Timing hygiene: >= 3 warm-up steps; each step works on one of several disjoint buffer sets
(each set is ~340 MB > the 126 MB L2, so nothing is re-read from cache); CUDA events on the
launching stream; max over ranks; SM clocks and throttle reasons sampled through NVML.

`--impl reference`: times the CPU restatement of the reference (oracle/lpg_literal.py, all host
threads) on the same workload (whole batch per step, the given step / warm-up counts); rank 0 only.

`extras` carries the numbers DESIGN.md / BASELINE.md quote next to the headline: the fused head kernels the decoder
actually launches (extras.heads), bf16 I/O (extras.bf16), the decoder-tail kernels (extras.tail), and BASELINE configs
3-5 (extras.decoder_config3/4/5: decoder images/s, inference and the data-parallel training step with its all-reduce).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lpg_fwd_bwd_algorithmic_GBps"
UNIT = "GB/s"


def parse_args():
    # `@file` arguments as in the reference's scripts (bts_train.py:45-53: one or more whitespace-separated options per line)
    ap = argparse.ArgumentParser(fromfile_prefix_chars="@")
    ap.convert_arg_line_to_args = lambda line: [a for a in line.split() if a.strip()]
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--skip-decoder", action="store_true", help="skip the decoder images/s context line (extras.decoder_config3)")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--sets", type=int, default=4, help="disjoint device buffer sets rotated between steps")
    ap.add_argument("--mode", default="auto", choices=["auto", "multi", "per-layer"],
                    help="one launch for all three scales (multi) or one launch per layer")
    ap.add_argument("--no-graph", action="store_true", help="launch from Python every step instead of replaying CUDA graphs")
    ap.add_argument("--graph-repeat", type=int, default=4, help="steps per CUDA graph = sets * graph_repeat")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-extras", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--ref-sample-batch", type=int, default=0, help="--impl reference: images per step (0 = the whole batch, the default)")
    ap.add_argument("--ref-max-seconds", type=float, default=240.0, help="--impl reference: wall-clock guard of the timed loop")
    ap.add_argument("--skip-train", action="store_true", help="skip the training-step context lines (extras.decoder_config4 / 5)")
    ap.add_argument("--decoder-steps", type=int, default=5)
    return ap.parse_args()


def workload_name(a):
    return "LPG layer microbench: scales 8/4/2 fwd+bwd, batch %d at %dx%d, %s" % (a.batch, a.height, a.width, a.dtype)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------
# CPU restatement of the reference (the "port"): used by cpu_baseline and --impl reference
# --------------------------------------------------------------------------------------------
def cpu_reference_pass(sample_batch, H, W, seed=0):
    """One fwd+bwd of the three LPG layers on torch-CPU, literal op order of custom_layers.py:47-56.
    Returns a closure that runs the pass once, and the algorithmic bytes it covers."""
    import torch
    from oracle import lpg_literal
    from bts_fully_tf_b200.host_io import DECODER_SCALES, algorithmic_bytes
    g = torch.Generator().manual_seed(seed)
    items = []
    for r, d in DECODER_SCALES:
        coef = torch.sigmoid(torch.randn(sample_batch, H // r, W // r, 3, generator=g))
        g_full = torch.randn(sample_batch, H, W, 1, generator=g)
        g_ds = torch.randn(sample_batch, H // d, W // d, 1, generator=g) if d else None
        layer = lpg_literal.LocalPlanarGuidanceLiteral(r)
        layer.build(tuple(coef.shape))          # the constant is built once, like Keras build()
        items.append((layer, coef, g_full, g_ds, d))

    def run():
        for layer, coef, g_full, g_ds, d in items:
            x = coef.clone().requires_grad_(True)
            out = layer(x)
            outs, grads = [out], [g_full]
            if d:
                outs.append(lpg_literal.downsample(out, d))
                grads.append(g_ds)
            torch.autograd.backward(outs, grads)
        return x.grad

    fwd, bwd, _ = algorithmic_bytes(sample_batch, H, W, 4)
    return run, fwd + bwd


def time_cpu_baseline(a, budget_s):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sb = max(1, min(a.batch, 2))
    run, nbytes = cpu_reference_pass(sb, a.height, a.width)
    run()                                        # warm-up
    t0 = time.perf_counter()
    run()
    one = time.perf_counter() - t0
    # grow the sample towards the whole batch while a pass stays under ~0.6 s, then fill the budget
    while sb < a.batch and one * 2 < 0.6:
        sb = min(a.batch, sb * 2)
        run, nbytes = cpu_reference_pass(sb, a.height, a.width)
        run()
        t0 = time.perf_counter()
        run()
        one = time.perf_counter() - t0
    reps = max(3, min(200, int(budget_s / max(one, 1e-3))))
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return {"value": round(nbytes / med / 1e9, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "restated-reference CPU (torch-CPU literal of custom_layers.py:47-56 + autograd), %d of %d images "
                      "of the same workload, fp32, median of %d passes (%.3f s each)" % (sb, a.batch, reps, med)}


def run_reference_arm(a):
    """--impl reference: rank 0 only; the CPU restatement of the reference on all host threads, on the SAME config as the
    B200 arm: every step is one forward + backward of the three LPG layers over the whole batch (a full pass takes ~0.3 s on
    32 cores), exactly --steps timed steps after --warmup warm-up steps.  --ref-sample-batch N (< batch) restricts a step to
    N images for slow hosts; a wall-clock guard (--ref-max-seconds) stops early and reports the steps actually timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sb = a.batch if a.ref_sample_batch <= 0 else max(1, min(a.batch, a.ref_sample_batch))
    run, nbytes = cpu_reference_pass(sb, a.height, a.width)
    steps, warm = max(1, a.steps), max(a.warmup, 3)          # the B200 arm clamps its warm-up the same way
    t_guard = time.perf_counter()
    for _ in range(warm):
        run()
        if time.perf_counter() - t_guard > a.ref_max_seconds / 4:
            break
    done = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
        done += 1
        if time.perf_counter() - t0 > a.ref_max_seconds:
            break
    dt = time.perf_counter() - t0
    value = nbytes * done / dt / 1e9
    sample = "%d steps x %d of %d images, fwd+bwd of the three LPG layers at %dx%d" % (done, sb, a.batch, a.height, a.width)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": a.gpus, "steps": done,
        "warmup": warm, "ms_per_step": round(dt / done * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "step_sample": "%d of %d images per step" % (sb, a.batch),
                   "note": "TensorFlow is not installable in this image: this is the op-by-op torch-CPU restatement "
                           "of the reference layer (oracle/lpg_literal.py), not TensorFlow"},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the GPU works."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.hot = index, [], False
        self.stop_flag = False
        self.ok = False
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((self.hot, mhz, reasons))
            except Exception:
                pass
            time.sleep(0.0005 if self.hot else 0.01)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "note": "NVML unavailable"}
        hot = [s for s in self.samples if s[0]]
        window = "timed regions"
        if len(hot) < 3:
            hot, window = self.samples, "whole bench (timed regions too short to sample)"
        bits = 0
        for _, _, r in hot:
            bits |= r
        return {"sm_mhz": statistics.median(s[1] for s in hot), "sm_max_mhz": self.sm_max,
                "reasons": sorted(name for bit, name in self.REASONS.items() if bits & bit),
                "samples": len(hot), "window": window}


# --------------------------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------------------------
def _tools():
    tp = os.path.join(ROOT, "tools")
    if tp not in sys.path:
        sys.path.insert(0, tp)


def decoder_context(cfg_id, rank, local_rank, world, steps=5, warmup=3):
    """BASELINE configs 3-5 (tools/bench_decoder.py): the decoder with this repo's kernels around cuDNN's convolutions, global
    batch sharded over the ranks (strong scaling), synthetic encoder taps of the reference's shapes (bts.py:72,80).  Config 3:
    batched inference as one CUDA graph.  Configs 4 / 5: the data-parallel training step of trainer.DataParallelStep -- the
    gradient all-reduce is the one collective of the design, so these are the lines that exercise NCCL."""
    import torch
    _tools()
    import bench_decoder
    torch.backends.cudnn.benchmark = True            # TensorFlow autotunes cuDNN by default as well
    return bench_decoder.run_config(cfg_id, rank, local_rank, world, steps=steps, warmup=warmup)


def heads_context(peak):
    """The fused reduction-head + LPG kernels the decoder launches (bts_decoder.py:79-94) at config-2 shapes, both encoders,
    float32: 12 points (forward and backward of r = 8 / 4 / 2), same timing hygiene as the headline."""
    _tools()
    import sweep_head
    out = sweep_head.collect(dtypes=("f32",))
    pts = out["points"]
    fr = [p["frac"] for p in pts]
    return {"points": pts, "min_frac": min(fr), "max_frac": max(fr), "peak": out["peak"],
            "note": "B=32, 480x640; densenet161 C=128/128/64, resnet50 C=64/64/32; algorithmic bytes per BASELINE.md section 3"}


def iconv_context():
    """iconv1 without concat1 (tcgen05 implicit GEMM, bts_decoder.py:98-100) against the fused concat kernel + cuDNN convolution."""
    _tools()
    import bench_iconv
    return bench_iconv.collect()


def wgrad_context():
    """Weight gradient of the full-resolution / few-channel 3x3 convolutions (bts_decoder.py:98, :100, :38-44): the tcgen05 kernel against
    the library's weight-gradient kernels at the decoder's shapes."""
    _tools()
    import bench_wgrad
    return bench_wgrad.collect()


def tail_context():
    """silog fwd/bwd, the nine metrics, fused ELU + concat1 fwd/bwd, the last convolution fwd/bwd at B=32, 480x640, float32."""
    _tools()
    import bench_tail
    r = bench_tail.collect(["--skip-cpu", "--skip-literal", "--steps", "64"])
    keep = ("silog_fwd", "silog_bwd", "eval_metrics", "concat1_fwd", "concat1_bwd", "depthconv_fwd_C32", "depthconv_bwd_C32",
            "depthconv_fwd_C16", "depthconv_bwd_C16")
    return {k: {kk: r[k][kk] for kk in ("us", "GBps", "frac_of_peak", "kernel") if kk in r[k]} for k in keep if k in r}


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    import torch.distributed as dist
    from bts_fully_tf_b200 import ops
    from bts_fully_tf_b200.host_io import DeviceSet, HostLpgPipeline, HostSet, algorithmic_bytes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the LPG path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created; keep stdout
        # for the one JSON line by pointing fd 1 at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from bts_fully_tf_b200 import _cabi
    _cabi.load()          # fail loudly now if libbtslpg.so is missing: there is no fallback path

    dtype = torch.float32 if a.dtype == "f32" else torch.bfloat16
    es = 4 if a.dtype == "f32" else 2
    B, H, W, K, WU = a.batch, a.height, a.width, a.steps, max(a.warmup, 3)
    fwd_b, bwd_b, per_layer = algorithmic_bytes(B, H, W, es)
    step_bytes = fwd_b + bwd_b

    cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
    nvml_index = local_rank
    if cvd:
        ids = [x.strip() for x in cvd.split(",") if x.strip()]
        if local_rank < len(ids) and ids[local_rank].isdigit():
            nvml_index = int(ids[local_rank])
    sampler = ClockSampler(nvml_index)
    sampler.start()

    gen = torch.Generator(device=device).manual_seed(rank)
    all_sets = sets = [DeviceSet(B, H, W, dtype, device, generator=gen) for _ in range(a.sets)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_steps(run, n_steps, n_warm):
        """barrier+sync, exactly n_steps steps between two CUDA events on the current stream, barrier+sync."""
        if hasattr(run, "prepare"):
            run.prepare(n_steps)
            run.prepare(n_warm)
        run(n_warm)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.hot = True
        e0.record()
        run(n_steps)
        e1.record()
        barrier()
        sampler.hot = False
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def make_runner(fused, what="both", sets=None):
        sets = sets if sets is not None else all_sets
        """Returns (run(n_steps), launches_per_step).  Steps rotate over the buffer sets.  With CUDA graphs
        (default) a chunk of len(sets) consecutive steps is one graph, so the timed region holds
        back-to-back kernels and not one host call per kernel."""
        def body(s):
            if what in ("both", "fwd"):
                s.forward(fused)
            if what in ("both", "bwd"):
                s.backward(fused)
        ops.reset_launch_count()
        body(sets[0])
        per_step = ops.launch_count()
        if a.no_graph:
            def run_plain(n):
                for k in range(n):
                    body(sets[k % len(sets)])
            return run_plain, per_step
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream())
        cache = {}

        def graph_of(n):
            if n not in cache:
                with torch.cuda.stream(side):
                    gph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gph, stream=side):
                        for k in range(n):
                            body(sets[k % len(sets)])
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                cache[n] = gph
            return cache[n]

        chunk = len(sets) * a.graph_repeat

        def prepare(n):
            graph_of(chunk)
            if n % chunk:
                graph_of(n % chunk)

        def run_graph(n):
            for _ in range(n // chunk):
                cache[chunk].replay()
            if n % chunk:
                cache[n % chunk].replay()
        run_graph.prepare = prepare
        return run_graph, per_step

    # ---- headline: device-resident fwd+bwd of the three scales
    results = {}
    modes = ["multi", "per-layer"] if a.mode == "auto" else [a.mode]
    for m in modes:
        fn, per_step = make_runner(m == "multi")
        ms = time_steps(fn, K, WU)
        results[m] = {"ms_per_step": ms / K, "gbps": world * step_bytes * K / (ms * 1e-3) / 1e9, "launches_per_step": per_step}
    best = max(results, key=lambda m: results[m]["gbps"])
    value = results[best]["gbps"]
    ms_per_step = results[best]["ms_per_step"]
    gpu_launches = results[best]["launches_per_step"] * K

    # ---- roofline of the dominant kernel: the backward launch (most bytes), timed alone, back to back
    peak, peak_src = measured_peak()
    if best == "multi":
        dom_name, dom_bytes = "lpg_bwd_multi<%s,n3>" % a.dtype, bwd_b
        fn_dom, _ = make_runner(True, "bwd")
    else:
        dom_name, dom_bytes = "lpg_bwd_vec<%s,r2>" % a.dtype, [b for (r, f, b) in per_layer if r == 2][0]
        only2 = [DeviceSet.__new__(DeviceSet) for _ in sets]
        for o, s in zip(only2, sets):
            o.layers = [L for L in s.layers if L["upratio"] == 2]
        fn_dom, _ = make_runner(False, "bwd", only2)
    n_dom = max(K, 50)
    ms_dom = time_steps(fn_dom, n_dom, WU)
    achieved = dom_bytes / (ms_dom / n_dom * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "us_per_launch": round(ms_dom / n_dom * 1e3, 2),
                "step_frac_of_peak": round(value / world / peak, 4)}
    tr = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get(dom_name)
        except Exception:
            pass

    extras = {"modes": {m: {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()} for m, r in results.items()}}

    # ---- per-kernel breakdown (same timing method), for DESIGN.md / BASELINE.md tables
    if not a.skip_extras:
        per = {}
        for what, nb in (("fwd", fwd_b), ("bwd", bwd_b)):
            fnw, _ = make_runner(best == "multi", what)
            msw = time_steps(fnw, max(K, 50), WU)
            per[what] = {"us": round(msw / max(K, 50) * 1e3, 2), "GBps": round(nb / (msw / max(K, 50) * 1e-3) / 1e9, 1),
                         "frac_of_peak": round(nb / (msw / max(K, 50) * 1e-3) / 1e9 / peak, 4)}
        extras["per_pass"] = per

    # ---- bf16 I/O (float32 arithmetic) on the same launches: half the bytes, same instruction count
    if not a.skip_extras and a.dtype == "f32" and best == "multi":
        try:
            bsets = [DeviceSet(B, H, W, torch.bfloat16, device, generator=gen) for _ in range(a.sets)]
            bf = {}
            for what, nb in (("both", step_bytes // 2), ("fwd", fwd_b // 2), ("bwd", bwd_b // 2)):
                fnb, _ = make_runner(True, what, bsets)
                msb = time_steps(fnb, max(K, 50), WU)
                us = msb / max(K, 50) * 1e3
                bf[what] = {"us": round(us, 2), "GBps": round(nb / us / 1e3, 1), "frac_of_peak": round(nb / us / 1e3 / peak, 4)}
            extras["bf16"] = {"workload": "the same three-scale launches with bfloat16 tensors (float32 arithmetic)", **bf}
            del bsets
        except Exception as exc:  # noqa: BLE001
            extras["bf16"] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}

    # ---- end to end: host buffers in, host buffers out, through the public API
    def time_e2e(run_kernels=True, return_ds=True, slots=3, n_e2e=20):
        pipe = HostLpgPipeline(B, H, W, dtype, device, slots=slots, fused=(best == "multi"), return_ds=return_ds, run_kernels=run_kernels)
        host = HostSet(sets[0])
        for _ in range(3):
            pipe.step(host)
        pipe.drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.hot = True
        e0.record()
        pipe.s_in.wait_event(e0)
        for _ in range(n_e2e):
            pipe.step(host)
        pipe.join()
        e1.record()
        barrier()
        sampler.hot = False
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        chk = float(host.layers[-1]["g_coef"].float().abs().sum())       # the results really are on the host
        out = {"value": round(world * step_bytes * n_e2e / (ms * 1e-3) / 1e9, 2), "ms_per_step": round(ms / n_e2e, 3), "steps": n_e2e,
               "h2d": host.bytes_in(), "d2h": host.bytes_out(return_ds), "checksum": chk,
               "pcie_GBps_per_gpu": round((host.bytes_in() + host.bytes_out(return_ds)) * n_e2e / (ms * 1e-3) / 1e9, 2)}
        del pipe, host
        return out

    e2e = None
    if not a.skip_e2e:
        n_e2e = max(3, min(K, 20))
        main_leg = time_e2e(True, True, 3, n_e2e)
        copy_leg = time_e2e(False, True, 3, n_e2e)          # same buffers / streams / events, no kernels: the platform's ceiling
        e2e = {"value": main_leg["value"], "unit": UNIT, "h2d_bytes_per_step": main_leg["h2d"], "d2h_bytes_per_step": main_leg["d2h"],
               "steps": n_e2e, "ms_per_step": main_leg["ms_per_step"], "result_checksum": main_leg["checksum"],
               "pcie_GBps_per_gpu": main_leg["pcie_GBps_per_gpu"],
               "copy_only_ceiling": {"value": copy_leg["value"], "ms_per_step": copy_leg["ms_per_step"], "pcie_GBps_per_gpu": copy_leg["pcie_GBps_per_gpu"],
                                     "what": "the same pinned buffers, three streams and events with the kernel launches removed"},
               "frac_of_copy_ceiling": round(main_leg["value"] / copy_leg["value"], 4) if copy_leg["value"] else None,
               "path": "bts_fully_tf_b200.host_io.HostLpgPipeline -> libbtslpg.so (pinned host in/out, 3 streams, 3 device slots)"}
        if not a.skip_extras:
            try:
                nods = time_e2e(True, False, 3, n_e2e)
                extras["e2e_without_ds_readback"] = {"value": nods["value"], "ms_per_step": nods["ms_per_step"], "d2h_bytes_per_step": nods["d2h"],
                                                     "note": "out_ds = out_full[:, ::d, ::d] is not copied back (the host can slice it)"}
            except Exception as exc:  # noqa: BLE001
                extras["e2e_without_ds_readback"] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}

    sampler.stop_flag = True
    sampler.join(timeout=2)

    del all_sets, sets
    torch.cuda.empty_cache()

    # ---- the kernels the decoder actually launches (fused heads) and the decoder-tail kernels, same hygiene (N = 1 only)
    if not a.skip_extras and world == 1 and a.dtype == "f32":
        for key, fn in (("heads", lambda: heads_context(peak)), ("tail", tail_context), ("iconv1_tcgen05", iconv_context),
                        ("wgrad_tcgen05", wgrad_context)):
            try:
                extras[key] = fn()
            except Exception as exc:  # noqa: BLE001
                extras[key] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
            torch.cuda.empty_cache()

    # ---- context for the second half of BASELINE.json's metric ("BTS decoder images/s at 1/2/4/8 B200"): config 3 (batched
    # inference) and the data-parallel training steps of configs 4 / 5 -- the only lines with a collective (the gradient
    # all-reduce); a failure here never touches the headline line
    if not a.skip_decoder:
        for cfg_id in (3,) + (() if a.skip_train else (5, 4)):
            try:
                extras["decoder_config%d" % cfg_id] = decoder_context(cfg_id, rank, local_rank, world, steps=a.decoder_steps)
            except Exception as exc:  # noqa: BLE001
                extras["decoder_config%d" % cfg_id] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
            torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not a.skip_cpu:
        cpu = time_cpu_baseline(a, a.cpu_seconds)

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": WU,
            "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": workload_name(a), "launch_mode": best,
                       "cache": "%d disjoint buffer sets of %.0f MB rotated between steps (each exceeds the 126 MB L2)" % (
                           a.sets, (step_bytes) / 1e6),
                       "cuda_graph": not a.no_graph, "steps_per_graph": (0 if a.no_graph else a.sets * a.graph_repeat), "parallelism": "dp%d (batch shards, no data-path collective)" % world,
                       "algorithmic_bytes_per_step": step_bytes},
            "clocks": sampler.summary(),
            "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu, "extras": extras,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
