#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference source.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):        python tests/golden/make_golden.py

How: oracle/tf_shim (a torch-CPU stand-in for the few tf symbols the reference
imports) is put on sys.path, then /root/reference/custom_layers.py and
/root/reference/bts_decoder.py are imported as they are and executed.  Inputs are
seeded; outputs and autograd gradients (TF autodiff's counterpart) are saved.

Files written
  lpg_r{8,4,2}.npz    one LocalPlanarGuidance layer (custom_layers.py:25-61) + the strided
                      slice of bts_decoder.py:81,88; fp32 run and a float64 run
  lpg_pole_r8.npz     U(0,1) coefficients that reach the theta->pi/3 pole (den <= 0)
  decoder_small.npz   bts_decoder.py:26-105 whole, num_filters=32, float64 run on
                      float32-representable inputs/weights, inference BN and training BN
  decoder_f256.npz    the same at num_filters=256 (the channel counts of the fused fast paths); kernels are redrawn by
                      oracle/decoder_fixture.regen_kernels (checked bit for bit here), gradients of large kernels sampled
  lpg_full_size_samples_shapes.npz   the same at 2 x 352 x 1216 (KITTI Eigen) and 2 x 416 x 544 (NYU training crop)
  lpg_full_size_samples.npz   the layer at FULL size (2 x 480 x 640, r = 8/4/2, seeded numpy inputs): 4096 sampled
                      outputs per scale instead of the 2.4 MB maps (the "full-size" pin of SURVEY 8(c))
  tail_silog.npz      bts.py:27-41 si_log_loss (nyu and kitti thresholds) on depth_est =
                      sigmoid(logit)*max_depth (bts_decoder.py:102-103): loss and autograd gradients
                      with respect to depth_est and to the logit, fp32 and float64 runs
  tail_metrics.npz    custom_eval_metrics.py:21-88, the nine metrics on maps that contain
                      out-of-range ground truth and NaN / inf / out-of-range predictions
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("BTS_REFERENCE", "/root/reference")

sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import custom_layers  # noqa: E402  (the reference file, unmodified)
import bts_decoder  # noqa: E402    (the reference file, unmodified)
import bts  # noqa: E402            (the reference file, unmodified: si_log_loss_wrapper)
import custom_eval_metrics  # noqa: E402  (the reference file, unmodified)
from tensorflow.keras import layers as shim_layers  # noqa: E402


def run_lpg(coef, r, d, g_full, g_ds):
    x = coef.clone().requires_grad_(True)
    layer = custom_layers.LocalPlanarGuidance(upratio=r, name="depth_%dx%d_scaled" % (r, r))
    out = layer(x)
    outs, grads = [out], [g_full.to(out.dtype)]
    ds = None
    if d:
        ds = (lambda t: t[:, ::d, ::d, ...])(out)          # bts_decoder.py:81,88
        outs.append(ds)
        grads.append(g_ds.to(out.dtype))
    torch.autograd.backward(outs, grads)
    return layer, out.detach(), None if ds is None else ds.detach(), x.grad


def golden_lpg():
    for r, d in [(8, 4), (4, 2), (2, 0)]:
        g = torch.Generator().manual_seed(100 + r)
        B, h, w = 2, 5, 7
        coef = torch.sigmoid(torch.randn(B, h, w, 3, generator=g))
        g_full = torch.randn(B, h * r, w * r, 1, generator=g)
        g_ds = torch.randn(B, h * r // d, w * r // d, 1, generator=g) if d else None
        layer, out, ds, gc = run_lpg(coef, r, d, g_full, g_ds)
        _, out64, ds64, gc64 = run_lpg(coef.double(), r, d, g_full.double(), None if g_ds is None else g_ds.double())
        cfg = layer.get_config()
        np.savez(os.path.join(HERE, "lpg_r%d.npz" % r),
                 upratio=r, ds_stride=d, coef=coef.numpy(), g_full=g_full.numpy(),
                 g_ds=np.zeros(0, np.float32) if g_ds is None else g_ds.numpy(),
                 out=out.numpy(), out_ds=np.zeros(0, np.float32) if ds is None else ds.numpy(),
                 g_coef=gc.numpy(), out64=out64.numpy(), g_coef64=gc64.numpy(),
                 pixel_dir_unit=layer.pixel_dir_unit.numpy(),
                 config_name=cfg["name"], config_upratio=cfg["upratio"])
        print("lpg_r%d: out %s |g_coef|max %.3g" % (r, tuple(out.shape), float(gc.abs().max())))


def golden_pole():
    g = torch.Generator().manual_seed(7)
    r = 8
    coef = torch.rand(2, 12, 16, 3, generator=g)
    coef[0, 0, 0] = torch.tensor([0.375, 1.0, 0.5])       # theta = pi/3 exactly at the pole side
    coef[0, 0, 1] = torch.tensor([0.875, 0.999, 0.0])     # n4 == 0 next to the pole
    layer = custom_layers.LocalPlanarGuidance(upratio=r)
    out = layer(coef)
    out64 = custom_layers.LocalPlanarGuidance(upratio=r)(coef.double())
    np.savez_compressed(os.path.join(HERE, "lpg_pole_r8.npz"), upratio=r, coef=coef.numpy(), out=out.numpy(), out64=out64.numpy())
    print("lpg_pole_r8: min out %.3g  n(out<0)=%d" % (float(out.min()), int((out < 0).sum())))


def golden_decoder():
    F = 32
    B, H, W = 1, 64, 64
    chans = dict(dense=16, s2=6, s4=6, s8=8, s16=12)
    out = {}
    for tag, training in (("infer", False), ("train", True)):
        shim_layers.reset(seed=1234, dtype=torch.float64)
        g = torch.Generator().manual_seed(99)
        mk = lambda s, c: torch.randn(B, H // s, W // s, c, generator=g).float().double().requires_grad_(True)  # noqa: E731
        feats = [mk(32, chans["dense"]), mk(2, chans["s2"]), mk(4, chans["s4"]), mk(8, chans["s8"]), mk(16, chans["s16"])]
        depth = bts_decoder.decoder_model(feats, 10.0, num_filters=F, is_training=training)
        g_depth = torch.randn(depth.shape, generator=g).float().double()
        depth.backward(g_depth)
        convs = [l for l in shim_layers.CREATED if isinstance(l, shim_layers.Conv2D)]
        named = {l.name: l for l in shim_layers.CREATED}
        if tag == "infer":
            for k, f in zip(["dense", "s2", "s4", "s8", "s16"], feats):
                out["feat_" + k] = f.detach().numpy()
            out["g_depth"] = g_depth.numpy()
            out["n_convs"] = len(convs)
            for i, l in enumerate(convs):
                out["kernel_%02d" % i] = l.kernel.detach().numpy()     # HWIO
        heads = [l for l in convs if l.filters == 3]
        assert len(heads) == 3
        for r, l in zip((8, 4, 2), heads):
            out["%s_head%d_in" % (tag, r)] = l.last_input.detach().numpy()
            out["%s_head%d_out" % (tag, r)] = l.last_output.detach().numpy()
            out["%s_head%d_gkernel" % (tag, r)] = l.kernel.grad.numpy()
            out["head%d_conv_index" % r] = convs.index(l)
        for r in (8, 4, 2):
            out["%s_depth_%dx%d_scaled" % (tag, r, r)] = named["depth_%dx%d_scaled" % (r, r)].last_output.detach().numpy()
        out[tag + "_depth_est"] = depth.detach().numpy()
        for i, l in enumerate(convs):
            out["%s_gkernel_%02d" % (tag, i)] = l.kernel.grad.numpy()
        for k, f in zip(["dense", "s2", "s4", "s8", "s16"], feats):
            out["%s_gfeat_%s" % (tag, k)] = f.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "decoder_small.npz"), **out)
    print("decoder_small: %d convs, depth %s" % (out["n_convs"], out["infer_depth_est"].shape))


def golden_decoder_f256():
    """bts_decoder.py:26-105 whole at num_filters = 256 (F/16 = 16, heads at C = 64 / 64 / 32): the channel counts that take the
    fused fast paths (head kernels, concat with pad channels, sub-pixel upconv, last convolution with the ELU folded in).
    Kernels are NOT stored (2 M weights): oracle/decoder_fixture.regen_kernels redraws them; checked here bit for bit."""
    from oracle import decoder_fixture
    F, seed = 256, 4321
    B, H, W = 1, 64, 64
    chans = dict(dense=16, s2=6, s4=6, s8=8, s16=12)
    out = {"seed": seed, "num_filters": F}
    for tag, training in (("infer", False), ("train", True)):
        shim_layers.reset(seed=seed, dtype=torch.float64)
        g = torch.Generator().manual_seed(77)
        mk = lambda s, c: torch.randn(B, H // s, W // s, c, generator=g).float().double().requires_grad_(True)  # noqa: E731
        feats = [mk(32, chans["dense"]), mk(2, chans["s2"]), mk(4, chans["s4"]), mk(8, chans["s8"]), mk(16, chans["s16"])]
        depth = bts_decoder.decoder_model(feats, 10.0, num_filters=F, is_training=training)
        g_depth = torch.randn(depth.shape, generator=g).float().double()
        depth.backward(g_depth)
        convs = [l for l in shim_layers.CREATED if isinstance(l, shim_layers.Conv2D)]
        named = {l.name: l for l in shim_layers.CREATED}
        shapes = [tuple(l.kernel.shape) for l in convs]
        regen = decoder_fixture.regen_kernels(shapes, seed)
        for l, k in zip(convs, regen):
            assert torch.equal(l.kernel.detach(), k), "regen_kernels does not reproduce the recorded run"
        if tag == "infer":
            for k, f in zip(["dense", "s2", "s4", "s8", "s16"], feats):
                out["feat_" + k] = f.detach().numpy().astype(np.float32)
            out["g_depth"] = g_depth.numpy().astype(np.float32)
            out["kernel_shapes"] = np.array(shapes, np.int64)
            out["kernel_sums"] = np.array([float(k.sum()) for k in regen], np.float64)
        heads = [l for l in convs if l.filters == 3]
        for r, l in zip((8, 4, 2), heads):
            out["%s_head%d_out" % (tag, r)] = l.last_output.detach().numpy()
        for r in (8, 4, 2):
            out["%s_depth_%dx%d_scaled" % (tag, r, r)] = named["depth_%dx%d_scaled" % (r, r)].last_output.detach().numpy().astype(np.float32)
        out[tag + "_depth_est"] = depth.detach().numpy()
        for i, l in enumerate(convs):
            gk = l.kernel.grad.reshape(-1)
            idx = decoder_fixture.sample_index(gk.numel())
            out["%s_gkernel_%02d" % (tag, i)] = gk[idx].numpy()
            out["%s_gkernel_absmax_%02d" % (tag, i)] = float(gk.abs().max())
            out["%s_gkernel_sum_%02d" % (tag, i)] = float(gk.sum())
        for k, f in zip(["dense", "s2", "s4", "s8", "s16"], feats):
            out["%s_gfeat_%s" % (tag, k)] = f.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "decoder_f256.npz"), **out)
    print("decoder_f256: %d convs, %d weights regenerated, depth %s, %d bytes" % (
        len(shapes), sum(int(np.prod(s)) for s in shapes), out["infer_depth_est"].shape, os.path.getsize(os.path.join(HERE, "decoder_f256.npz"))))


def full_size_inputs(r, B=2, H=480, W=640):
    """Seeded inputs of the full-size pin (numpy Generator streams are stable across platforms)."""
    rng = np.random.default_rng(4000 + r + (0 if (H, W) == (480, 640) else H * 10000 + W))
    z = rng.standard_normal((B, H // r, W // r, 3)).astype(np.float32)
    coef = (1.0 / (1.0 + np.exp(-z.astype(np.float64)))).astype(np.float32)
    idx = np.sort(rng.choice(B * H * W, size=4096, replace=False))
    return coef, idx


def golden_full_size():
    out = {}
    for r in (8, 4, 2):
        coef, idx = full_size_inputs(r)
        layer = custom_layers.LocalPlanarGuidance(upratio=r)
        y = layer(torch.from_numpy(coef)).numpy().reshape(-1)
        y64 = custom_layers.LocalPlanarGuidance(upratio=r)(torch.from_numpy(coef).double()).numpy().reshape(-1)
        out["r%d_idx" % r] = idx
        out["r%d_out" % r] = y[idx]
        out["r%d_out64" % r] = y64[idx]
        out["r%d_n_negative" % r] = int((y64 < 0).sum())
        print("lpg_full_size r=%d: sample mean %.6f, negatives %d" % (r, float(y64[idx].mean()), int((y64 < 0).sum())))
    np.savez_compressed(os.path.join(HERE, "lpg_full_size_samples.npz"), **out)


def golden_full_size_shapes():
    """The same pin at the reference's other two input sizes: KITTI Eigen 352 x 1216 (args/test_eigen.txt:8-9) and the NYU
    training crop 416 x 544 (args/train_nyu.txt:13-14).  A separate file: lpg_full_size_samples.npz stays bit-identical."""
    out = {}
    for H, W in ((352, 1216), (416, 544)):
        for r in (8, 4, 2):
            coef, idx = full_size_inputs(r, 2, H, W)
            y = custom_layers.LocalPlanarGuidance(upratio=r)(torch.from_numpy(coef)).numpy().reshape(-1)
            y64 = custom_layers.LocalPlanarGuidance(upratio=r)(torch.from_numpy(coef).double()).numpy().reshape(-1)
            key = "h%dw%d_r%d" % (H, W, r)
            out[key + "_idx"] = idx
            out[key + "_out"] = y[idx]
            out[key + "_out64"] = y64[idx]
            out[key + "_n_negative"] = int((y64 < 0).sum())
            print("lpg_full_size %dx%d r=%d: sample mean %.6f, negatives %d" % (H, W, r, float(y64[idx].mean()), int((y64 < 0).sum())))
    np.savez_compressed(os.path.join(HERE, "lpg_full_size_samples_shapes.npz"), **out)


def golden_tail():
    g = torch.Generator().manual_seed(2024)
    out = {}
    for dataset, max_depth, shape in (("nyu", 10.0, (2, 13, 18, 1)), ("kitti", 80.0, (1, 11, 23, 1))):
        logit = torch.randn(shape, generator=g) * 1.5
        y_true = torch.rand(shape, generator=g) * max_depth * 1.05
        y_true[torch.rand(shape, generator=g) < 0.3] = 0.0                 # missing ground truth (below the threshold)
        loss_fn = bts.si_log_loss_wrapper(dataset)
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            z = logit.detach().clone().to(dt).requires_grad_(True)
            y_pred = torch.sigmoid(z) * max_depth                          # bts_decoder.py:102-103
            y_pred.retain_grad()
            loss = loss_fn(y_true.to(dt), y_pred)
            loss.backward()
            out["%s_%s_loss" % (dataset, tag)] = loss.detach().numpy()
            out["%s_%s_depth_est" % (dataset, tag)] = y_pred.detach().numpy()
            out["%s_%s_g_depth" % (dataset, tag)] = y_pred.grad.numpy()
            out["%s_%s_g_logit" % (dataset, tag)] = z.grad.numpy()
        out[dataset + "_logit"] = logit.numpy()
        out[dataset + "_y_true"] = y_true.numpy()
        out[dataset + "_max_depth"] = max_depth
        print("tail_silog %s: loss %.6f" % (dataset, float(out[dataset + "_f64_loss"])))
    np.savez(os.path.join(HERE, "tail_silog.npz"), **out)

    class Args:                                                            # the argparse names bts_eval.py passes
        min_depth_eval, max_depth_eval = 1e-3, 10.0
        garg_crop = eigen_crop = False
        dataset = "nyu"
    shape = (2, 17, 22, 1)
    y_true = torch.rand(shape, generator=g) * 12.0                         # some beyond max_depth_eval
    y_true[torch.rand(shape, generator=g) < 0.25] = 0.0
    y_pred = y_true * torch.exp(torch.randn(shape, generator=g) * 0.3) + 0.05
    y_pred.view(-1)[3] = float("nan")
    y_pred.view(-1)[10] = float("inf")
    y_pred.view(-1)[20] = -float("inf")
    y_pred.view(-1)[30] = 50.0
    y_pred.view(-1)[40] = -1.0
    y_true.view(-1)[[3, 10, 20, 30, 40]] = torch.tensor([2.0, 3.0, 4.0, 5.0, 6.0])
    out = {"y_true": y_true.numpy(), "y_pred": y_pred.numpy(), "min_depth_eval": Args.min_depth_eval, "max_depth_eval": Args.max_depth_eval}
    fns = custom_eval_metrics.metrics_list_factory(Args)
    out["names"] = np.array([f.__name__ for f in fns])
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        out["values_" + tag] = np.array([float(f(y_true.to(dt), y_pred.to(dt))) for f in fns], np.float64)
    np.savez(os.path.join(HERE, "tail_metrics.npz"), **out)
    print("tail_metrics:", dict(zip(out["names"], np.round(out["values_f64"], 5))))


if __name__ == "__main__":
    golden_full_size()
    golden_full_size_shapes()
    golden_tail()
    golden_lpg()
    golden_pole()
    golden_decoder()
    golden_decoder_f256()
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print("total fixture bytes:", tot)
