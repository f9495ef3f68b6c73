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

import custom_layers  # noqa: E402  (the reference file, unmodified)
import bts_decoder  # noqa: E402    (the reference file, unmodified)
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


if __name__ == "__main__":
    golden_lpg()
    golden_pole()
    golden_decoder()
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print("total fixture bytes:", tot)
