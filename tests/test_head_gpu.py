"""GPU parity tests of the fused reduction-head + LPG kernels (bts_decoder.py:79-81, 86-88, 93-94)."""
import numpy as np
import pytest
import torch

import bts_fully_tf_b200 as pkg
from bts_fully_tf_b200 import ops
from oracle import c_oracle
import lpg_parity as parity

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RS = [(8, 4), (4, 2), (2, 0)]


def npf(t):
    return t.detach().float().cpu().numpy()


def make(B, h, w, C, r, d, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    feat = torch.nn.functional.elu(torch.randn(B, h, w, C, generator=g)).to(dtype)     # post-ELU like iconv*/daspp_feat
    lim = (6.0 / (C + 3)) ** 0.5
    kern = (torch.rand(C, 3, generator=g) * 2 - 1) * lim                                 # glorot_uniform
    g_full = torch.randn(B, h * r, w * r, 1, generator=g).to(dtype)
    g_ds = torch.randn(B, h * r // d, w * r // d, 1, generator=g).to(dtype) if d else None
    return feat, kern, g_full, g_ds


def oracle_head(feat, kern, g_full, g_ds, r, d, coef_seen=None):
    """float64 oracle of the fused op.  coef_seen: the coefficients as stored by the kernel (bf16 case)."""
    f, k = npf(feat), kern.numpy()
    coef = c_oracle.head_forward_f64(f, k)
    cs = coef if coef_seen is None else np.asarray(coef_seen, np.float64)
    g_coef = c_oracle.lpg_backward_f64(cs, npf(g_full), r, None if g_ds is None else npf(g_ds), d)
    g_feat, g_w = c_oracle.head_backward_f64(f, k, cs, g_coef)
    return coef, g_coef, g_feat, g_w


@pytest.mark.parametrize("r,d", RS)
@pytest.mark.parametrize("C", [32, 64, 128])
@pytest.mark.parametrize("B,h,w", [(2, 6, 10), (1, 13, 17)])
def test_fused_head_f32(B, h, w, C, r, d):
    feat, kern, g_full, g_ds = make(B, h, w, C, r, d, seed=C + r)
    f, k, gf = feat.to(DEV), kern.to(DEV), g_full.to(DEV)
    gd = g_ds.to(DEV) if d else None
    coef, full, ds = ops.reduce_lpg_forward(f, k, r, d)
    assert ops.last_kernel().startswith("head_lpg_fwd") and ("<f32,r%d," % r) in ops.last_kernel(), ops.last_kernel()
    ref_coef, ref_gc, ref_gf, ref_gw = oracle_head(feat, kern, g_full, g_ds, r, d, coef_seen=npf(coef))
    np.testing.assert_allclose(npf(coef), ref_coef, rtol=2e-6, atol=1e-7)
    parity.check_forward(npf(full), npf(coef), r, what="fused fwd")
    # identical to running the stand-alone LPG kernel on the stored coefficients
    full2, ds2 = ops.lpg_forward(coef, r, d)
    assert torch.equal(full, full2)
    if d:
        assert torch.equal(ds, ds2) and torch.equal(ds, full[:, ::d, ::d])

    g_feat, g_kern, g_coef = ops.reduce_lpg_backward(f, k, coef, gf, gd, r, d, need_g_coef=True)
    assert ops.last_kernel().startswith("head_lpg_bwd") and ("<f32,r%d," % r) in ops.last_kernel(), ops.last_kernel()
    # same per-pixel terms as the stand-alone kernel; at r=8 the latter adds the patch rows in a lane-group tree
    alone = ops.lpg_backward(coef, gf, gd, r, d)
    scale, _ = parity.backward_scale(npf(coef), npf(g_full), r, npf(g_ds) if d else None, d)
    assert (np.abs(npf(g_coef) - npf(alone)) <= 4e-7 * scale + 1e-30).all()
    if r != 8:
        assert torch.equal(g_coef, alone)
    parity.check_backward(npf(g_coef), npf(coef), npf(g_full), r, npf(g_ds) if d else None, d)
    # g_feat / g_kernel: compare with the oracle chain fed by the kernel's own (float32) g_coef
    gf_o, gw_o = c_oracle.head_backward_f64(npf(feat), kern.numpy(), npf(coef), npf(g_coef))
    np.testing.assert_allclose(npf(g_feat), gf_o, rtol=1e-5, atol=1e-5 * np.abs(gf_o).max())
    absw = np.abs(npf(feat)).reshape(-1, C).T @ np.abs(npf(g_coef).reshape(-1, 3) * npf(coef).reshape(-1, 3) * (1 - npf(coef).reshape(-1, 3)))
    assert (np.abs(npf(g_kern) - gw_o) <= 1e-5 * absw + 1e-30).all()
    # and end to end against the pure float64 chain (looser: float32 g_coef feeds the head)
    assert np.abs(npf(g_kern) - ref_gw).max() <= 1e-4 * np.abs(ref_gw).max() + 1e-6 * absw.max()


@pytest.mark.parametrize("r,d", RS)
def test_fused_head_deterministic_dw_and_partial_outputs(r, d):
    feat, kern, g_full, g_ds = make(4, 24, 32, 64, r, d, seed=1)
    f, k, gf = feat.to(DEV), kern.to(DEV), g_full.to(DEV)
    gd = g_ds.to(DEV) if d else None
    coef, full, ds = ops.reduce_lpg_forward(f, k, r, d)
    a = ops.reduce_lpg_backward(f, k, coef, gf, gd, r, d)
    b = ops.reduce_lpg_backward(f, k, coef, gf, gd, r, d)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])          # no float atomics: bit-reproducible
    only_k = ops.reduce_lpg_backward(f, k, coef, gf, gd, r, d, need_g_feat=False)
    only_f = ops.reduce_lpg_backward(f, k, coef, gf, gd, r, d, need_g_kernel=False)
    assert only_k[0] is None and torch.equal(only_k[1], a[1])
    assert only_f[1] is None and torch.equal(only_f[0], a[0])
    # g_kernel written straight into a slice of a flat gradient bucket
    bucket = torch.zeros(1000, device=DEV)
    view = bucket[100:100 + 64 * 3].view(64, 3)
    ops.reduce_lpg_backward(f, k, coef, gf, gd, r, d, need_g_feat=False, g_kernel_out=view)
    assert torch.equal(view, a[1]) and (bucket[:100] == 0).all() and (bucket[292:] == 0).all()


@pytest.mark.parametrize("r,d", RS)
def test_fused_head_bf16(r, d):
    C = 64
    feat, kern, g_full, g_ds = make(2, 12, 16, C, r, d, seed=7, dtype=torch.bfloat16)
    f, k, gf = feat.to(DEV), kern.to(DEV), g_full.to(DEV)
    gd = g_ds.to(DEV) if d else None
    coef, full, ds = ops.reduce_lpg_forward(f, k, r, d)
    assert "bf16" in ops.last_kernel() and coef.dtype == torch.bfloat16
    ref_coef = c_oracle.head_forward_f64(npf(feat), kern.numpy())
    np.testing.assert_allclose(npf(coef), ref_coef, rtol=1e-2)
    parity.check_forward(npf(full), npf(coef), r, rtol=1e-2, what="bf16 fused fwd")
    g_feat, g_kern, g_coef = ops.reduce_lpg_backward(f, k, coef, gf, gd, r, d, need_g_coef=True)
    parity.check_backward(npf(g_coef), npf(coef), npf(g_full), r, npf(g_ds) if d else None, d, rtol=1e-2)
    gf_o, gw_o = c_oracle.head_backward_f64(npf(feat), kern.numpy(), npf(coef), npf(g_coef))
    np.testing.assert_allclose(npf(g_feat), gf_o, rtol=1e-2, atol=1e-2 * np.abs(gf_o).max())
    assert np.abs(npf(g_kern) - gw_o).max() <= 1e-2 * np.abs(gw_o).max()


@pytest.mark.parametrize("C", [3, 48, 160])
def test_generic_channel_counts(C):
    """C not in {32, 64, 128}: generic kernels (1x1 conv + sigmoid, then the LPG dispatch)."""
    r, d = 4, 2
    feat, kern, g_full, g_ds = make(2, 5, 6, C, r, d, seed=C)
    f, k, gf, gd = feat.to(DEV), kern.to(DEV), g_full.to(DEV), g_ds.to(DEV)
    coef, full, ds = ops.reduce_lpg_forward(f, k, r, d)
    ref_coef = c_oracle.head_forward_f64(npf(feat), kern.numpy())
    np.testing.assert_allclose(npf(coef), ref_coef, rtol=2e-6, atol=1e-7)
    parity.check_forward(npf(full), npf(coef), r)
    g_feat, g_kern, g_coef = ops.reduce_lpg_backward(f, k, coef, gf, gd, r, d, need_g_coef=True)
    gf_o, gw_o = c_oracle.head_backward_f64(npf(feat), kern.numpy(), npf(coef), npf(g_coef))
    np.testing.assert_allclose(npf(g_feat), gf_o, rtol=1e-5, atol=1e-5 * np.abs(gf_o).max())
    assert np.abs(npf(g_kern) - gw_o).max() <= 2e-5 * np.abs(gw_o).max()
    # the Python wrapper supplies the scratch the generic path needs; the raw ABI refuses without it
    from bts_fully_tf_b200 import _cabi
    lib = _cabi.load()
    refs = [_cabi.as_ref(t) for t in (f, k, coef, gf, gd, torch.empty_like(f))]
    rc = lib.btslpg_reduce_backward(refs[0].ptr, refs[1].ptr, refs[2].ptr, refs[3].ptr, refs[4].ptr, r, d, refs[5].ptr, None, None,
                                    None, 0, _cabi.current_stream_ptr(f.device))
    assert rc == -5 and b"g_coef_out" in lib.btslpg_last_error()


def test_keras_hwio_kernel_and_module_autograd():
    """ReductionLPG owns the Conv2D kernel in HWIO (1,1,C,3) like Keras; autograd reaches feat and kernel."""
    torch.manual_seed(0)
    r, d, C = 8, 4, 128
    head = pkg.ReductionLPG(C, r, ds_stride=d).to(DEV)
    feat = torch.nn.functional.elu(torch.randn(2, 6, 8, C, device=DEV)).requires_grad_(True)
    coef, depth, depth_ds = head(feat)
    g_full = torch.randn_like(depth)
    g_ds = torch.randn_like(depth_ds)
    torch.autograd.backward([depth, depth_ds], [g_full, g_ds])
    assert tuple(head.kernel.grad.shape) == (1, 1, C, 3) and feat.grad.shape == feat.shape
    k2 = head.kernel.detach().reshape(C, 3).cpu()
    _, _, ref_gf, ref_gw = oracle_head(feat.detach().cpu(), k2, g_full.cpu(), g_ds.cpu(), r, d, coef_seen=npf(coef))
    assert np.abs(npf(head.kernel.grad).reshape(C, 3) - ref_gw).max() <= 1e-4 * np.abs(ref_gw).max()
    np.testing.assert_allclose(npf(feat.grad), ref_gf, rtol=1e-4, atol=1e-4 * np.abs(ref_gf).max())


def test_decoder_fixture_heads(golden_dir):
    """Head inputs / kernels recorded from the reference decoder run -> fused kernel reproduces
    reduction_NxN and depth_NxN_scaled (C = 8, 8, 4 here: generic-channel path)."""
    import os
    z = np.load(os.path.join(golden_dir, "decoder_small.npz"))
    for r, d in RS:
        k = torch.from_numpy(z["kernel_%02d" % int(z["head%d_conv_index" % r])][0, 0]).float().to(DEV)
        f = torch.from_numpy(z["infer_head%d_in" % r]).float().to(DEV)
        coef, full, ds = ops.reduce_lpg_forward(f, k, r, d)
        np.testing.assert_allclose(npf(coef), z["infer_head%d_out" % r], rtol=5e-6, atol=1e-7)
        np.testing.assert_allclose(npf(full), z["infer_depth_%dx%d_scaled" % (r, r)], rtol=2e-5)


def test_full_size_head_r2_properties():
    """BASELINE config 2 shape for the largest head (r=2, C=64): determinism + sampled oracle parity."""
    B, h, w, C, r = 8, 240, 320, 64, 2
    g = torch.Generator(device=DEV).manual_seed(0)
    feat = torch.nn.functional.elu(torch.randn(B, h, w, C, device=DEV, generator=g))
    kern = (torch.rand(C, 3, device=DEV, generator=g) * 2 - 1) * (6.0 / (C + 3)) ** 0.5
    coef, full, _ = ops.reduce_lpg_forward(feat, kern, r)
    coef2, full2, _ = ops.reduce_lpg_forward(feat, kern, r)
    assert torch.equal(coef, coef2) and torch.equal(full, full2)
    ref = c_oracle.head_forward_f64(npf(feat[:1]), npf(kern))
    np.testing.assert_allclose(npf(coef[:1]), ref, rtol=2e-6, atol=1e-7)
    parity.check_forward(npf(full[:1]), npf(coef[:1]), r)
    gf = torch.randn(B, h * r, w * r, 1, device=DEV, generator=g)
    a = ops.reduce_lpg_backward(feat, kern, coef, gf, None, r)
    b = ops.reduce_lpg_backward(feat, kern, coef, gf, None, r)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # g_kernel == feat^T dz  (checksum over all pixels, float64 on the GPU via torch as plumbing only)
    gc = ops.lpg_backward(coef, gf, None, r)
    dz = (gc.double() * coef.double() * (1 - coef.double())).reshape(-1, 3)
    ref_gw = feat.double().reshape(-1, C).T @ dz
    assert (a[1].double() - ref_gw).abs().max() <= 2e-4 * ref_gw.abs().max()
