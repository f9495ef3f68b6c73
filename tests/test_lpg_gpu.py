"""GPU parity tests of the LPG kernels (forward, backward, multi-layer launches) against the CPU
oracle, through the C ABI.  Run on the B200 box:  python -m pytest tests -m gpu"""
import os

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


def make_inputs(B, h, w, r, d, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    coef = torch.sigmoid(torch.randn(B, h, w, 3, generator=g)).to(dtype)
    g_full = torch.randn(B, h * r, w * r, 1, generator=g).to(dtype)
    g_ds = torch.randn(B, h * r // d, w * r // d, 1, generator=g).to(dtype) if d else None
    return coef, g_full, g_ds


def npf(t):
    return t.detach().float().cpu().numpy()


# ------------------------------------------------------------------------------------------------
# golden fixtures generated from the unmodified reference source
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("r,d", RS)
def test_golden_fixture(golden_dir, r, d):
    z = np.load(os.path.join(golden_dir, "lpg_r%d.npz" % r))
    coef = torch.from_numpy(z["coef"]).to(DEV)
    full, ds = ops.lpg_forward(coef, r, d)
    parity.check_forward(npf(full), z["coef"], r, what="golden r=%d" % r)
    np.testing.assert_allclose(npf(full), z["out"], rtol=1e-5)            # the reference's own float32 result
    np.testing.assert_allclose(npf(full), z["out64"], rtol=1e-5)
    if d:
        assert torch.equal(ds, full[:, ::d, ::d])                          # bts_decoder.py:81,88
        np.testing.assert_allclose(npf(ds), z["out_ds"], rtol=1e-5)
    g_full = torch.from_numpy(z["g_full"]).to(DEV)
    g_ds = torch.from_numpy(z["g_ds"]).to(DEV) if d else None
    gc = ops.lpg_backward(coef, g_full, g_ds, r, d)
    parity.check_backward(npf(gc), z["coef"], z["g_full"], r, z["g_ds"] if d else None, d, what="golden bwd r=%d" % r)
    scale = np.abs(z["g_coef64"]).max()
    assert np.abs(npf(gc) - z["g_coef64"]).max() <= 2e-5 * scale
    assert np.abs(npf(gc) - z["g_coef"]).max() <= 4e-5 * scale             # vs the reference's float32 autograd


def test_golden_pole_no_clamping(golden_dir):
    """den crosses zero at r=8 for theta -> pi/3: signs, huge values and the n4 == 0 case follow the reference."""
    z = np.load(os.path.join(golden_dir, "lpg_pole_r8.npz"))
    coef = torch.from_numpy(z["coef"]).to(DEV)
    full, _ = ops.lpg_forward(coef, 8)
    parity.check_forward(npf(full), z["coef"], 8, what="pole")
    ref64, den = c_oracle.lpg_forward_f64(z["coef"], 8, return_den=True)
    ok = np.abs(den) > 1e-3
    assert np.array_equal(np.sign(npf(full)[..., 0][ok]), np.sign(ref64[ok]))
    neg = den < -1e-3
    assert neg.any() and np.signbit(npf(full)[..., 0][neg]).all()          # -0.0 where n4 == 0, like the reference


# ------------------------------------------------------------------------------------------------
# seeded parity at assorted shapes: vector paths with every PX, ragged widths, generic path
# ------------------------------------------------------------------------------------------------
SHAPES = [
    (2, 15, 20),    # w % 4 == 0: widest vectors
    (1, 13, 17),    # NYU 416x544 at /32: odd width -> narrow vectors / generic
    (3, 7, 34),     # w % 2 == 0 only
    (2, 11, 38),    # KITTI 352x1216 at /32
    (1, 1, 1),
]


@pytest.mark.parametrize("r,d", RS)
@pytest.mark.parametrize("B,h,w", SHAPES)
@pytest.mark.parametrize("with_ds", [True, False])
def test_forward_backward_f32(B, h, w, r, d, with_ds):
    d = d if with_ds else 0
    coef, g_full, g_ds = make_inputs(B, h, w, r, d, seed=B * 100 + w)
    c, gf = coef.to(DEV), g_full.to(DEV)
    gd = g_ds.to(DEV) if d else None
    full, ds = ops.lpg_forward(c, r, d)
    kern_f = ops.last_kernel()
    parity.check_forward(npf(full), coef.numpy(), r, what="%s" % kern_f)
    if d:
        assert torch.equal(ds, full[:, ::d, ::d])
    gc = ops.lpg_backward(c, gf, gd, r, d)
    parity.check_backward(npf(gc), coef.numpy(), g_full.numpy(), r, g_ds.numpy() if d else None, d, what=ops.last_kernel())
    # only one of the two gradients present
    gc1 = ops.lpg_backward(c, gf, None, r, 0)
    parity.check_backward(npf(gc1), coef.numpy(), g_full.numpy(), r, what="g_full only")
    if d:
        gc2 = ops.lpg_backward(c, None, gd, r, d)
        parity.check_backward(npf(gc2), coef.numpy(), None, r, g_ds.numpy(), d, what="g_ds only")


def test_kernel_variant_selection():
    """The reference's shapes must hit the vectorised kernels, not the generic fallback."""
    for (B, H, W) in [(2, 480, 640), (1, 416, 544), (1, 352, 1216)]:
        for r, d in RS:
            coef, _, _ = make_inputs(B, H // r, W // r, r, d)
            ops.lpg_forward(coef.to(DEV), r, d)
            assert ops.last_kernel().startswith("lpg_fwd_vec<f32,r%d" % r), ops.last_kernel()
            ops.lpg_forward(coef.to(DEV).bfloat16(), r, d)
            assert ops.last_kernel().startswith("lpg_fwd_vec<bf16,r%d" % r), ops.last_kernel()


@pytest.mark.parametrize("r,d", RS)
def test_generic_path_bit_identical_to_vector_path(r, d):
    """Same arithmetic in both paths: force the generic kernel with a non-unit column stride."""
    coef, g_full, g_ds = make_inputs(2, 6, 8, r, d, seed=5)
    c = coef.to(DEV)
    full, ds = ops.lpg_forward(c, r, d)
    assert "vec" in ops.last_kernel()
    slot = torch.zeros(2, 6 * r, 8 * r, 3, device=DEV)          # NHWC concat buffer, LPG map = channel 1
    full2, ds2 = ops.lpg_forward(c, r, d, out_full=slot[..., 1:2])
    assert "generic" in ops.last_kernel()
    assert torch.equal(slot[..., 1:2], full) and (slot[..., 0] == 0).all() and (slot[..., 2] == 0).all()
    if d:
        assert torch.equal(ds2, ds)
    gslot = torch.randn(2, 6 * r, 8 * r, 3, device=DEV)
    gd = g_ds.to(DEV) if d else None
    ga = ops.lpg_backward(c, gslot[..., 1:2].contiguous(), gd, r, d)
    gb = ops.lpg_backward(c, gslot[..., 1:2], gd, r, d)
    assert "generic" in ops.last_kernel()
    # same per-pixel terms; only the association of the patch sum may differ (lane-group tree at r=8)
    scale, _ = parity.backward_scale(coef.numpy(), npf(gslot[..., 1:2]), r, g_ds.numpy() if d else None, d)
    assert (np.abs(npf(ga) - npf(gb)) <= 4e-7 * scale + 1e-30).all()


@pytest.mark.parametrize("r", [1, 3, 16])
def test_generic_upratio(r):
    """The reference layer accepts any upratio; only 2/4/8 are vectorised."""
    coef, g_full, _ = make_inputs(2, 3, 5, r, 0, seed=r)
    c = coef.to(DEV)
    full, _ = ops.lpg_forward(c, r)
    assert "generic" in ops.last_kernel()
    parity.check_forward(npf(full), coef.numpy(), r)
    gc = ops.lpg_backward(c, g_full.to(DEV), None, r)
    parity.check_backward(npf(gc), coef.numpy(), g_full.numpy(), r)


def test_planar_concat_slot_keeps_vector_path():
    """concat1 = [upconv1, d2, d4, d8] (bts_decoder.py:99) held channels-first: each LPG map is a
    contiguous plane of the buffer and is written in place by the vector kernels."""
    B, H, W, C = 2, 32, 64, 5
    buf = torch.zeros(B, C + 3, H, W, device=DEV)
    for k, (r, d) in enumerate([(2, 0), (4, 2), (8, 4)]):
        coef, _, _ = make_inputs(B, H // r, W // r, r, d, seed=k)
        plane = buf[:, C + k].unsqueeze(-1)                      # (B,H,W,1) view, strides (C+3)*H*W, W, 1, 1
        full, _ = ops.lpg_forward(coef.to(DEV), r, d, out_full=plane)
        assert "vec" in ops.last_kernel(), ops.last_kernel()
        ref, _ = ops.lpg_forward(coef.to(DEV), r, d)
        assert torch.equal(buf[:, C + k], ref[..., 0])
    assert (buf[:, :C] == 0).all()


@pytest.mark.parametrize("r,d", RS)
def test_bf16(r, d):
    coef, g_full, g_ds = make_inputs(2, 12, 16, r, d, seed=3, dtype=torch.bfloat16)
    c, gf = coef.to(DEV), g_full.to(DEV)
    gd = g_ds.to(DEV) if d else None
    full, ds = ops.lpg_forward(c, r, d)
    assert full.dtype == torch.bfloat16 and "bf16" in ops.last_kernel()
    parity.check_forward(npf(full), npf(coef), r, rtol=1e-2, what="bf16 fwd")
    # against the float32 kernel on the same (bf16-valued) inputs, rounded to bf16: at most 1 bf16 ulp apart
    f32, _ = ops.lpg_forward(c.float(), r, d)
    assert (npf(full) - npf(f32.bfloat16())).__abs__().max() <= 2 ** -7 * np.abs(npf(f32)).max()
    if d:
        assert torch.equal(ds, full[:, ::d, ::d])
    gc = ops.lpg_backward(c, gf, gd, r, d)
    parity.check_backward(npf(gc), npf(coef), npf(g_full), r, npf(g_ds) if d else None, d, rtol=1e-2, what="bf16 bwd")


# ------------------------------------------------------------------------------------------------
# multi-layer launches
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_multi_equals_single(dtype):
    B, H, W = 2, 64, 96
    layers_f, layers_b, singles = [], [], []
    for k, (r, d) in enumerate(RS):
        coef, g_full, g_ds = make_inputs(B, H // r, W // r, r, d, seed=40 + k, dtype=dtype)
        c, gf = coef.to(DEV), g_full.to(DEV)
        gd = g_ds.to(DEV) if d else None
        full, ds = ops.lpg_forward(c, r, d)
        gc = ops.lpg_backward(c, gf, gd, r, d)
        singles.append((full, ds, gc))
        layers_f.append(dict(coef=c, upratio=r, ds_stride=d, out_full=torch.empty_like(full), out_ds=torch.empty_like(ds) if d else None))
        layers_b.append(dict(coef=c, g_full=gf, g_ds=gd, upratio=r, ds_stride=d, g_coef=torch.empty_like(gc)))
    ops.reset_launch_count()
    ops.lpg_forward_multi(layers_f)
    assert ops.launch_count() == 1 and ops.last_kernel().startswith("lpg_fwd_multi"), ops.last_kernel()
    ops.lpg_backward_multi(layers_b)
    assert ops.launch_count() == 2 and ops.last_kernel().startswith("lpg_bwd_multi"), ops.last_kernel()
    for (full, ds, gc), lf, lb in zip(singles, layers_f, layers_b):
        assert torch.equal(lf["out_full"], full)
        if ds is not None:
            assert torch.equal(lf["out_ds"], ds)
        assert torch.equal(lb["g_coef"], gc)


def test_multi_falls_back_per_layer():
    coef, _, _ = make_inputs(1, 5, 7, 3, 0)
    c = coef.to(DEV)
    out = torch.empty(1, 15, 21, 1, device=DEV)
    ops.reset_launch_count()
    ops.lpg_forward_multi([dict(coef=c, upratio=3, ds_stride=0, out_full=out, out_ds=None)])
    assert ops.launch_count() == 1 and "generic" in ops.last_kernel()
    parity.check_forward(npf(out), coef.numpy(), 3)


# ------------------------------------------------------------------------------------------------
# autograd / layer surface
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("r,d", RS)
def test_layer_autograd_matches_oracle(r, d):
    coef, g_full, g_ds = make_inputs(2, 6, 10, r, d, seed=9)
    layer = pkg.LocalPlanarGuidance(upratio=r, ds_stride=d, name="depth_%dx%d_scaled" % (r, r))
    x = coef.to(DEV).requires_grad_(True)
    out = layer(x)
    if d:
        full, ds = out
        torch.autograd.backward([full, ds], [g_full.to(DEV), g_ds.to(DEV)])
    else:
        full = out
        full.backward(g_full.to(DEV))
    assert tuple(full.shape) == layer.compute_output_shape(tuple(coef.shape))
    parity.check_forward(npf(full), coef.numpy(), r)
    parity.check_backward(npf(x.grad), coef.numpy(), g_full.numpy(), r, g_ds.numpy() if d else None, d)
    with pytest.raises(ValueError, match="was built for"):
        layer(torch.rand(1, 3, 3, 3, device=DEV))


def test_layer_inside_torch_graph():
    """Gradient flows through the layer into an upstream op (sigmoid head in plain torch)."""
    torch.manual_seed(0)
    feat = torch.randn(2, 6, 8, 16, device=DEV)
    kern = (torch.randn(16, 3, device=DEV) * 0.3).requires_grad_(True)
    layer = pkg.LocalPlanarGuidance(4)
    depth = layer(torch.sigmoid(feat @ kern))
    loss = (depth * torch.linspace(0, 1, depth.numel(), device=DEV).reshape(depth.shape)).sum()
    loss.backward()
    x64 = torch.sigmoid(feat.double().cpu() @ kern.detach().double().cpu())
    g = torch.linspace(0, 1, depth.numel()).reshape(depth.shape).numpy()
    gcoef = c_oracle.lpg_backward_f64(x64.numpy(), g, 4)
    _, gw = c_oracle.head_backward_f64(feat.double().cpu().numpy(), kern.detach().double().cpu().numpy(), x64.numpy(), gcoef)
    assert np.abs(npf(kern.grad) - gw).max() <= 2e-5 * np.abs(gw).max()


# ------------------------------------------------------------------------------------------------
# full-size, size-independent properties (BASELINE config 2: B=32, 480x640)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("r,d", RS)
def test_full_size_properties(r, d):
    B, H, W = 32, 480, 640
    g = torch.Generator(device=DEV).manual_seed(0)
    coef = torch.sigmoid(torch.randn(B, H // r, W // r, 3, device=DEV, generator=g))
    full, ds = ops.lpg_forward(coef, r, d)
    assert ops.last_kernel().startswith("lpg_fwd_vec<f32,r%d" % r)
    # (1) ds is the strided slice of full
    if d:
        assert torch.equal(ds, full[:, ::d, ::d])
    # (2) linear in the distance channel: doubling n4 doubles the depth exactly (power of two)
    coef2 = coef.clone()
    coef2[..., 2] *= 2
    full2, _ = ops.lpg_forward(coef2, r, 0)
    assert torch.equal(full2, full * 2)
    # (3) determinism: bit-identical across runs
    again, _ = ops.lpg_forward(coef, r, d)
    assert torch.equal(again, full)
    # (4) a sample of the batch against the oracle
    parity.check_forward(npf(full[:2]), npf(coef[:2]), r)
    # (5) backward is linear in the upstream gradient and deterministic
    gf = torch.randn(B, H, W, 1, device=DEV, generator=g)
    gd = torch.randn(B, H // d, W // d, 1, device=DEV, generator=g) if d else None
    ga = ops.lpg_backward(coef, gf, gd, r, d)
    gb = ops.lpg_backward(coef, gf, gd, r, d)
    assert torch.equal(ga, gb)
    g2 = ops.lpg_backward(coef, gf * 2, gd * 2 if d else None, r, d)
    assert torch.equal(g2, ga * 2)
    # (6) superposition: grad(g_full) + grad(g_ds) == grad(both) up to float32 rounding of the sums
    if d:
        g_only_f = ops.lpg_backward(coef, gf, None, r, 0)
        g_only_d = ops.lpg_backward(coef, None, gd, r, d)
        scale, _ = parity.backward_scale(npf(coef[:1]), npf(gf[:1]), r, npf(gd[:1]), d)
        assert (np.abs(npf((g_only_f + g_only_d - ga)[:1])) <= 1e-5 * scale + 1e-30).all()
    parity.check_backward(npf(ga[:2]), npf(coef[:2]), npf(gf[:2]), r, npf(gd[:2]) if d else None, d)
    # (7) d(sum of depth)/d(dist) == sum over the patch of 1/den: check via the forward itself
    ones = torch.ones(B, H, W, 1, device=DEV)
    g1 = ops.lpg_backward(coef, ones, None, r, 0)
    unit = coef.clone()
    unit[..., 2] = 1.0
    inv_den, _ = ops.lpg_forward(unit, r, 0)
    patches = inv_den.reshape(B, H // r, r, W // r, r)
    patch_sum = patches.sum(dim=(2, 4))
    tame = patches.abs().amax(dim=(2, 4)) < 100.0               # skip the handful of patches that touch the pole
    torch.testing.assert_close(g1[..., 2][tame], patch_sum[tame], rtol=2e-5, atol=0)


@pytest.mark.parametrize("r,d", RS)
def test_full_size_pin(golden_dir, r, d):
    """Full-size run (2 x 480 x 640) against 4096 sampled outputs of the UNMODIFIED reference layer."""
    from test_oracle import full_size_inputs
    z = np.load(os.path.join(golden_dir, "lpg_full_size_samples.npz"))
    coef, idx = full_size_inputs(r)
    full, _ = ops.lpg_forward(torch.from_numpy(coef).to(DEV), r, d)
    assert ops.last_kernel().startswith("lpg_fwd_vec<f32,r%d" % r)
    _, den = c_oracle.lpg_forward_f64(coef, r, return_den=True)
    good = den.reshape(-1)[idx] >= parity.DEN_OK
    got = npf(full).reshape(-1)[idx]
    np.testing.assert_allclose(got[good], z["r%d_out64" % r][good], rtol=1e-5)
    np.testing.assert_allclose(got[good], z["r%d_out" % r][good], rtol=1e-5)     # the reference's own float32 run
    assert int((full < 0).sum()) == int(z["r%d_n_negative" % r])
    parity.check_forward(npf(full), coef, r, what="full-size pin r=%d" % r)


@pytest.mark.parametrize("H,W", [(352, 1216), (416, 544)])
@pytest.mark.parametrize("r,d", RS)
def test_full_size_pin_other_shapes(golden_dir, r, d, H, W):
    """The same pin at KITTI Eigen 352 x 1216 and the NYU training crop 416 x 544: forward against sampled outputs of the
    UNMODIFIED reference layer, backward against the float64 oracle."""
    from test_oracle import full_size_inputs
    z = np.load(os.path.join(golden_dir, "lpg_full_size_samples_shapes.npz"))
    key = "h%dw%d_r%d" % (H, W, r)
    coef, idx = full_size_inputs(r, 2, H, W)
    full, ds = ops.lpg_forward(torch.from_numpy(coef).to(DEV), r, d)
    assert ops.last_kernel().startswith("lpg_fwd_vec<f32,r%d" % r)
    _, den = c_oracle.lpg_forward_f64(coef, r, return_den=True)
    good = den.reshape(-1)[idx] >= parity.DEN_OK
    got = npf(full).reshape(-1)[idx]
    np.testing.assert_allclose(got[good], z[key + "_out64"][good], rtol=1e-5)
    np.testing.assert_allclose(got[good], z[key + "_out"][good], rtol=1e-5)
    assert int((full < 0).sum()) == int(z[key + "_n_negative"])
    if d:
        assert torch.equal(ds, full[:, ::d, ::d])
    g = torch.Generator().manual_seed(H + r)
    g_full = torch.randn(2, H, W, 1, generator=g)
    gc = ops.lpg_backward(torch.from_numpy(coef).to(DEV), g_full.to(DEV), None, r, 0)
    parity.check_backward(npf(gc), coef, g_full.numpy(), r, what="full-size bwd %dx%d r=%d" % (H, W, r))


# ------------------------------------------------------------------------------------------------
# errors and edge cases through the ABI
# ------------------------------------------------------------------------------------------------
def test_empty_batch():
    coef = torch.empty(0, 4, 4, 3, device=DEV)
    full, ds = ops.lpg_forward(coef, 8, 4)
    assert full.shape == (0, 32, 32, 1) and ds.shape == (0, 8, 8, 1)
    gc = ops.lpg_backward(coef, full, ds, 8, 4)
    assert gc.shape == coef.shape


def test_errors():
    coef = torch.rand(1, 4, 4, 3, device=DEV)
    with pytest.raises(ValueError, match="expected"):
        ops.lpg_forward(coef, 8, 0, out_full=torch.empty(1, 32, 31, 1, device=DEV))
    with pytest.raises(ValueError, match="dtype"):
        ops.lpg_forward(coef, 8, 0, out_full=torch.empty(1, 32, 32, 1, device=DEV, dtype=torch.bfloat16))
    with pytest.raises(ValueError, match="float32 or bfloat16"):
        ops.lpg_forward(coef.double(), 8)
    with pytest.raises(ValueError, match="divide"):
        ops.lpg_forward(coef, 8, 3, out_ds=torch.empty(1, 10, 10, 1, device=DEV))
    with pytest.raises(ValueError, match="no CPU fallback"):
        ops.lpg_forward(coef.cpu(), 8)


def test_non_default_stream_and_graph_capture():
    """Launches go to the caller's current stream and are capturable into a CUDA graph."""
    coef, g_full, g_ds = make_inputs(2, 8, 8, 8, 4, seed=2)
    c, gf, gd = coef.to(DEV), g_full.to(DEV), g_ds.to(DEV)
    ref_full, ref_ds = ops.lpg_forward(c, 8, 4)
    ref_g = ops.lpg_backward(c, gf, gd, 8, 4)
    full, ds, gc = torch.empty_like(ref_full), torch.empty_like(ref_ds), torch.empty_like(ref_g)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.lpg_forward(c, 8, 4, out_full=full, out_ds=ds)       # warm-up on the side stream
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            ops.lpg_forward(c, 8, 4, out_full=full, out_ds=ds)
            ops.lpg_backward(c, gf, gd, 8, 4, g_coef=gc)
        full.zero_(), ds.zero_(), gc.zero_()
        graph.replay()
    s.synchronize()
    assert torch.equal(full, ref_full) and torch.equal(ds, ref_ds) and torch.equal(gc, ref_g)


# ------------------------------------------------------------------------------------------------
# block-size knob: a patch split over LPP warps must keep its warp group inside one CTA
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("threads", [32, 64, 96, 192, 256, 1000])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_block_threads_knob_keeps_parity(threads, dtype):
    """btslpg_set_block_threads(t): the r = 8 kernels split a patch over 4 consecutive warps that exchange partial sums
    through shared memory indexed by the CTA-local warp id, so the CTA size is rounded to a multiple of 128 there
    (lpg_api.cu split_threads); every setting must give the same result as the default."""
    try:
        for r, d in RS:
            coef, g_full, g_ds = make_inputs(2, 9, 12, r, d, seed=r, dtype=dtype)
            c, gf, gd = coef.to(DEV), g_full.to(DEV), (g_ds.to(DEV) if d else None)
            ops.set_block_threads(0, 0)
            full0, _ = ops.lpg_forward(c, r, d)
            gc0 = ops.lpg_backward(c, gf, gd, r, d)
            ops.set_block_threads(threads, threads)
            full1, ds1 = ops.lpg_forward(c, r, d)
            gc1 = ops.lpg_backward(c, gf, gd, r, d)
            torch.cuda.synchronize()
            assert torch.equal(full0, full1) and torch.equal(gc0, gc1), (r, threads, ops.last_kernel())
            if dtype == torch.float32:
                parity.check_backward(npf(gc1), coef.numpy(), g_full.numpy(), r, g_ds.numpy() if d else None, d, what="threads=%d" % threads)
    finally:
        ops.set_block_threads(0, 0)
