"""GPU tests of the fused activation + NHWC concat (bts_decoder.py:98-99 and :42) through the
C ABI, against oracle/tail_oracle.py and against the torch ops it replaces.  Copies are bit-exact; the
activation is within float32 rounding of expm1."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from bts_fully_tf_b200 import ops
from oracle import tail_oracle as T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def npf(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


@pytest.fixture(autouse=True, params=["chunked", "staged"])
def concat_impl(request):
    """Every test runs with both forward variants: the shared-memory-staged kernel (default) and the chunked kernel
    (tuning key 10; where each 16-byte output chunk has one source -- other shapes fall back by themselves)."""
    ops.set_tuning(10, 1 if request.param == "chunked" else 0)
    yield request.param
    ops.set_tuning(10, 0)


def make(B, H, W, ca, cb, n_planes, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    a = (torch.randn(B, H, W, ca, generator=g) * 2).to(dtype)
    b = torch.randn(B, H, W, cb, generator=g).to(dtype) if cb else None
    planes = [torch.randn(B, H, W, 1, generator=g).to(dtype) for _ in range(n_planes)]
    g_out = torch.randn(B, H, W, ca + cb + n_planes, generator=g).to(dtype)
    return a, b, planes, g_out


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", [True, False])
@pytest.mark.parametrize("B,H,W,ca,cb,n_planes", [
    (2, 16, 24, 32, 0, 3),      # concat1, densenet161 (F/16 = 32): 35 channels
    (1, 8, 16, 16, 0, 3),       # concat1, resnet50 (F/16 = 16): 19 channels
    (1, 3, 5, 32, 0, 3),        # 15 pixels: ragged tail only
    (2, 7, 9, 16, 24, 1),       # conv_block concat [upconv, skip, lpg_ds], 126 pixels
    (1, 20, 30, 64, 96, 1),     # 161 channels: small tiles
    (1, 4, 8, 8, 0, 0),         # nothing to append
    (1, 8, 8, 2, 0, 3),         # F/16 = 2 (the reference-decoder fixture, num_filters = 32): scalar path
    (2, 5, 7, 6, 3, 2),         # odd channel counts everywhere
])
def test_concat_matches_oracle_and_torch(B, H, W, ca, cb, n_planes, act, dtype, concat_impl):
    a, b, planes, g_out = make(B, H, W, ca, cb, n_planes, seed=ca + cb + W, dtype=dtype)
    ad = a.to(DEV).requires_grad_(True)
    bd = b.to(DEV).requires_grad_(True) if b is not None else None
    pd = [p.to(DEV).requires_grad_(True) for p in planes]
    out = ops.concat_nhwc(ad, pd, b=bd, act=act)
    tag = "<%s,%s," % ("f32" if dtype == torch.float32 else "bf16", "elu" if act else "id")
    assert ops.last_kernel().startswith(("concat_fwd" + tag, "concat_fwd_chunk" + tag)), ops.last_kernel()
    V = 4 if dtype == torch.float32 else 8
    if ca % V == 0 and cb % V == 0 and n_planes == 0:
        assert ops.last_kernel().startswith("concat_fwd_chunk" if concat_impl == "chunked" else "concat_fwd<"), ops.last_kernel()
    ref = T.concat_elu(npf(a), [npf(p) for p in planes], None if b is None else npf(b), act)
    tol = 1e-6 if dtype == torch.float32 else 2 ** -8
    np.testing.assert_allclose(npf(out), ref, rtol=tol, atol=tol * 1e-2)
    # copied channels are bit-exact
    assert torch.equal(out[..., ca + cb:], torch.cat([p.to(DEV) for p in planes], 3)) if n_planes else True
    if b is not None:
        assert torch.equal(out[..., ca:ca + cb], b.to(DEV))
    if not act:
        assert torch.equal(out[..., :ca], a.to(DEV))
    # gradients vs the oracle and vs torch autograd of the unfused ops
    out.backward(g_out.to(DEV))
    ga, gb, gp = T.concat_elu_grad(npf(g_out), npf(a), ca, cb, n_planes, act)
    gtol = 2e-6 if dtype == torch.float32 else 2 ** -7
    assert np.abs(npf(ad.grad) - ga).max() <= gtol * max(np.abs(ga).max(), 1e-30)
    if b is not None:
        np.testing.assert_array_equal(npf(bd.grad), gb)
    for k in range(n_planes):
        np.testing.assert_array_equal(npf(pd[k].grad), gp[k])
    a2 = a.to(DEV).requires_grad_(True)
    parts = [F.elu(a2) if act else a2] + ([b.to(DEV)] if b is not None else []) + [p.to(DEV) for p in planes]
    out2 = torch.cat(parts, 3)
    out2.backward(g_out.to(DEV))
    np.testing.assert_allclose(npf(out), npf(out2), rtol=tol, atol=tol * 1e-2)
    assert np.abs(npf(ad.grad) - npf(a2.grad)).max() <= gtol * max(np.abs(npf(a2.grad)).max(), 1e-30)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,ca,cb,n_planes,pad", [
    (2, 16, 24, 32, 0, 3, 1),       # concat1: 35 -> 36
    (2, 7, 9, 64, 96, 1, 3),        # block2 concat: 161 -> 164
    (1, 6, 10, 128, 96, 1, 3),      # block3 concat: 225 -> 228
    (1, 3, 5, 512, 384, 0, 0),      # block5 concat [up, skip]: 896 channels, 8-pixel tiles
    (1, 5, 7, 6, 3, 2, 1),          # scalar path with padding
    (1, 33, 37, 32, 8, 3, 5),       # 1221 pixels (more than one 1024-pixel work item of the chunked kernel); bfloat16: 3 + 5 = one 8-element chunk
    (3, 9, 11, 16, 0, 1, 3),        # single plane + 3 zero channels
])
def test_concat_pad_and_folded_batchnorm(B, H, W, ca, cb, n_planes, pad, dtype, concat_impl):
    a, b, planes, g_out = make(B, H, W, ca, cb, n_planes, seed=ca + pad, dtype=dtype)
    g = torch.Generator().manual_seed(7)
    scale = torch.rand(ca, generator=g) + 0.5
    shift = torch.randn(ca, generator=g)
    bd = b.to(DEV) if b is not None else None
    out = ops.concat_forward(a.to(DEV), [p.to(DEV) for p in planes], bd, act=True, pad=pad, scale=scale.to(DEV), shift=shift.to(DEV))
    assert "elu+affine" in ops.last_kernel() and ops.last_kernel().endswith("+%d>" % pad), ops.last_kernel()
    V = 4 if dtype == torch.float32 else 8
    if concat_impl == "chunked" and ca % V == 0 and cb % V == 0 and n_planes + pad in (0, V):
        assert ops.last_kernel().startswith("concat_fwd_chunk<"), ops.last_kernel()
    ref = T.concat_elu(npf(a), [npf(p) for p in planes], None if b is None else npf(b), True, pad=pad, scale=scale.numpy(), shift=shift.numpy())
    tol = 2e-6 if dtype == torch.float32 else 2 ** -7
    np.testing.assert_allclose(npf(out), ref, rtol=tol, atol=tol)
    if pad:
        assert (out[..., ca + cb + n_planes:] == 0).all()
    # autograd form with padding: the gradient of the pad channels is dropped, everything else as before
    ad = a.to(DEV).requires_grad_(True)
    pd = [p.to(DEV).requires_grad_(True) for p in planes]
    out2 = ops.concat_nhwc(ad, pd, b=bd, act=True, pad=pad)
    g_pad = torch.randn(B, H, W, ca + cb + n_planes + pad, generator=g).to(dtype)
    out2.backward(g_pad.to(DEV))
    ga, _, gp = T.concat_elu_grad(npf(g_pad), npf(a), ca, cb, n_planes, True)
    gtol = 2e-6 if dtype == torch.float32 else 2 ** -6
    assert np.abs(npf(ad.grad) - ga).max() <= gtol * max(np.abs(ga).max(), 1e-30)
    for k in range(n_planes):
        np.testing.assert_array_equal(npf(pd[k].grad), gp[k])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,h,w,cin,cout,pad", [(2, 4, 6, 8, 8, 1), (1, 5, 7, 16, 32, 1), (1, 3, 3, 4, 16, 0)])
def test_subpixel_source_equals_upsample_then_conv(B, h, w, cin, cout, pad, dtype):
    """UpSampling2D(2) + Conv2D(3x3) + ELU + concat  ==  low-res conv with ops.subpixel_kernel + concat(a_subpixel=True),
    forward and every gradient (bts_decoder.py:97-99)."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        g = torch.Generator().manual_seed(cin + cout)
        x = torch.randn(B, cin, h, w, generator=g).to(DEV)
        wt = (torch.randn(cout, cin, 3, 3, generator=g) * 0.2).to(DEV)
        planes = [torch.randn(B, 2 * h, 2 * w, 1, generator=g).to(dtype).to(DEV) for _ in range(3)]
        g_out = torch.randn(B, 2 * h, 2 * w, cout + 3 + pad, generator=g).to(dtype).to(DEV)
        # reference form (float32 convs, the concat in `dtype`)
        x1, w1 = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
        up = F.conv2d(F.interpolate(x1, scale_factor=2, mode="nearest"), w1, padding=1).permute(0, 2, 3, 1).to(dtype)
        ref = ops.concat_nhwc(up, planes, act=True, pad=pad)
        ref.backward(g_out)
        # sub-pixel form
        x2, w2 = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
        up4 = F.conv2d(x2, ops.subpixel_kernel(w2), padding=1).permute(0, 2, 3, 1).to(dtype)
        out = ops.concat_nhwc(up4, planes, act=True, pad=pad, a_subpixel=True)
        assert "+subpixel" in ops.last_kernel()
        out.backward(g_out)
        tol = 2e-5 if dtype == torch.float32 else 2e-2
        assert float((out - ref).abs().max()) <= tol * float(ref.abs().max())
        assert float((x2.grad - x1.grad).abs().max()) <= tol * float(x1.grad.abs().max())
        assert float((w2.grad - w1.grad).abs().max()) <= tol * float(w1.grad.abs().max())
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_concat_full_size_properties():
    """B=8 480x640 (1/4 of BASELINE config 2): channel-slot identities at a size the CPU oracle does not need to see."""
    B, H, W, ca = 8, 480, 640, 32
    g = torch.Generator(device=DEV).manual_seed(0)
    a = torch.randn(B, H, W, ca, generator=g, device=DEV)
    planes = [torch.randn(B, H, W, 1, generator=g, device=DEV) for _ in range(3)]
    out = ops.concat_forward(a, planes, act=True)
    assert torch.equal(out[..., 32], planes[0][..., 0]) and torch.equal(out[..., 33], planes[1][..., 0]) and torch.equal(out[..., 34], planes[2][..., 0])
    ref = F.elu(a)
    assert float((out[..., :32] - ref).abs().max()) <= 2e-7 * float(ref.abs().max())
    g_out = torch.randn_like(out)
    g_a, _, g_p = ops.concat_backward(g_out, out, True, ca, 0, 3)
    for k in range(3):
        assert torch.equal(g_p[k][..., 0], g_out[..., 32 + k])
    exp = g_out[..., :32] * torch.where(a > 0, torch.ones_like(a), torch.exp(a))
    assert float((g_a - exp).abs().max()) <= 1e-6 * float(exp.abs().max())


def test_concat_errors():
    a, b, planes, _ = make(1, 4, 8, 8, 0, 1, seed=0)
    with pytest.raises(ValueError, match="not a CUDA tensor"):
        ops.concat_forward(a, planes)
    with pytest.raises(ValueError, match="last dimension must be"):
        ops.concat_forward(a.to(DEV), [planes[0].to(DEV)], out=torch.empty(1, 4, 8, 12, device=DEV))
    with pytest.raises(ValueError, match="differs"):
        ops.concat_forward(a.to(DEV), [torch.zeros(1, 4, 4, 1, device=DEV)])
