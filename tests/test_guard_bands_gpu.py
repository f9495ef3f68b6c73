"""Out-of-bounds detection without a sanitizer (compute-sanitizer is closed on the GPU pool): every tensor
of a call is carved from the middle of a larger buffer.  Input padding holds NaN -- an out-of-bounds READ
poisons a result, which the parity check then catches; output padding holds a sentinel -- an out-of-bounds
WRITE changes it.  Shapes are ragged on purpose (odd widths, pixel counts that are not multiples of the
vector widths / warp tiles / ring stages)."""
import numpy as np
import pytest
import torch

from bts_fully_tf_b200 import ops
from oracle import c_oracle, tail_oracle
import lpg_parity as parity

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PAD = 4096            # elements of padding on both sides (a multiple of every vector width -> alignment preserved)
SENTINEL = -12288.0      # exactly representable in bfloat16


class Arena:
    """Hands out tensors surrounded by guard bands and checks the bands afterwards."""

    def __init__(self):
        self.items = []

    def _carve(self, shape, dtype, fill):
        n = int(np.prod(shape))
        n_al = (n + 7) // 8 * 8
        buf = torch.full((PAD + n_al + PAD,), fill, dtype=dtype, device=DEV)
        view = buf[PAD:PAD + n].view(shape)
        self.items.append((buf, n, fill))
        return view

    def input(self, t):
        v = self._carve(tuple(t.shape), t.dtype, float("nan"))
        v.copy_(t.to(DEV))
        return v

    def output(self, shape, dtype=torch.float32):
        return self._carve(shape, dtype, SENTINEL)

    def check(self):
        torch.cuda.synchronize()
        for buf, n, fill in self.items:
            lo, hi = buf[:PAD], buf[PAD + (n + 7) // 8 * 8:]
            slack = buf[PAD + n:PAD + (n + 7) // 8 * 8]
            for band in (lo, hi, slack):
                if band.numel() == 0:
                    continue
                if fill != fill:
                    assert torch.isnan(band).all(), "input guard band was overwritten"
                else:
                    assert (band.float() == fill).all(), "write outside the output tensor"


def npf(t):
    return t.detach().float().cpu().numpy()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("r,d", [(8, 4), (4, 2), (2, 0)])
@pytest.mark.parametrize("B,h,w", [(1, 1, 1), (1, 3, 5), (2, 7, 9), (1, 5, 33), (3, 9, 34)])
def test_lpg_guard_bands(r, d, B, h, w, dtype):
    g = torch.Generator().manual_seed(h * 100 + w)
    coef = torch.sigmoid(torch.randn(B, h, w, 3, generator=g)).to(dtype)
    g_full = torch.randn(B, h * r, w * r, 1, generator=g).to(dtype)
    g_ds = torch.randn(B, h * r // d, w * r // d, 1, generator=g).to(dtype) if d else None
    rtol = 1e-5 if dtype == torch.float32 else 1e-2
    A = Arena()
    c = A.input(coef)
    full = A.output((B, h * r, w * r, 1), dtype)
    ds = A.output((B, h * r // d, w * r // d, 1), dtype) if d else None
    ops.lpg_forward(c, r, d, out_full=full, out_ds=ds)
    gc = A.output((B, h, w, 3), dtype)
    ops.lpg_backward(c, A.input(g_full), A.input(g_ds) if d else None, r, d, g_coef=gc)
    A.check()
    parity.check_forward(npf(full), npf(coef), r, rtol=rtol, what="guarded fwd")
    parity.check_backward(npf(gc), npf(coef), npf(g_full), r, npf(g_ds) if d else None, d, rtol=rtol, what="guarded bwd")
    if d:
        assert torch.equal(ds, full[:, ::d, ::d])


@pytest.mark.parametrize("B,h,w", [(1, 2, 3), (2, 5, 7)])
def test_lpg_multi_guard_bands(B, h, w):
    """the three scales of one decoder in one launch (static layer slots), ragged sizes"""
    g = torch.Generator().manual_seed(11)
    A = Arena()
    fw, bw, keep = [], [], []
    for r, d in ((8, 4), (4, 2), (2, 0)):
        hh, ww = h * (8 // r), w * (8 // r)
        coef = torch.sigmoid(torch.randn(B, hh, ww, 3, generator=g))
        g_full = torch.randn(B, hh * r, ww * r, 1, generator=g)
        g_ds = torch.randn(B, hh * r // d, ww * r // d, 1, generator=g) if d else None
        c = A.input(coef)
        full = A.output((B, hh * r, ww * r, 1))
        ds = A.output((B, hh * r // d, ww * r // d, 1)) if d else None
        gc = A.output((B, hh, ww, 3))
        fw.append(dict(coef=c, upratio=r, ds_stride=d, out_full=full, out_ds=ds))
        bw.append(dict(coef=c, g_full=A.input(g_full), g_ds=A.input(g_ds) if d else None, upratio=r, ds_stride=d, g_coef=gc))
        keep.append((r, d, coef, g_full, g_ds, full, gc))
    ops.lpg_forward_multi(fw)
    ops.lpg_backward_multi(bw)
    A.check()
    for r, d, coef, g_full, g_ds, full, gc in keep:
        parity.check_forward(npf(full), coef.numpy(), r, what="guarded multi fwd r=%d" % r)
        parity.check_backward(npf(gc), coef.numpy(), g_full.numpy(), r, g_ds.numpy() if d else None, d, what="guarded multi bwd r=%d" % r)


@pytest.mark.parametrize("r,d", [(8, 4), (4, 2), (2, 0)])
@pytest.mark.parametrize("C", [32, 64, 128])
@pytest.mark.parametrize("B,h,w", [(1, 1, 1), (1, 3, 11), (2, 5, 13)])
def test_head_guard_bands(r, d, C, B, h, w):
    """fused head: pixel counts that are not multiples of 32 (warp tile) or of the TMA ring stage"""
    g = torch.Generator().manual_seed(C + r + w)
    feat = torch.nn.functional.elu(torch.randn(B, h, w, C, generator=g))
    kern = (torch.rand(C, 3, generator=g) * 2 - 1) * (6.0 / (C + 3)) ** 0.5
    g_full = torch.randn(B, h * r, w * r, 1, generator=g)
    g_ds = torch.randn(B, h * r // d, w * r // d, 1, generator=g) if d else None
    A = Arena()
    f, k = A.input(feat), A.input(kern)
    coef = A.output((B, h, w, 3))
    full = A.output((B, h * r, w * r, 1))
    ds = A.output((B, h * r // d, w * r // d, 1)) if d else None
    ops.reduce_lpg_forward(f, k, r, d, out_full=full, out_ds=ds, coef_out=coef)
    assert ops.last_kernel().startswith("head_lpg_fwd")
    gk = A.output((C, 3))
    g_feat, g_kern, g_coef = ops.reduce_lpg_backward(f, k, coef, A.input(g_full), A.input(g_ds) if d else None, r, d,
                                                    g_kernel_out=gk, need_g_coef=True)
    A.check()
    ref = c_oracle.head_forward_f64(feat.numpy(), kern.numpy())
    np.testing.assert_allclose(npf(coef), ref, rtol=2e-6, atol=1e-7)
    parity.check_forward(npf(full), npf(coef), r, what="guarded head fwd")
    assert torch.isfinite(g_feat).all() and torch.isfinite(gk).all() and torch.isfinite(g_coef).all()
    parity.check_backward(npf(g_coef), npf(coef), g_full.numpy(), r, g_ds.numpy() if d else None, d, what="guarded head bwd")
    # head gradients from the kernel's own g_coef (float64 restatement of SURVEY 8(a) a8)
    x = npf(coef).astype(np.float64)
    dz = npf(g_coef).astype(np.float64) * x * (1 - x)
    ref_gf = dz.reshape(-1, 3) @ kern.numpy().astype(np.float64).T
    ref_gk = feat.numpy().astype(np.float64).reshape(-1, C).T @ dz.reshape(-1, 3)
    assert np.abs(npf(g_feat).reshape(-1, C) - ref_gf).max() <= 1e-5 * max(np.abs(ref_gf).max(), 1e-30)
    assert np.abs(npf(gk) - ref_gk).max() <= 2e-5 * max(np.abs(ref_gk).max(), 1e-30)


@pytest.mark.parametrize("n", [1, 3, 4, 7, 8, 1023, 1025, 65537])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_tail_guard_bands(n, dtype):
    g = torch.Generator().manual_seed(n)
    shape = (1, 1, n, 1)
    logit = (torch.randn(shape, generator=g) * 1.5).to(dtype)
    y_true = (torch.rand(shape, generator=g) * 10.5).to(dtype)
    if n > 4:
        y_true[0, 0, ::3] = 0
    A = Arena()
    z, yt = A.input(logit), A.input(y_true)
    depth = A.output(shape, dtype)
    ws = ops.tail_workspace(DEV)
    _, loss, _ = ops.silog_forward(z, yt, 10.0, 0.1, depth_est=depth, workspace=ws)
    gz = A.output(shape, dtype)
    ops.silog_backward(depth, yt, 10.0, 0.1, ws, None, True, g_out=gz)
    m = A.output((10,))
    ops.eval_metrics(yt, depth, 1e-3, 10.0, out=m)
    A.check()
    rtol = 1e-5 if dtype == torch.float32 else 1e-2
    ref_loss, (nv, _, var) = tail_oracle.si_log_loss(npf(y_true), npf(depth), 0.1)
    if nv >= 2 and var > 1e-6:
        np.testing.assert_allclose(float(loss), ref_loss, rtol=rtol)
        ref_g = tail_oracle.si_log_loss_grad(npf(y_true), npf(depth), 0.1, max_depth=10.0)
        assert np.abs(npf(gz) - ref_g).max() <= rtol * np.abs(ref_g).max()
    ref_m = tail_oracle.eval_metrics(npf(y_true), npf(depth), 1e-3, 10.0)
    assert int(m[9]) == ref_m["n_valid"]
    if ref_m["n_valid"]:
        np.testing.assert_allclose(float(m[3]), ref_m["rmse"], rtol=2e-5 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", [True, False])
@pytest.mark.parametrize("B,H,W,ca,cb,n_planes", [(1, 1, 1, 8, 0, 3), (1, 3, 5, 32, 0, 3), (2, 7, 9, 16, 24, 1), (1, 13, 31, 32, 0, 3)])
def test_concat_guard_bands(B, H, W, ca, cb, n_planes, act, dtype):
    g = torch.Generator().manual_seed(ca + W)
    a = (torch.randn(B, H, W, ca, generator=g) * 2).to(dtype)
    b = torch.randn(B, H, W, cb, generator=g).to(dtype) if cb else None
    planes = [torch.randn(B, H, W, 1, generator=g).to(dtype) for _ in range(n_planes)]
    g_out = torch.randn(B, H, W, ca + cb + n_planes, generator=g).to(dtype)
    A = Arena()
    out = A.output((B, H, W, ca + cb + n_planes), dtype)
    ops.concat_forward(A.input(a), [A.input(p) for p in planes], A.input(b) if cb else None, act, out=out)
    A.check()
    ref = tail_oracle.concat_elu(npf(a), [npf(p) for p in planes], None if b is None else npf(b), act)
    tol = 1e-6 if dtype == torch.float32 else 2 ** -8
    np.testing.assert_allclose(npf(out), ref, rtol=tol, atol=tol * 1e-2)
    B2 = Arena()
    g_a, g_b, g_p = ops.concat_backward(B2.input(g_out), B2.input(out) if act else None, act, ca, cb, n_planes)
    B2.check()
    ga, gb, gp = tail_oracle.concat_elu_grad(npf(g_out), npf(a), ca, cb, n_planes, act)
    assert np.abs(npf(g_a) - ga).max() <= (2e-6 if dtype == torch.float32 else 2 ** -6) * max(np.abs(ga).max(), 1e-30)
    for k in range(n_planes):
        np.testing.assert_array_equal(npf(g_p[k]), gp[k])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,h,w,C", [(1, 1, 1, 4), (2, 3, 5, 64), (1, 5, 3, 6), (1, 7, 9, 136)])
def test_upsample_guard_bands(B, h, w, C, dtype):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(B, h, w, C, generator=g).to(dtype)
    A = Arena()
    out = A.output((B, 2 * h, 2 * w, C), dtype)
    ops.upsample2x_forward(A.input(x), out=out)
    A.check()
    assert np.array_equal(npf(out), tail_oracle.upsample2x(npf(x)))
    B2 = Arena()
    g_in = B2.output((B, h, w, C), dtype)
    ops.upsample2x_backward(B2.input(out), g_in=g_in)
    B2.check()
    assert np.array_equal(npf(g_in), npf((x.float() * 4).to(dtype)))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,h,w,C", [(1, 1, 1, 8), (2, 3, 5, 384), (1, 5, 3, 10)])
def test_affine_act_guard_bands(B, h, w, C, dtype):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(B, h, w, C, generator=g).to(dtype)
    scale, shift = (torch.rand(C, generator=g) + 0.5), torch.randn(C, generator=g)
    A = Arena()
    out = A.output((B, h, w, C), dtype)
    ops.affine_act(A.input(x), dst=out, scale=A.input(scale), shift=A.input(shift), act=ops.ACT_RELU)
    A.check()
    tol = 2e-6 if dtype == torch.float32 else 2 ** -7
    np.testing.assert_allclose(npf(out), tail_oracle.affine_act(npf(x), npf(scale), npf(shift), 2), rtol=tol, atol=tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,C", [(1, 1, 1, 16), (2, 3, 5, 32), (1, 2, 131, 32), (1, 5, 7, 16)])
def test_depthconv_guard_bands(B, H, W, C, dtype):
    g = torch.Generator().manual_seed(W)
    x = torch.randn(B, H, W, C, generator=g).to(dtype)
    w = torch.randn(9 * C, generator=g) * 0.2
    go = torch.randn(B, H, W, 1, generator=g).to(dtype)
    A = Arena()
    g_x, g_k = ops.depthconv_backward(A.input(x), A.input(w), A.input(go))
    A.check()
    A2 = Arena()
    g_xe, g_ke = ops.depthconv_backward(A2.input(x), A2.input(w), A2.input(go), act_in=True)
    A2.check()
    ref_gxe, ref_gwe = tail_oracle.depth_tail_backward(npf(x), w.numpy(), npf(go))
    assert np.abs(npf(g_xe) - ref_gxe).max() <= (2e-6 if dtype == torch.float32 else 2 ** -7) * max(np.abs(ref_gxe).max(), 1e-30)
    assert np.abs(npf(g_ke).reshape(9, C) - ref_gwe).max() <= 1e-5 * max(np.abs(ref_gwe).max(), 1e-30)
    ref_gx, ref_gw = tail_oracle.depthconv_backward(npf(x), w.numpy(), npf(go))
    tol = 2e-6 if dtype == torch.float32 else 2 ** -7
    assert np.abs(npf(g_x) - ref_gx).max() <= tol * max(np.abs(ref_gx).max(), 1e-30)
    assert np.abs(npf(g_k).reshape(9, C) - ref_gw).max() <= 1e-5 * max(np.abs(ref_gw).max(), 1e-30)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,C", [(1, 1, 1, 16), (2, 3, 5, 32), (1, 2, 131, 32), (1, 33, 35, 16), (2, 31, 65, 32)])
def test_depthconv_forward_guard_bands(B, H, W, C, dtype):
    g = torch.Generator().manual_seed(W + C)
    x = torch.randn(B, H, W, C, generator=g).to(dtype)
    w = torch.randn(9 * C, generator=g) * 0.2
    A = Arena()
    out = A.output((B, H, W, 1), dtype)
    ops.depthconv_forward(A.input(x), A.input(w), act_in=True, sigmoid_scale=10.0, out=out)
    A.check()
    ref = tail_oracle.depth_tail_forward(npf(x), w.numpy(), act_in=True, max_depth=10.0)
    np.testing.assert_allclose(npf(out), ref, rtol=1e-5 if dtype == torch.float32 else 2 ** -7, atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,ca,cb,n_planes,pad", [(1, 1, 1, 8, 0, 3, 1), (1, 3, 5, 32, 0, 3, 1), (2, 7, 9, 16, 24, 1, 3), (1, 33, 37, 32, 8, 3, 5),
                                                      (1, 5, 3, 64, 32, 0, 0)])
def test_concat_chunked_guard_bands(B, H, W, ca, cb, n_planes, pad, dtype):
    """Shapes the chunked forward takes (float32: n_planes + pad == 4; bfloat16: == 8 or staged fallback), ragged pixel counts."""
    g = torch.Generator().manual_seed(ca + W + pad)
    a = (torch.randn(B, H, W, ca, generator=g) * 2).to(dtype)
    b = torch.randn(B, H, W, cb, generator=g).to(dtype) if cb else None
    planes = [torch.randn(B, H, W, 1, generator=g).to(dtype) for _ in range(n_planes)]
    scale = torch.rand(ca, generator=g) + 0.5
    shift = torch.randn(ca, generator=g)
    A = Arena()
    out = A.output((B, H, W, ca + cb + n_planes + pad), dtype)
    ops.set_tuning(10, 1)
    try:
        ops.concat_forward(A.input(a), [A.input(p) for p in planes], A.input(b) if cb else None, True, pad=pad, scale=A.input(scale), shift=A.input(shift), out=out)
    finally:
        ops.set_tuning(10, 0)
    A.check()
    V = 4 if dtype == torch.float32 else 8
    if ca % V == 0 and cb % V == 0 and n_planes + pad in (0, V):
        assert ops.last_kernel().startswith("concat_fwd_chunk<"), ops.last_kernel()
    ref = tail_oracle.concat_elu(npf(a), [npf(p) for p in planes], None if b is None else npf(b), True, pad=pad, scale=scale.numpy(), shift=shift.numpy())
    tol = 2e-6 if dtype == torch.float32 else 2 ** -7
    np.testing.assert_allclose(npf(out), ref, rtol=tol, atol=tol)
