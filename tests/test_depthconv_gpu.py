"""GPU tests of the fused backward of the decoder's last convolution (bts_decoder.py:102) through the C ABI."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from bts_fully_tf_b200 import ops
from oracle import tail_oracle as T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _no_tf32():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def npf(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C", [16, 32])
@pytest.mark.parametrize("B,H,W", [(1, 1, 1), (2, 3, 5), (1, 7, 37), (2, 5, 130), (1, 4, 257), (1, 16, 128)])
def test_depthconv_backward_vs_oracle(B, H, W, C, dtype):
    g = torch.Generator().manual_seed(H * 1000 + W + C)
    x = torch.randn(B, H, W, C, generator=g).to(dtype)
    w9c = (torch.randn(9 * C, generator=g) * 0.2)
    g_out = torch.randn(B, H, W, 1, generator=g).to(dtype)
    g_x, g_k = ops.depthconv_backward(x.to(DEV), w9c.to(DEV), g_out.to(DEV))
    assert ops.last_kernel() == "depthconv_bwd<%s,C%d>" % ("f32" if dtype == torch.float32 else "bf16", C)
    ref_gx, ref_gw = T.depthconv_backward(npf(x), w9c.numpy(), npf(g_out))
    tol = 2e-6 if dtype == torch.float32 else 2 ** -7
    assert np.abs(npf(g_x) - ref_gx).max() <= tol * max(np.abs(ref_gx).max(), 1e-30)
    assert np.abs(npf(g_k).reshape(9, C) - ref_gw).max() <= 1e-5 * max(np.abs(ref_gw).max(), 1e-30)   # float32 sums of B*H*W terms
    # only one of the two outputs
    g_x2, none = ops.depthconv_backward(x.to(DEV), w9c.to(DEV), g_out.to(DEV), need_g_kernel=False)
    assert none is None and torch.equal(g_x2, g_x)
    none, g_k2 = ops.depthconv_backward(x.to(DEV), w9c.to(DEV), g_out.to(DEV), need_g_x=False)
    assert none is None and torch.equal(g_k2, g_k)                                               # bit-reproducible reduction


@pytest.mark.parametrize("C", [16, 32])
def test_depth_conv_autograd_matches_library_convolution(C):
    torch.manual_seed(C)
    B, H, W = 2, 24, 40
    x = torch.randn(B, H, W, C, device=DEV, requires_grad=True)
    conv = torch.nn.Conv2d(C, 1, 3, padding=1, bias=False).to(DEV)
    y = ops.depth_conv(x, conv.weight)
    g = torch.randn_like(y)
    y.backward(g)
    gx, gw = x.grad.clone(), conv.weight.grad.clone()
    x.grad = None
    conv.weight.grad = None
    y2 = conv(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    y2.backward(g)
    assert torch.equal(y, y2)
    assert float((gx - x.grad).abs().max()) <= 2e-6 * float(x.grad.abs().max())
    assert float((gw - conv.weight.grad).abs().max()) <= 1e-5 * float(conv.weight.grad.abs().max())
    # forward against the oracle too (Keras HWIO (3,3,C,1) == [tap][c])
    w9c = conv.weight.detach().permute(2, 3, 1, 0).reshape(9, C)
    np.testing.assert_allclose(npf(y), T.depthconv_forward(npf(x), npf(w9c)), rtol=1e-5, atol=1e-6)


def test_depthconv_full_size_linearity():
    """B=8 480x640 C=32: linear in g_out and deterministic; sum over channels/pixels identity for a constant kernel."""
    B, H, W, C = 8, 480, 640, 32
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(B, H, W, C, device=DEV, generator=g)
    w = torch.randn(9 * C, device=DEV, generator=g) * 0.1
    go = torch.randn(B, H, W, 1, device=DEV, generator=g)
    gx1, gk1 = ops.depthconv_backward(x, w, go)
    gx2, gk2 = ops.depthconv_backward(x, w, go * 2)
    assert torch.equal(gx2, gx1 * 2) and torch.equal(gk2, gk1 * 2)
    gx3, gk3 = ops.depthconv_backward(x, w, go)
    assert torch.equal(gx3, gx1) and torch.equal(gk3, gk1)
    # centre tap of g_kernel == sum_p g[p] * x[p][c]
    ref = (go.double() * x.double()).sum(dim=(0, 1, 2))
    np.testing.assert_allclose(gk1.view(9, C)[4].double().cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=1e-3)


def test_depthconv_errors():
    x = torch.zeros(1, 4, 4, 32, device=DEV)
    with pytest.raises(ValueError, match="built for C = 16 and C = 32"):
        ops.depthconv_backward(torch.zeros(1, 4, 4, 8, device=DEV), torch.zeros(72, device=DEV), torch.zeros(1, 4, 4, 1, device=DEV))
    with pytest.raises(ValueError, match="differs from x"):
        ops.depthconv_backward(x, torch.zeros(288, device=DEV), torch.zeros(1, 4, 5, 1, device=DEV))
    with pytest.raises(ValueError, match="at least"):
        ops.depthconv_backward(x, torch.zeros(100, device=DEV), torch.zeros(1, 4, 4, 1, device=DEV))
