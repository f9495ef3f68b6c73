"""GPU tests of the decoder's last convolution (bts_decoder.py:100-103) through the C ABI: the fused forward
(ELU of iconv1 + Conv2D(1, 3x3) + sigmoid * max_depth in one pass) and the fused backward (both gradients in one pass)."""
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


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C", [16, 32])
@pytest.mark.parametrize("B,H,W", [(1, 1, 1), (2, 3, 5), (1, 7, 37), (2, 5, 130), (1, 16, 128)])
def test_depthconv_backward_with_elu_vs_oracle(B, H, W, C, dtype):
    """act_in = 1: x is iconv1's pre-activation (bts_decoder.py:100); g_kernel against elu(x), g_x = d loss / d x."""
    g = torch.Generator().manual_seed(H * 1000 + W + C + 1)
    x = (torch.randn(B, H, W, C, generator=g) * 1.5).to(dtype)
    w9c = (torch.randn(9 * C, generator=g) * 0.2)
    g_out = torch.randn(B, H, W, 1, generator=g).to(dtype)
    g_x, g_k = ops.depthconv_backward(x.to(DEV), w9c.to(DEV), g_out.to(DEV), act_in=True)
    assert ops.last_kernel() == "depthconv_bwd<%s,C%d,elu>" % ("f32" if dtype == torch.float32 else "bf16", C)
    ref_gx, ref_gw = T.depth_tail_backward(npf(x), w9c.numpy(), npf(g_out))
    tol = 2e-6 if dtype == torch.float32 else 2 ** -7
    assert np.abs(npf(g_x) - ref_gx).max() <= tol * max(np.abs(ref_gx).max(), 1e-30)
    assert np.abs(npf(g_k).reshape(9, C) - ref_gw).max() <= 1e-5 * max(np.abs(ref_gw).max(), 1e-30)
    g_x2, g_k2 = ops.depthconv_backward(x.to(DEV), w9c.to(DEV), g_out.to(DEV), act_in=True)
    assert torch.equal(g_x2, g_x) and torch.equal(g_k2, g_k)                                     # bit-reproducible


@pytest.mark.parametrize("C", [16, 32])
def test_depth_conv_with_elu_autograd_matches_framework_ops(C):
    """ops.depth_conv(x, w, act_in=True) == conv2d(elu(x)) of the framework (TF32 off), values and both gradients."""
    torch.manual_seed(C + 7)
    B, H, W = 2, 24, 40
    x = torch.randn(B, H, W, C, device=DEV, requires_grad=True)
    conv = torch.nn.Conv2d(C, 1, 3, padding=1, bias=False).to(DEV)
    y = ops.depth_conv(x, conv.weight, act_in=True)
    g = torch.randn_like(y)
    y.backward(g)
    gx, gw = x.grad.clone(), conv.weight.grad.clone()
    x.grad = None
    conv.weight.grad = None
    y2 = conv(F.elu(x.permute(0, 3, 1, 2))).permute(0, 2, 3, 1)
    y2.backward(g)
    assert float((y.detach() - y2.detach()).abs().max()) <= 2e-6 * float(y2.detach().abs().max())
    assert float((gx - x.grad).abs().max()) <= 2e-6 * float(x.grad.abs().max())
    assert float((gw - conv.weight.grad).abs().max()) <= 1e-5 * float(conv.weight.grad.abs().max())


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
    assert float((y.detach() - y2.detach()).abs().max()) <= 2e-6 * float(y2.detach().abs().max())    # float32-accurate on both sides
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


# ---------------------------------------------------------------------------------------------
# forward: bts_decoder.py:100-103 in one pass
# ---------------------------------------------------------------------------------------------
FWD_SHAPES = [(1, 1, 1), (2, 3, 5), (1, 7, 37), (2, 33, 65), (1, 32, 32), (1, 31, 34), (1, 64, 96), (3, 40, 130)]


@pytest.fixture(params=["tensor", "fp32pipe"])
def impl(request):
    """Both phase-1 variants of the forward: tensor cores with the 3xTF32 split (default) and the FP32 pipe."""
    ops.set_tuning(9, 1 if request.param == "fp32pipe" else 0)
    yield request.param
    ops.set_tuning(9, 0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C", [16, 32])
@pytest.mark.parametrize("B,H,W", FWD_SHAPES)
def test_depthconv_forward_vs_oracle(B, H, W, C, dtype, impl):
    g = torch.Generator().manual_seed(H * 1000 + W + C)
    x = (torch.randn(B, H, W, C, generator=g) * 1.5).to(dtype)
    w9c = torch.randn(9 * C, generator=g) * 0.2
    name = "f32" if dtype == torch.float32 else "bf16"
    tol = 2e-6 if dtype == torch.float32 else 2 ** -7          # bfloat16: one rounding of the output
    for act_in in (False, True):
        ref = T.depth_tail_forward(npf(x), w9c.numpy(), act_in=act_in)
        y = ops.depthconv_forward(x.to(DEV), w9c.to(DEV), act_in=act_in)
        assert ops.last_kernel() == "depthconv_fwd%s<%s,C%d,%s>" % ("_fp32pipe" if impl == "fp32pipe" else "", name, C, "elu" if act_in else "lin")
        assert y.shape == (B, H, W, 1)
        assert np.abs(npf(y) - ref).max() <= tol * max(np.abs(ref).max(), 1e-30), (act_in, np.abs(npf(y) - ref).max())
        y2 = ops.depthconv_forward(x.to(DEV), w9c.to(DEV), act_in=act_in)
        assert torch.equal(y, y2)                                                        # fixed summation order
    ref = T.depth_tail_forward(npf(x), w9c.numpy(), act_in=True, max_depth=10.0)         # sigmoid * max_depth (NYU)
    d = ops.depthconv_forward(x.to(DEV), w9c.to(DEV), act_in=True, sigmoid_scale=10.0)
    np.testing.assert_allclose(npf(d), ref, rtol=1e-5 if dtype == torch.float32 else 2 ** -7, atol=1e-6)


def test_depthconv_forward_elu_extremes(impl):
    """ELU inside the sum at its corners: 0, tiny negatives (expm1 regime), large negatives (-> -1), large positives."""
    C = 32
    vals = torch.tensor([0.0, -0.0, -1e-8, -1e-4, -0.3, -0.35, -1.0, -20.0, -200.0, 1e-8, 3.0, 50.0])
    x = vals.repeat(C * 9 * 4)[:4 * 9 * C].view(1, 4, 9, C).contiguous()
    w9c = torch.linspace(-0.3, 0.3, 9 * C)
    ref = T.depth_tail_forward(npf(x), w9c.numpy(), act_in=True)
    y = ops.depthconv_forward(x.to(DEV), w9c.to(DEV), act_in=True)
    assert np.abs(npf(y) - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("C", [16, 32])
def test_depthconv_forward_matches_library_path(C):
    """The op-by-op library path the kernel replaces: F.elu -> conv2d (TF32 off) -> sigmoid * max_depth."""
    torch.manual_seed(C)
    B, H, W = 2, 96, 160
    x = torch.randn(B, H, W, C, device=DEV)
    conv = torch.nn.Conv2d(C, 1, 3, padding=1, bias=False).to(DEV)
    with torch.no_grad():
        ref = torch.sigmoid(conv(F.elu(x.permute(0, 3, 1, 2)))).permute(0, 2, 3, 1) * 80.0
        d = ops.depthconv_forward(x, ops.kernel9c(conv.weight), act_in=True, sigmoid_scale=80.0)
    assert float((d - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_depthconv_forward_full_size_properties(impl):
    """B=8 480x640 C=32 (the oracle takes minutes there): linear in the kernel and in x without activations, tap
    selection with one-hot kernels (exact on the FP32 pipe, to the 3xTF32 split's 2^-20 on the tensor cores),
    deterministic, and equal to the library convolution."""
    B, H, W, C = 8, 480, 640, 32
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(B, H, W, C, device=DEV, generator=g)
    w = torch.randn(9 * C, device=DEV, generator=g) * 0.1
    y1 = ops.depthconv_forward(x, w)
    assert torch.equal(ops.depthconv_forward(x, w * 2), y1 * 2) and torch.equal(ops.depthconv_forward(x * 2, w), y1 * 2)
    assert torch.equal(ops.depthconv_forward(x, w), y1)
    for tap, c in ((0, 0), (4, 17), (8, 31), (2, 5), (6, 30)):
        onehot = torch.zeros(9 * C, device=DEV)
        onehot[tap * C + c] = 1.0
        dy, dx = tap // 3 - 1, tap % 3 - 1
        ref = torch.zeros(B, H, W, device=DEV)
        ys, xs = slice(max(0, -dy), H - max(0, dy)), slice(max(0, -dx), W - max(0, dx))
        yd, xd = slice(max(0, dy), H - max(0, -dy)), slice(max(0, dx), W - max(0, -dx))
        ref[:, ys, xs] = x[:, yd, xd, c]
        got = ops.depthconv_forward(x, onehot)[..., 0]
        if impl == "fp32pipe":
            assert torch.equal(got, ref), (tap, c)
        else:
            assert float((got - ref).abs().max()) <= 2 ** -20 * float(ref.abs().max()), (tap, c)
    lib = F.conv2d(F.elu(x.permute(0, 3, 1, 2)), w.view(3, 3, C, 1).permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    y2 = ops.depthconv_forward(x, w, act_in=True)
    assert float((y2 - lib).abs().max()) <= 1e-5 * float(lib.abs().max())


def test_depthconv_forward_errors():
    x = torch.zeros(1, 4, 4, 32, device=DEV)
    with pytest.raises(ValueError, match="built for C = 16 and C = 32"):
        ops.depthconv_forward(torch.zeros(1, 4, 4, 8, device=DEV), torch.zeros(72, device=DEV))
    with pytest.raises(ValueError, match="at least"):
        ops.depthconv_forward(x, torch.zeros(100, device=DEV))
    with pytest.raises(ValueError, match="differs from x"):
        ops.depthconv_forward(x, torch.zeros(288, device=DEV), out=torch.zeros(1, 4, 5, 1, device=DEV))
    with pytest.raises(ValueError, match="contiguous"):
        ops.depthconv_forward(torch.zeros(1, 4, 4, 64, device=DEV)[..., :32], torch.zeros(288, device=DEV))
