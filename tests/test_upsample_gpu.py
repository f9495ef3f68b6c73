"""GPU tests of the nearest x2 up-sampling kernels (bts_decoder.py:31, :38, :97) through the C ABI:
bit-exact copies forward, fixed-order 4-term sums backward."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from bts_fully_tf_b200 import ops
from oracle import tail_oracle as T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,h,w,C", [(1, 1, 1, 4), (2, 3, 5, 64), (1, 7, 9, 128), (2, 4, 6, 2208), (1, 5, 3, 6), (3, 2, 2, 1), (1, 15, 20, 512)])
def test_upsample_matches_oracle_and_torch(B, h, w, C, dtype):
    g = torch.Generator().manual_seed(C + w)
    x = torch.randn(B, h, w, C, generator=g).to(dtype)
    xd = x.to(DEV).requires_grad_(True)
    y = ops.upsample2x_nhwc(xd)
    vec = (C * x.element_size()) % 16 == 0
    assert ops.last_kernel().startswith("upsample2x_fwd<" if vec else "upsample2x_fwd_generic<"), ops.last_kernel()
    ref = T.upsample2x(x.float().numpy())
    assert np.array_equal(y.detach().float().cpu().numpy(), ref)                       # a copy: bit-exact
    ref_t = F.interpolate(x.to(DEV).permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(y.detach(), ref_t)
    g_out = torch.randn(B, 2 * h, 2 * w, C, generator=g).to(dtype)
    y.backward(g_out.to(DEV))
    ref_g = T.upsample2x_grad(g_out.float().numpy())
    tol = 1e-6 if dtype == torch.float32 else 2 ** -7
    assert np.abs(xd.grad.float().cpu().numpy() - ref_g).max() <= tol * max(np.abs(ref_g).max(), 1e-30)
    # deterministic
    g2 = ops.upsample2x_backward(g_out.to(DEV))
    assert torch.equal(g2, xd.grad)


def test_upsample_full_size():
    """the decoder's largest: (8, 240, 320, 64) -> (8, 480, 640, 64)"""
    x = torch.randn(8, 240, 320, 64, device=DEV)
    y = ops.upsample2x_forward(x)
    assert torch.equal(y[:, ::2, ::2], x) and torch.equal(y[:, 1::2, ::2], x) and torch.equal(y[:, ::2, 1::2], x) and torch.equal(y[:, 1::2, 1::2], x)
    g = ops.upsample2x_backward(y)
    assert torch.equal(g, x * 4)                                                       # (x + x) + (x + x) is exact


def test_upsample_errors():
    with pytest.raises(ValueError, match="not a CUDA tensor"):
        ops.upsample2x_forward(torch.zeros(1, 2, 2, 4))
    with pytest.raises(ValueError, match="expected"):
        ops.upsample2x_forward(torch.zeros(1, 2, 2, 4, device=DEV), out=torch.zeros(1, 4, 5, 4, device=DEV))
