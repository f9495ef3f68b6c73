"""GPU tests of the strided per-channel affine + activation copy (DenseASPP glue) through the C ABI."""
import numpy as np
import pytest
import torch

from bts_fully_tf_b200 import ops
from oracle import tail_oracle as T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def npf(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", [ops.ACT_NONE, ops.ACT_ELU, ops.ACT_RELU])
@pytest.mark.parametrize("B,h,w,ctot,c0,C", [(2, 6, 8, 896, 0, 384), (1, 5, 7, 896, 256, 128), (1, 3, 3, 40, 8, 24), (2, 2, 5, 37, 3, 10), (1, 1, 1, 16, 0, 16)])
def test_gather_and_scatter_of_channel_slices(B, h, w, ctot, c0, C, act, dtype):
    g = torch.Generator().manual_seed(ctot + C)
    buf = torch.randn(B, h, w, ctot, generator=g).to(dtype).to(DEV)
    scale = (torch.rand(C, generator=g) + 0.5).to(DEV)
    shift = torch.randn(C, generator=g).to(DEV)
    tol = 2e-6 if dtype == torch.float32 else 2 ** -7
    # gather: strided slice -> contiguous, with the affine
    out = ops.affine_act(buf[..., c0:c0 + C], scale=scale, shift=shift, act=act)
    assert out.is_contiguous() and ops.last_kernel().startswith("affine_act_")
    ref = T.affine_act(npf(buf[..., c0:c0 + C]), npf(scale), npf(shift), act)
    np.testing.assert_allclose(npf(out), ref, rtol=tol, atol=tol)
    # scatter: contiguous -> slice of a wider buffer, plain copy: bit-exact, neighbours untouched
    dst = torch.full((B, h, w, ctot), 7.0, dtype=dtype, device=DEV)
    ops.affine_act(out, dst=dst[..., c0:c0 + C])
    assert torch.equal(dst[..., c0:c0 + C], out)
    assert (dst[..., :c0] == 7).all() and (dst[..., c0 + C:] == 7).all()
    # in place on a slice
    before = buf.clone()
    ops.affine_act(buf[..., c0:c0 + C], dst=buf[..., c0:c0 + C], scale=scale, shift=shift)
    np.testing.assert_allclose(npf(buf[..., c0:c0 + C]), T.affine_act(npf(before[..., c0:c0 + C]), npf(scale), npf(shift), 0), rtol=tol, atol=tol)
    assert torch.equal(buf[..., :c0], before[..., :c0]) and torch.equal(buf[..., c0 + C:], before[..., c0 + C:])


def test_affine_act_errors():
    x = torch.zeros(1, 2, 2, 8, device=DEV)
    with pytest.raises(ValueError, match="not a CUDA tensor"):
        ops.affine_act(torch.zeros(1, 2, 2, 8))
    with pytest.raises(ValueError, match="shape differs"):
        ops.affine_act(x, dst=torch.zeros(1, 2, 2, 4, device=DEV))
    with pytest.raises(ValueError, match="channel stride"):
        ops.affine_act(x.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1))
    with pytest.raises(ValueError, match="together"):
        ops.affine_act(x, scale=torch.ones(8, device=DEV))
