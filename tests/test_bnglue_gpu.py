"""GPU parity of the TRAINING-mode conv block glue (ELU + BatchNormalization with batch statistics + concat; bts_decoder.py:32-42):
forward against the float64 oracle, forward and backward against torch autograd of the same three framework ops, moving averages,
determinism."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from bts_fully_tf_b200 import ops
from oracle import tail_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(B, H, W, C, CB, n_planes, seed):
    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(B, H, W, C, generator=g) * 1.5 + 0.2
    skip = torch.randn(B, H, W, CB, generator=g)
    planes = [torch.rand(B, H, W, 1, generator=g) for _ in range(n_planes)]
    bn = torch.nn.BatchNorm2d(C, eps=1.1e-5, momentum=0.01)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(C, generator=g) * 0.3)
        bn.running_mean.copy_(torch.randn(C, generator=g))
        bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    return raw, skip, planes, bn


@pytest.mark.parametrize("B,H,W,C,CB,n_planes", [(2, 9, 13, 32, 8, 1), (1, 5, 7, 8, 4, 0), (3, 16, 20, 64, 96, 1), (2, 6, 10, 256, 192, 0),
                                                (1, 4, 6, 512, 384, 0), (2, 33, 41, 16, 7, 1)])
def test_glue_matches_oracle_and_autograd(B, H, W, C, CB, n_planes):
    raw, skip, planes, bn = _case(B, H, W, C, CB, n_planes, seed=C + H)
    pad = ops.pad_to(C + CB + n_planes)
    bn_ref = torch.nn.BatchNorm2d(C, eps=1.1e-5, momentum=0.01)
    bn_ref.load_state_dict(bn.state_dict())
    bn, bn_ref = bn.to(DEV).train(), bn_ref.to(DEV).train()
    r1 = raw.to(DEV).requires_grad_(True)
    s1 = skip.to(DEV).requires_grad_(True)
    p1 = [p.to(DEV).requires_grad_(True) for p in planes]
    out = ops.conv_block_glue(r1, s1, p1, bn, pad=pad)
    ref, mean, var = tail_oracle.conv_block_glue(raw.numpy(), skip.numpy(), [p.numpy() for p in planes], bn_ref.weight.detach().cpu().numpy(),
                                                 bn_ref.bias.detach().cpu().numpy(), 1.1e-5, pad=pad)
    # normalised values are differences of O(1) numbers divided by a standard deviation: 2e-5 of the largest entry
    assert np.abs(out.detach().cpu().numpy() - ref).max() <= 2e-5 * np.abs(ref).max()
    # the framework's three ops with autograd
    r2 = raw.to(DEV).requires_grad_(True)
    s2 = skip.to(DEV).requires_grad_(True)
    p2 = [p.to(DEV).requires_grad_(True) for p in planes]
    up = bn_ref(F.elu(r2.permute(0, 3, 1, 2))).permute(0, 2, 3, 1)
    cat = torch.cat([up, s2] + p2 + ([torch.zeros(B, H, W, pad, device=DEV)] if pad else []), 3)
    g_out = torch.randn(cat.shape, generator=torch.Generator().manual_seed(1)).to(DEV)
    cat.backward(g_out)
    out.backward(g_out)
    torch.cuda.synchronize()
    tol = lambda t: 3e-5 * float(t.abs().max()) + 1e-7          # noqa: E731
    assert float((r1.grad - r2.grad).abs().max()) <= tol(r2.grad)
    assert torch.equal(s1.grad, s2.grad)
    for a, b in zip(p1, p2):
        assert torch.equal(a.grad, b.grad)
    assert float((bn.weight.grad - bn_ref.weight.grad).abs().max()) <= tol(bn_ref.weight.grad)
    assert float((bn.bias.grad - bn_ref.bias.grad).abs().max()) <= tol(bn_ref.bias.grad)
    # moving averages: (1 - momentum) * old + momentum * batch (unbiased variance), as the framework's fused batch norm
    torch.testing.assert_close(bn.running_mean, bn_ref.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bn.running_var, bn_ref.running_var, rtol=1e-5, atol=1e-6)


def test_glue_is_deterministic():
    raw, skip, planes, bn = _case(2, 31, 45, 64, 96, 1, seed=3)
    bn = bn.to(DEV).train()
    outs, grads = [], []
    for _ in range(3):
        r = raw.to(DEV).requires_grad_(True)
        out = ops.conv_block_glue(r, skip.to(DEV), [p.to(DEV) for p in planes], bn, pad=3)
        out.backward(torch.ones_like(out) * 0.5 + out.detach() * 0.1)
        outs.append(out.detach().clone())
        grads.append((r.grad.clone(), bn.weight.grad.clone()))
        bn.weight.grad = None
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[2][1])


def test_glue_rejects_unsupported_channel_counts():
    assert not ops.bn_glue_supported(48, 52, torch.float32)          # not a power of two: the decoder falls back to the framework's ops
    assert not ops.bn_glue_supported(32, 41, torch.float32)          # concat width not a multiple of 4
    assert ops.bn_glue_supported(32, 44, torch.float32)
    raw = torch.randn(1, 4, 4, 48, device=DEV)
    with pytest.raises(ValueError):
        ops.bn_elu_stats(raw, torch.ones(48, device=DEV), torch.zeros(48, device=DEV), None, None, 0.01, 1e-5)
