"""GPU parity of the tcgen05 weight gradient of the full-resolution 3x3 convolutions (bts_decoder.py:98, :100) against the float64
oracle (oracle/tail_oracle.conv3x3_wgrad, pinned to conv2d's own autograd in tests/test_train_glue_oracle_cpu.py) and, structurally,
against one-hot inputs that isolate single taps."""
import numpy as np
import pytest
import torch

from bts_fully_tf_b200 import ops
from oracle import tail_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 16, 32, 32), (2, 11, 37, 32, 16), (1, 16, 48, 20, 16), (2, 9, 33, 64, 32), (1, 24, 40, 36, 32),
                                            (3, 5, 7, 8, 4), (1, 1, 1, 4, 4), (2, 13, 21, 32, 64), (1, 9, 40, 64, 128), (1, 10, 17, 24, 100), (2, 8, 16, 64, 64), (1, 12, 20, 100, 32), (1, 6, 18, 164, 64), (1, 5, 9, 256, 128), (1, 7, 11, 40, 24)])
def test_wgrad_matches_float64_definition(B, H, W, Cin, Cout):
    gen = torch.Generator().manual_seed(B * 1000 + H * 10 + Cin)
    x = torch.randn(B, H, W, Cin, generator=gen)
    g = torch.randn(B, H, W, Cout, generator=gen)
    out = ops.conv3x3_wgrad(x.to(DEV), g.to(DEV)).cpu()
    out = out.double().numpy()
    exact = tail_oracle.conv3x3_wgrad(x.numpy(), g.numpy(), tf32_operands=True)       # same operand bits: only float32 accumulation differs
    assert np.abs(out - exact).max() <= 2e-5 * np.abs(exact).max() + 1e-6
    full = tail_oracle.conv3x3_wgrad(x.numpy(), g.numpy())                            # untruncated operands: TF32's 2^-10 per operand
    assert np.abs(out - full).max() <= 3e-3 * np.abs(full).max()


def test_wgrad_one_hot_isolates_taps():
    """x one-hot at (y0, x0, ci), g one-hot at (y1, x1, co): the only non-zero entry is [y0 - y1 + 1][x0 - x1 + 1][ci][co]."""
    B, H, W, Cin, Cout = 1, 12, 20, 32, 16
    for (y0, x0, ci), (y1, x1, co) in [((3, 4, 5), (3, 4, 6)), ((0, 0, 31), (1, 1, 0)), ((11, 19, 0), (10, 18, 15)), ((5, 16, 7), (6, 15, 3)),
                                       ((7, 8, 9), (7, 10, 2))]:
        x = torch.zeros(B, H, W, Cin)
        g = torch.zeros(B, H, W, Cout)
        x[0, y0, x0, ci] = 2.0
        g[0, y1, x1, co] = 3.0
        out = ops.conv3x3_wgrad(x.to(DEV), g.to(DEV)).cpu()
        want = torch.zeros(3, 3, Cin, Cout)
        ky, kx = y0 - y1 + 1, x0 - x1 + 1
        if 0 <= ky < 3 and 0 <= kx < 3:
            want[ky, kx, ci, co] = 6.0
        assert torch.equal(out, want), ((y0, x0, ci), (y1, x1, co), out.nonzero().tolist())


def test_wgrad_is_deterministic_and_rejects_bad_shapes():
    x = torch.randn(2, 40, 64, 32, device=DEV)
    g = torch.randn(2, 40, 64, 16, device=DEV)
    a, b = ops.conv3x3_wgrad(x, g), ops.conv3x3_wgrad(x, g)
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        ops.conv3x3_wgrad(torch.randn(1, 4, 4, 260, device=DEV), torch.randn(1, 4, 4, 16, device=DEV))
    with pytest.raises(ValueError):
        ops.conv3x3_wgrad(torch.randn(1, 4, 4, 32, device=DEV), torch.randn(1, 4, 4, 132, device=DEV))
    with pytest.raises(ValueError):
        ops.conv3x3_wgrad(torch.randn(1, 4, 4, 32, device=DEV), torch.randn(1, 4, 5, 16, device=DEV))


def test_decoder_training_gradients_with_and_without_the_tensor_core_wgrad():
    """A training step of the decoder with the tcgen05 weight gradients (the default while torch's TF32 switch is on) against the same
    step with the library's TF32 kernels: every parameter gradient agrees to TF32 accuracy."""
    from bts_fully_tf_b200 import decoder as decoder_mod
    from bts_fully_tf_b200.decoder import BtsDecoder
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True              # the switch the tensor-core path follows (another test module may have left it off)
    torch.manual_seed(3)
    B, H, W, F = 4, 128, 256, 128
    chans = [24, 8, 8, 12, 16]
    feats = [torch.randn(B, H // s, W // s, c, device=DEV) for s, c in zip((32, 2, 4, 8, 16), chans)]
    gt = torch.rand(B, H, W, 1, device=DEV) * 10.0
    dec = BtsDecoder(chans, 10.0, num_filters=F).to(DEV).train()
    grads = []
    used = []
    real = ops.conv3x3_wgrad
    try:
        for on in (True, False):
            decoder_mod.TENSOR_CORE_WGRAD = on
            ops.conv3x3_wgrad = lambda x, g, out=None: (used.append((x.shape[-1], g.shape[-1])), real(x, g, out))[1]
            dec.zero_grad(set_to_none=True)
            _, loss = dec.forward_loss(feats, gt, "nyu")
            loss.backward()
            grads.append({n: p.grad.clone() for n, p in dec.named_parameters()})
    finally:
        decoder_mod.TENSOR_CORE_WGRAD = True
        ops.conv3x3_wgrad = real
        torch.backends.cudnn.allow_tf32 = old_tf32
    assert len(used) >= 3, used                     # up4 (sub-pixel upconv1), iconv1 and conv block 2 at least
    for n, ga in grads[0].items():
        gb = grads[1][n]
        scale = float(gb.abs().max())
        assert float((ga - gb).abs().max()) <= 1e-2 * scale + 1e-8, n


def test_wgrad_leaves_the_shared_workspace_header_alone():
    """ops shares one scratch buffer per stream between the kernels that need one; its first 256 bytes hold the counters of the
    last-CTA reductions (always left zero).  A fused head backward before and after a weight gradient gives identical bits."""
    gen = torch.Generator().manual_seed(0)
    feat = torch.randn(2, 24, 32, 32, generator=gen).to(DEV)
    kernel = (torch.randn(32, 3, generator=gen) * 0.3).to(DEV)

    def head_grads():
        f, k = feat.clone().requires_grad_(True), kernel.clone().requires_grad_(True)
        _, full, ds = ops.reduce_lpg(f, k, 4, 2)
        (full.sum() * 0.5 + (ds * ds).sum()).backward()
        return f.grad.clone(), k.grad.clone()

    before = head_grads()
    ops.conv3x3_wgrad(torch.randn(2, 40, 64, 32, device=DEV), torch.randn(2, 40, 64, 16, device=DEV))
    ws = ops._workspace(torch.device(DEV), 256)
    assert int(ws[:256].to(torch.int32).abs().sum()) == 0
    after = head_grads()
    assert torch.equal(before[0], after[0]) and torch.equal(before[1], after[1])


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(32, 416, 544, 20, 16), (16, 176, 608, 64, 128)])
def test_wgrad_at_benchmark_sizes(B, H, W, Cin, Cout):
    """BASELINE-size launches (config 5's iconv1, config 4's sub-pixel upconv1 at half the batch), where the float64 oracle would take
    minutes: against the library's float32 (TF32 off) weight gradient, and linearity in the gradient operand."""
    gen = torch.Generator(device=DEV).manual_seed(7)
    x = torch.randn(B, H, W, Cin, device=DEV, generator=gen)
    g1 = torch.randn(B, H, W, Cout, device=DEV, generator=gen)
    g2 = torch.randn(B, H, W, Cout, device=DEV, generator=gen)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        w = torch.zeros(Cout, Cin, 3, 3, device=DEV)
        ref = torch.ops.aten.convolution_backward(g1.permute(0, 3, 1, 2), x.permute(0, 3, 1, 2), w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                                  [False, True, False])[1].permute(2, 3, 1, 0)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    a = ops.conv3x3_wgrad(x, g1)
    scale = float(ref.abs().max())
    assert float((a - ref).abs().max()) <= 3e-3 * scale
    # linearity: exact products, float32 accumulation -> the sum of two launches equals the launch on the sum up to accumulation rounding
    # (g1 + g2 is rounded to float32 and then cut to TF32, so allow TF32's own 2^-10 on the operand)
    b = ops.conv3x3_wgrad(x, g2)
    c = ops.conv3x3_wgrad(x, g1 + g2)
    assert float((a + b - c).abs().max()) <= 3e-3 * float(c.abs().max())
    # border handling at full size: zero the interior of g, only the outermost ring contributes
    ring = torch.zeros_like(g1)
    ring[:, 0], ring[:, -1], ring[:, :, 0], ring[:, :, -1] = g1[:, 0], g1[:, -1], g1[:, :, 0], g1[:, :, -1]
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref_ring = torch.ops.aten.convolution_backward(ring.permute(0, 3, 1, 2), x.permute(0, 3, 1, 2), w, None, [1, 1], [1, 1], [1, 1], False,
                                                       [0, 0], 1, [False, True, False])[1].permute(2, 3, 1, 0)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    d = ops.conv3x3_wgrad(x, ring)
    assert float((d - ref_ring).abs().max()) <= 3e-3 * float(ref_ring.abs().max())
