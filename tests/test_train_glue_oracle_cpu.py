"""CPU pins of the oracle restatements behind the training-glue kernels (oracle/tail_oracle.py: conv_block_glue, bn_relu_backward,
conv3x3_wgrad) against the framework's own layers with autograd in float64 -- the same layer semantics the reference builds with
Keras (bts_decoder.py:30-54, :98-100: Conv2D, BatchNormalization(training), ELU / ReLU, Concatenate)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import tail_oracle


def test_conv3x3_wgrad_oracle_matches_conv2d_autograd():
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(2, 7, 9, 5, generator=gen, dtype=torch.float64)
    g = torch.randn(2, 7, 9, 3, generator=gen, dtype=torch.float64)
    w = torch.zeros(3, 5, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.permute(0, 3, 1, 2), w, padding=1).backward(g.permute(0, 3, 1, 2))
    ref = w.grad.permute(2, 3, 1, 0).numpy()                                   # OIHW -> HWIO
    np.testing.assert_allclose(tail_oracle.conv3x3_wgrad(x.float().numpy(), g.float().numpy()), ref, rtol=1e-5, atol=1e-5)
    # TF32 operands: the low 13 mantissa bits are cut
    cut = tail_oracle.conv3x3_wgrad(x.float().numpy(), g.float().numpy(), tf32_operands=True)
    assert 0 < np.abs(cut - ref).max() <= 4e-3 * np.abs(ref).max()


def test_bn_relu_backward_oracle_matches_autograd():
    gen = torch.Generator().manual_seed(1)
    x = (torch.randn(3, 5, 6, 8, generator=gen, dtype=torch.float64) * 2 + 0.3).requires_grad_(True)
    gamma = (torch.rand(8, generator=gen, dtype=torch.float64) + 0.5).requires_grad_(True)
    beta = torch.randn(8, generator=gen, dtype=torch.float64).requires_grad_(True)
    z = F.batch_norm(x.permute(0, 3, 1, 2), None, None, gamma, beta, training=True, eps=1.1e-5).permute(0, 2, 3, 1)
    g = torch.randn(3, 5, 6, 8, generator=gen, dtype=torch.float64)
    g2 = torch.randn(3, 5, 6, 8, generator=gen, dtype=torch.float64)
    ((torch.relu(z) * g).sum() + (z * g2).sum()).backward()
    d_x, d_gamma, d_beta = tail_oracle.bn_relu_backward(g.numpy(), x.detach().numpy(), gamma.detach().numpy(), beta.detach().numpy(), 1.1e-5,
                                                        g2=g2.numpy())
    np.testing.assert_allclose(d_x, x.grad.numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(d_gamma, gamma.grad.numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(d_beta, beta.grad.numpy(), rtol=1e-9, atol=1e-10)


def test_conv_block_glue_oracle_matches_framework_layers():
    gen = torch.Generator().manual_seed(2)
    raw = torch.randn(2, 4, 5, 8, generator=gen, dtype=torch.float64)
    skip = torch.randn(2, 4, 5, 3, generator=gen, dtype=torch.float64)
    plane = torch.rand(2, 4, 5, 1, generator=gen, dtype=torch.float64)
    gamma, beta = torch.rand(8, generator=gen, dtype=torch.float64) + 0.5, torch.randn(8, generator=gen, dtype=torch.float64)
    up = F.batch_norm(F.elu(raw).permute(0, 3, 1, 2), None, None, gamma, beta, training=True, eps=1.1e-5).permute(0, 2, 3, 1)
    ref = torch.cat([up, skip, plane, torch.zeros(2, 4, 5, 4, dtype=torch.float64)], 3).numpy()
    out, mean, var = tail_oracle.conv_block_glue(raw.numpy(), skip.numpy(), [plane.numpy()], gamma.numpy(), beta.numpy(), 1.1e-5, pad=4)
    np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(mean, F.elu(raw).reshape(-1, 8).mean(0).numpy(), rtol=1e-12)


def test_sub_grid_form_of_a_dilated_convolution_is_the_same_convolution():
    """decoder._dilation_split: Conv2D(kernel_size=3, dilation_rate=d, padding='same') (bts_decoder.py:53) == 2 x 2 independent rate-d/2
    convolutions on the interleaved sub-grids, forward and both gradients, in float64 on the CPU (host logic, no kernel involved)."""
    from bts_fully_tf_b200 import decoder as decoder_mod
    for rate, H, W in ((18, 10, 14), (24, 12, 8), (24, 44, 20)):
        torch.manual_seed(rate + H)
        conv = decoder_mod._conv(6, 4, k=3, dilation=rate).double()
        x = torch.randn(2, H, W, 6, dtype=torch.float64)
        g = torch.randn(2, H, W, 4, dtype=torch.float64)
        assert decoder_mod._dilation_split(conv, H, W) == 2
        assert decoder_mod._dilation_split(conv, H + 1, W) == 1 and decoder_mod._dilation_split(decoder_mod._conv(6, 4, k=3, dilation=12), H, W) == 1
        xs = x.clone().requires_grad_(True)
        y = decoder_mod._conv_nhwc(xs, conv)
        y.backward(g)
        gw_split, gx_split = conv.weight.grad.clone(), xs.grad.clone()
        conv.weight.grad = None
        xr = x.clone().requires_grad_(True)
        y_ref = F.conv2d(xr.permute(0, 3, 1, 2), conv.weight, None, 1, rate, rate).permute(0, 2, 3, 1)
        y_ref.backward(g)
        np.testing.assert_allclose(y.detach().numpy(), y_ref.detach().numpy(), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(gx_split.numpy(), xr.grad.numpy(), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(gw_split.numpy(), conv.weight.grad.numpy(), rtol=1e-12, atol=1e-12)
        # the sub-grid re-ordering itself
        s = decoder_mod._s2b(x, 2)
        assert torch.equal(s[1], x[0, 0::2, 1::2]) and torch.equal(s[2], x[0, 1::2, 0::2]) and torch.equal(s[4], x[1, 0::2, 0::2])
        assert torch.equal(decoder_mod._b2s(s, 2), x)


def test_dispatch_predicates_of_the_training_paths():
    """Host logic only: which layers take the hand-written training paths (and which fall back to the framework's)."""
    from bts_fully_tf_b200 import decoder as decoder_mod
    from bts_fully_tf_b200 import ops
    # conv-block glue: float32, power-of-two channel count, concat width a multiple of 4
    assert ops.bn_glue_supported(64, 64 + 96 + 1 + 3, torch.float32)
    assert not ops.bn_glue_supported(48, 48 + 16, torch.float32) and not ops.bn_glue_supported(64, 161, torch.float32)
    assert not ops.bn_glue_supported(64, 164, torch.bfloat16)
    # DenseASPP glue: half-width pieces must be whole 16-byte vectors, slices at most 1024 channels
    assert ops.bn_slices_supported(256, torch.float32) and ops.bn_slices_supported(128, torch.float32)
    assert not ops.bn_slices_supported(100, torch.float32) and not ops.bn_slices_supported(1024, torch.float32)
    # tensor-core weight gradient: never on the CPU / without grad / for wide layers / with the TF32 switch off
    w = torch.zeros(16, 20, 3, 3, requires_grad=True)
    x = torch.zeros(32, 20, 64, 64)
    assert not decoder_mod._tc_wgrad_applies(x, w)                         # CPU tensor
    class _Cuda:                                                           # shape / flag logic without a device
        is_cuda, dtype = True, torch.float32
        def __init__(self, shape): self.shape = shape
    old = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = True
        assert decoder_mod._tc_wgrad_applies(_Cuda((32, 20, 64, 64)), w)
        assert not decoder_mod._tc_wgrad_applies(_Cuda((1, 20, 64, 64)), w)                                          # too few pixels
        assert not decoder_mod._tc_wgrad_applies(_Cuda((32, 128, 64, 64)), torch.zeros(128, 128, 3, 3, requires_grad=True))   # wide: library
        assert not decoder_mod._tc_wgrad_applies(_Cuda((32, 20, 64, 64)), torch.zeros(16, 20, 1, 1, requires_grad=True))      # not 3x3
        assert not decoder_mod._tc_wgrad_applies(_Cuda((32, 20, 64, 64)), w.detach())                                # no gradient wanted
        with torch.no_grad():
            assert not decoder_mod._tc_wgrad_applies(_Cuda((32, 20, 64, 64)), w)
        torch.backends.cudnn.allow_tf32 = False
        assert not decoder_mod._tc_wgrad_applies(_Cuda((32, 20, 64, 64)), w)                                         # float32 convolutions asked for
    finally:
        torch.backends.cudnn.allow_tf32 = old
