"""GPU parity of the TRAINING-mode DenseASPP glue (bts_decoder.py:46-76 with is_training): the fused path (shared per-channel batch
moments, affine_act forward, statistics + accumulate backward) against the framework's own layers (torch.cat + BatchNorm2d(train) +
ReLU with autograd) in float64 on the CPU, and the primitive kernels against closed forms."""
import copy

import numpy as np
import pytest
import torch

from bts_fully_tf_b200 import ops
from bts_fully_tf_b200.decoder import BtsDecoder

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _exact_convs():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def _decoder(num_filters, seed):
    torch.manual_seed(seed)
    dec = BtsDecoder([num_filters, 8, 8, 8, 8], max_depth=10.0, num_filters=num_filters)
    with torch.no_grad():
        for m in dec.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0.0, 0.3)
                m.running_mean.normal_()
                m.running_var.uniform_(0.5, 1.5)
    return dec.train()


@pytest.mark.parametrize("num_filters,B,h,w", [(64, 2, 7, 9), (128, 1, 12, 16), (32, 3, 5, 6), (64, 2, 52, 68)])
def test_daspp_training_matches_framework_layers(num_filters, B, h, w, monkeypatch):
    """Even map sizes also cover the rate-18 / rate-24 convolutions run as 2 x 2 interleaved sub-grids (decoder._dilation_split); the
    float64 reference runs every layer as the framework defines it."""
    from bts_fully_tf_b200 import decoder as decoder_mod
    dec = _decoder(num_filters, seed=num_filters)
    ref = copy.deepcopy(dec).double()
    ref.fused_training_glue = False
    dec = dec.to(DEV)
    nf = num_filters // 2
    x = torch.randn(B, nf, h, w, generator=torch.Generator().manual_seed(1)) * 1.3 + 0.1
    x_gpu = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    x_ref = x.double().requires_grad_(True)
    out = dec._daspp(x_gpu)
    assert (decoder_mod._dilation_split(dec.daspp_24.conv2, h, w) == 2) == (h % 2 == 0 and w % 2 == 0)
    monkeypatch.setattr(decoder_mod, "SPLIT_DILATION_FROM", 10 ** 9)            # the reference: no splitting
    out_ref = ref._daspp(x_ref)
    monkeypatch.undo()
    g = torch.randn(out_ref.shape, generator=torch.Generator().manual_seed(2))
    out.backward(g.to(DEV))
    out_ref.backward(g.double())
    torch.cuda.synchronize()

    def close(a, b, what, rel=2e-4):
        a, b = a.detach().cpu().double(), b.detach().double()
        err, scale = float((a - b).abs().max()), float(b.abs().max())
        assert err <= rel * scale + 1e-7, "%s: max error %.3g against a scale of %.3g" % (what, err, scale)

    close(out, out_ref, "daspp_feat")
    close(x_gpu.grad, x_ref.grad, "d iconv4")
    named, named_ref = dict(dec.named_parameters()), dict(ref.named_parameters())
    checked = 0
    for name, p in named.items():
        if named_ref[name].grad is None:
            assert p.grad is None, name
            continue
        close(p.grad, named_ref[name].grad, "d " + name)
        checked += 1
    assert checked == 2 + 4 * 5 + 2 * 4 + 1                     # bn4, five blocks (conv1, bn2 x2, conv2), four bn_first, daspp_feat
    bufs, bufs_ref = dict(dec.named_buffers()), dict(ref.named_buffers())
    for name, b in bufs.items():
        if name.startswith(("bn4", "daspp_")):
            close(b.float(), bufs_ref[name].double(), name, rel=1e-5)


def test_daspp_training_is_deterministic_and_leaves_the_upstream_gradient_alone():
    torch.backends.cudnn.deterministic = True          # the library's own weight-gradient algorithms may use atomics otherwise
    try:
        dec = _decoder(64, seed=5).to(DEV)
        x = torch.randn(2, 32, 9, 11, device=DEV).contiguous(memory_format=torch.channels_last)
        g = torch.randn(2, 16, 9, 11, device=DEV)
        res = []
        for _ in range(2):
            dec.zero_grad(set_to_none=True)
            xi = x.clone().requires_grad_(True)
            out = dec._daspp(xi)
            g_in = g.clone()
            out.backward(g_in)
            assert torch.equal(g_in, g)
            res.append((out.detach().clone(), xi.grad.clone(), dec.daspp_12.bn_first.weight.grad.clone(), dec.daspp_3.conv1.weight.grad.clone()))
    finally:
        torch.backends.cudnn.deterministic = False
    for what, a, b in zip(("daspp_feat", "d iconv4", "d gamma", "d conv1"), *res):
        assert torch.equal(a, b), what


def test_bn_moments_and_backward_on_slices():
    """The primitives on channel slices of a wider buffer, against float64 closed forms."""
    gen = torch.Generator().manual_seed(0)
    B, H, W, CT, c0, C = 2, 6, 7, 96 + 8, 8, 96                  # 96 channels: 24 vectors per pixel, a 240-thread CTA
    buf = (torch.randn(B, H, W, CT, generator=gen) * 2 + 0.5).to(DEV)
    x = buf[..., c0:c0 + C]
    mean, var = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    ops.bn_moments(x, mean, var)
    xd = x.double().reshape(-1, C)
    torch.testing.assert_close(mean.double(), xd.mean(0), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(var.double(), xd.var(0, unbiased=False), rtol=1e-5, atol=1e-6)
    gamma, beta = (torch.rand(C, generator=gen) + 0.5).to(DEV), torch.randn(C, generator=gen).to(DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    vecs = ops.bn_fold(mean, var, gamma, beta, rm, rv, 0.01, 1.1e-5, B * H * W)
    n = B * H * W
    torch.testing.assert_close(rm.double(), 0.01 * xd.mean(0), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(rv.double(), 0.99 + 0.01 * xd.var(0, unbiased=True), rtol=1e-5, atol=1e-7)
    # backward against autograd in float64
    xr = xd.clone().requires_grad_(True)
    gr, br = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    xhat = (xr - xr.mean(0)) / torch.sqrt(xr.var(0, unbiased=False) + 1.1e-5)
    z = xhat * gr + br
    y = torch.relu(z)
    g = torch.randn(n, C, generator=gen).to(DEV)
    g2 = torch.randn(n, C, generator=gen).to(DEV)
    (y * g.double()).sum().backward(retain_graph=True)
    (z * g2.double()).sum().backward()
    gg, gb = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    dst = torch.full((B, H, W, CT), 3.0, device=DEV)
    ops.bn_act_backward(g.reshape(B, H, W, C), x, vecs, gg, gb, dst[..., c0:c0 + C], relu=True, accumulate=True, g2=g2.reshape(B, H, W, C))
    torch.testing.assert_close((dst[..., c0:c0 + C] - 3.0).double().reshape(-1, C), xr.grad, rtol=1e-4, atol=2e-5)
    assert bool((dst[..., :c0] == 3.0).all()) and bool((dst[..., c0 + C:] == 3.0).all())       # nothing outside the slice
    torch.testing.assert_close(gg.double(), gr.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(gb.double(), br.grad, rtol=1e-4, atol=1e-4)


def test_bn_slice_argument_checks():
    x = torch.randn(1, 4, 4, 6, device=DEV)
    with pytest.raises(ValueError):
        ops.bn_moments(x, torch.empty(6, device=DEV), torch.empty(6, device=DEV))          # 6 channels: not whole 16-byte vectors
    buf = torch.randn(1, 4, 4, 18, device=DEV)
    with pytest.raises(ValueError):
        ops.bn_moments(buf[..., 1:9], torch.empty(8, device=DEV), torch.empty(8, device=DEV))   # misaligned slice
    assert not ops.bn_slices_supported(36, torch.float32) and ops.bn_slices_supported(128, torch.float32)
    assert not ops.bn_slices_supported(128, torch.bfloat16)


@pytest.mark.parametrize("rate,H,W", [(18, 12, 20), (24, 44, 152), (24, 6, 8)])
def test_split_dilated_convolution_equals_the_layer(rate, H, W, monkeypatch):
    """Forward, d input and d kernel of a high-rate dilated convolution through the sub-grid form against the layer as it is."""
    from bts_fully_tf_b200 import decoder as decoder_mod
    torch.manual_seed(rate)
    conv = decoder_mod._conv(16, 8, k=3, dilation=rate).to(DEV)
    x = torch.randn(2, H, W, 16, device=DEV)
    g = torch.randn(2, H, W, 8, device=DEV)
    assert decoder_mod._dilation_split(conv, H, W) == 2
    y = decoder_mod._conv_nhwc(x, conv)
    g_x, g_w = decoder_mod._conv_backward(g, x, conv)
    monkeypatch.setattr(decoder_mod, "SPLIT_DILATION_FROM", 10 ** 9)
    y0 = decoder_mod._conv_nhwc(x, conv)
    g_x0, g_w0 = decoder_mod._conv_backward(g, x, conv)
    for a, b in ((y, y0), (g_x, g_x0), (g_w, g_w0)):
        a, b = a.detach(), b.detach()
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) + 1e-6


@pytest.mark.parametrize("s,B,H,W,CT,c0,C", [(2, 2, 6, 8, 24, 4, 16), (2, 1, 44, 152, 128, 0, 128), (3, 2, 6, 9, 7, 2, 5)])
def test_affine_act_sub_grid_forms(s, B, H, W, CT, c0, C):
    """ops.affine_act with one side in sub-grid form == the plain pass followed / preceded by the re-ordering copy."""
    from bts_fully_tf_b200 import decoder as decoder_mod
    gen = torch.Generator().manual_seed(s * 100 + C)
    buf = torch.randn(B, H, W, CT, generator=gen).to(DEV)
    scale, shift = (torch.rand(C, generator=gen) + 0.5).to(DEV), torch.randn(C, generator=gen).to(DEV)
    src = buf[..., c0:c0 + C]
    plain = ops.affine_act(src, scale=scale, shift=shift, act=ops.ACT_RELU)
    to_split = ops.affine_act(src, scale=scale, shift=shift, act=ops.ACT_RELU, dst_split=s)
    assert to_split.shape == (B * s * s, H // s, W // s, C)
    assert torch.equal(to_split, decoder_mod._s2b(plain, s))
    back = torch.full((B, H, W, CT), -7.0, device=DEV)
    ops.affine_act(to_split, dst=back[..., c0:c0 + C], src_split=s)
    assert torch.equal(back[..., c0:c0 + C], plain)
    assert bool((back[..., :c0] == -7.0).all()) and bool((back[..., c0 + C:] == -7.0).all())
    with pytest.raises(ValueError):
        ops.affine_act(src, dst=torch.empty(B * s * s, H // s, W // s + 1, C, device=DEV), dst_split=s)
