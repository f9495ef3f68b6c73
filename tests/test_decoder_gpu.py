"""GPU test of the decoder wiring (bts_decoder.py:26-105) around the fused LPG heads against the
fixture recorded from the UNMODIFIED reference decoder_model (tests/golden/decoder_small.npz),
forward and backward, inference-mode and training-mode BatchNorm."""
import os

import numpy as np
import pytest
import torch

from bts_fully_tf_b200 import LocalPlanarGuidance, ops
from bts_fully_tf_b200.decoder import BtsDecoder, decoder_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _load(golden_dir):
    z = np.load(os.path.join(golden_dir, "decoder_small.npz"))
    feats = [torch.from_numpy(z["feat_" + k]).float().to(DEV).requires_grad_(True) for k in ("dense", "s2", "s4", "s8", "s16")]
    kernels = [z["kernel_%02d" % i] for i in range(int(z["n_convs"]))]
    return z, feats, kernels


@pytest.mark.parametrize("tag,training", [("infer", False), ("train", True)])
def test_decoder_matches_reference_fixture(golden_dir, tag, training):
    z, feats, kernels = _load(golden_dir)
    dec = BtsDecoder([f.shape[-1] for f in feats], 10.0, num_filters=32).to(DEV)
    dec.load_keras_kernels(kernels)
    depth = decoder_model(feats, 10.0, num_filters=32, is_training=training, decoder=dec)
    assert tuple(depth.shape) == z[tag + "_depth_est"].shape
    for r in (8, 4, 2):
        np.testing.assert_allclose(dec.intermediates["reduction_%dx%d" % (r, r)].detach().cpu().numpy(), z["%s_head%d_out" % (tag, r)],
                                   rtol=2e-4, atol=2e-6)
        np.testing.assert_allclose(dec.intermediates["depth_%dx%d_scaled" % (r, r)].detach().cpu().numpy(),
                                   z["%s_depth_%dx%d_scaled" % (tag, r, r)], rtol=5e-4, atol=1e-5)
    np.testing.assert_allclose(depth.detach().cpu().numpy(), z[tag + "_depth_est"], rtol=5e-4, atol=1e-5)

    depth.backward(torch.from_numpy(z["g_depth"]).float().to(DEV))
    grads = dec.keras_kernel_grads()
    for i, g in enumerate(grads):
        ref = z["%s_gkernel_%02d" % (tag, i)]
        assert np.abs(g.detach().cpu().numpy() - ref).max() <= 2e-3 * np.abs(ref).max() + 1e-7, "kernel %d" % i
    for k, f in zip(("dense", "s2", "s4", "s8", "s16"), feats):
        ref = z["%s_gfeat_%s" % (tag, k)]
        assert np.abs(f.grad.cpu().numpy() - ref).max() <= 2e-3 * np.abs(ref).max() + 1e-7, k


@pytest.mark.parametrize("tag,training", [("infer", False), ("train", True)])
def test_decoder_f256_matches_reference_fixture(golden_dir, tag, training):
    """The UNMODIFIED reference decoder_model at num_filters = 256 (tests/golden/decoder_f256.npz): the channel counts
    that take every fused fast path -- TMA-staged heads at C = 64 / 64 / 32, concats with pad channels, sub-pixel upconvs,
    the last convolution with iconv1's ELU (and, in inference, sigmoid * max_depth) folded in, its fused backward."""
    from oracle import decoder_fixture
    z = np.load(os.path.join(golden_dir, "decoder_f256.npz"))
    kernels = decoder_fixture.regen_kernels([tuple(s) for s in z["kernel_shapes"]], int(z["seed"]))
    np.testing.assert_allclose([float(k.sum()) for k in kernels], z["kernel_sums"], rtol=0, atol=1e-9)     # the same draws as the recorded run
    feats = [torch.from_numpy(z["feat_" + k]).float().to(DEV).requires_grad_(True) for k in ("dense", "s2", "s4", "s8", "s16")]
    dec = BtsDecoder([f.shape[-1] for f in feats], 10.0, num_filters=int(z["num_filters"])).to(DEV)
    dec.load_keras_kernels([k.float() for k in kernels])
    if not training:
        with torch.no_grad():                                   # the inference path proper (folded BatchNorm, fused tail)
            d_inf = decoder_model([f.detach() for f in feats], 10.0, num_filters=256, is_training=False, decoder=dec)
        assert ops.last_kernel() == "depthconv_fwd<f32,C16,elu>", ops.last_kernel()
        np.testing.assert_allclose(d_inf.cpu().numpy(), z["infer_depth_est"], rtol=5e-4, atol=1e-5)
    depth = decoder_model(feats, 10.0, num_filters=256, is_training=training, decoder=dec)
    for r in (8, 4, 2):
        np.testing.assert_allclose(dec.intermediates["reduction_%dx%d" % (r, r)].detach().cpu().numpy(), z["%s_head%d_out" % (tag, r)],
                                   rtol=2e-4, atol=2e-6)
        np.testing.assert_allclose(dec.intermediates["depth_%dx%d_scaled" % (r, r)].detach().cpu().numpy(),
                                   z["%s_depth_%dx%d_scaled" % (tag, r, r)], rtol=5e-4, atol=1e-5)
    np.testing.assert_allclose(depth.detach().cpu().numpy(), z[tag + "_depth_est"], rtol=5e-4, atol=1e-5)
    depth.backward(torch.from_numpy(z["g_depth"]).float().to(DEV))
    for i, g in enumerate(dec.keras_kernel_grads()):
        flat = g.detach().reshape(-1).cpu()
        got = flat[decoder_fixture.sample_index(flat.numel())].numpy()
        scale = float(z["%s_gkernel_absmax_%02d" % (tag, i)])
        assert np.abs(got - z["%s_gkernel_%02d" % (tag, i)]).max() <= 2e-3 * scale + 1e-7, "kernel %d" % i
    for k, f in zip(("dense", "s2", "s4", "s8", "s16"), feats):
        ref = z["%s_gfeat_%s" % (tag, k)]
        assert np.abs(f.grad.cpu().numpy() - ref).max() <= 2e-3 * np.abs(ref).max() + 1e-7, k


def test_fused_heads_equal_unfused_composition():
    """ReductionLPG (one kernel) == torch 1x1 conv + sigmoid -> LocalPlanarGuidance layer -> slice, at
    channel counts that take the fused fast path (F=256: C = 64, 64, 32)."""
    torch.manual_seed(0)
    B, H, W, F = 2, 64, 96, 256
    chans = [48, 24, 24, 32, 40]
    feats = [torch.randn(B, H // s, W // s, c, device=DEV) for s, c in zip((32, 2, 4, 8, 16), chans)]
    dec = BtsDecoder(chans, 10.0, num_filters=F).to(DEV).eval()
    depth = dec(feats)
    assert ops.launch_count() > 0
    for r, d, head in ((8, 4, dec.reduction_8x8), (4, 2, dec.reduction_4x4), (2, 0, dec.reduction_2x2)):
        red = dec.intermediates["reduction_%dx%d" % (r, r)]
        layer = LocalPlanarGuidance(upratio=r, name="depth_%dx%d_scaled" % (r, r))
        np.testing.assert_array_equal(layer(red).cpu().numpy(), dec.intermediates["depth_%dx%d_scaled" % (r, r)].detach().cpu().numpy())
    depth = depth.detach()
    assert torch.isfinite(depth).all() and float(depth.max()) <= 10.0 and float(depth.min()) >= 0.0


@pytest.mark.parametrize("dataset,max_depth", [("nyu", 10.0), ("kitti", 80.0)])
def test_forward_loss_equals_unfused_training_step(dataset, max_depth):
    """decoder.forward_loss (fused sigmoid*max_depth + si_log_loss kernels) == forward() + the torch
    restatement of bts.py:27-41, for the loss and for every decoder gradient."""
    from bts_fully_tf_b200.decoder import si_log_loss
    torch.manual_seed(1)
    B, H, W, F = 2, 64, 96, 64
    chans = [24, 8, 8, 12, 16]
    feats = [torch.randn(B, H // s, W // s, c, device=DEV) for s, c in zip((32, 2, 4, 8, 16), chans)]
    gt = torch.rand(B, H, W, 1, device=DEV) * max_depth
    gt[:, :10] = 0.0
    dec = BtsDecoder(chans, max_depth, num_filters=F).to(DEV).train()
    depth_a, loss_a = dec.forward_loss(feats, gt, dataset)
    loss_a.backward()
    grads_a = [p.grad.clone() for p in dec.parameters()]
    dec.zero_grad(set_to_none=True)
    depth_b = dec(feats)
    loss_b = si_log_loss(gt, depth_b, dataset)
    loss_b.backward()
    np.testing.assert_allclose(depth_a.cpu().numpy(), depth_b.detach().cpu().numpy(), rtol=2e-6)
    np.testing.assert_allclose(float(loss_a.detach()), float(loss_b.detach()), rtol=1e-5)
    for ga, p in zip(grads_a, dec.parameters()):
        scale = float(p.grad.abs().max())
        assert float((ga - p.grad).abs().max()) <= 2e-4 * scale + 1e-9


def test_inference_fast_path_equals_eval_path():
    """Under torch.no_grad() in eval mode the conv blocks fold ELU + BatchNormalization into the concat pass;
    the result must equal the autograd-capable eval path (torch ELU + BatchNorm + the plain concat kernel)."""
    torch.manual_seed(3)
    B, H, W, F = 2, 64, 96, 64
    chans = [24, 8, 8, 12, 16]
    feats = [torch.randn(B, H // s, W // s, c, device=DEV) for s, c in zip((32, 2, 4, 8, 16), chans)]
    dec = BtsDecoder(chans, 10.0, num_filters=F).to(DEV)
    dec.train()
    for _ in range(2):                                   # give the BatchNorm running statistics non-trivial values
        dec(feats)
    dec.eval()
    ops.reset_launch_count()
    with torch.no_grad():
        fast = dec(feats)
    assert ops.launch_count() >= 3 + 5 + 5               # heads, up-samplings, concats
    slow = dec(feats)                                    # grad enabled: unfused BatchNorm path
    np.testing.assert_allclose(fast.cpu().numpy(), slow.detach().cpu().numpy(), rtol=2e-4, atol=2e-5)


def test_inference_with_tcgen05_iconv1_matches_reference_fixture(golden_dir):
    """With the framework's TF32 switch on (its default), inference runs iconv1 as the tcgen05 implicit GEMM over the concat's
    sources (no concat1).  TF32 operands: the depth map agrees with the float64 reference run of the UNMODIFIED bts_decoder.py to
    the tolerance of a TF32 convolution chain (5e-3), and with this repo's own float32 path to the same."""
    from oracle import decoder_fixture
    z = np.load(os.path.join(golden_dir, "decoder_f256.npz"))
    feats = [torch.from_numpy(z["feat_" + k]).float().to(DEV) for k in ("dense", "s2", "s4", "s8", "s16")]
    dec = BtsDecoder([f.shape[-1] for f in feats], 10.0, num_filters=int(z["num_filters"])).to(DEV).eval()
    dec.load_keras_kernels([k.float() for k in decoder_fixture.regen_kernels([tuple(s) for s in z["kernel_shapes"]], int(z["seed"]))])
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        with torch.no_grad():
            exact = dec(feats)
        torch.backends.cudnn.allow_tf32 = True
        ops.reset_launch_count()
        with torch.no_grad():
            fast = dec(feats)
        torch.cuda.synchronize()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    np.testing.assert_allclose(exact.cpu().numpy(), z["infer_depth_est"], rtol=5e-4, atol=1e-5)
    np.testing.assert_allclose(fast.cpu().numpy(), z["infer_depth_est"], rtol=5e-3, atol=1e-4)
    dec.tensor_core_iconv1 = False
    torch.backends.cudnn.allow_tf32 = True
    try:
        with torch.no_grad():
            lib = dec(feats)
    finally:
        torch.backends.cudnn.allow_tf32 = old[0]
    np.testing.assert_allclose(fast.cpu().numpy(), lib.cpu().numpy(), rtol=5e-3, atol=1e-4)
