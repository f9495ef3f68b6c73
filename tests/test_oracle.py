"""CPU tests of the oracle itself: it is pinned to the UNMODIFIED reference source (run over
oracle/tf_shim, fixtures in tests/golden/) and its two restatements are cross-checked."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, lpg_closed, lpg_literal

RS = [(8, 4), (4, 2), (2, 0)]


def _load(golden_dir, r):
    return np.load(os.path.join(golden_dir, "lpg_r%d.npz" % r))


def _ulp_diff(a, b):
    ai = a.astype(np.float32).view(np.int32).astype(np.int64)
    bi = b.astype(np.float32).view(np.int32).astype(np.int64)
    return np.abs(ai - bi)


@pytest.mark.parametrize("r,d", RS)
def test_c_literal_matches_reference_fixture(golden_dir, r, d):
    """lpg_oracle.c (literal float32 op order) vs the reference source run over the stand-in runtime:
    same arithmetic, different libm -> a handful of ulps at most."""
    z = _load(golden_dir, r)
    out = c_oracle.lpg_forward_f32(z["coef"], r)
    ref = z["out"][..., 0]
    assert out.shape == ref.shape
    assert _ulp_diff(out, ref).max() <= 16
    np.testing.assert_allclose(out, ref, rtol=2e-6, atol=0)


@pytest.mark.parametrize("r,d", RS)
def test_pixel_dir_constant_bit_exact(golden_dir, r, d):
    """custom_layers.py:30-45 constant: bit-exact vs the fixture, and it only holds r*r distinct vectors."""
    z = _load(golden_dir, r)
    ref = z["pixel_dir_unit"][0]
    H, W = ref.shape[:2]
    got = c_oracle.pixel_dir_f32(H, W, r)
    assert np.array_equal(got, ref)
    tiled = np.tile(ref[:r, :r], (H // r, W // r, 1))
    assert np.array_equal(tiled, ref)


@pytest.mark.parametrize("r,d", RS)
def test_closed_form_f64_matches_reference_fixture(golden_dir, r, d):
    z = _load(golden_dir, r)
    out64 = c_oracle.lpg_forward_f64(z["coef"], r)
    # fixture's float64 run uses double pi and the float32 direction constant: agree to ~1e-7
    np.testing.assert_allclose(out64, z["out64"][..., 0], rtol=5e-7)
    np.testing.assert_allclose(out64, z["out"][..., 0], rtol=3e-6)
    g = c_oracle.lpg_backward_f64(z["coef"], z["g_full"], r, z["g_ds"] if d else None, d)
    scale = np.abs(z["g_coef64"]).max()
    assert np.abs(g - z["g_coef64"]).max() <= 2e-6 * scale
    assert np.abs(g - z["g_coef"]).max() <= 2e-5 * scale      # float32 autograd of the reference


@pytest.mark.parametrize("r,d", RS)
def test_ds_is_strided_slice(golden_dir, r, d):
    if not d:
        pytest.skip("no down-sampled copy at r=2 (bts_decoder.py:93-94)")
    z = _load(golden_dir, r)
    assert np.array_equal(z["out_ds"], z["out"][:, ::d, ::d])
    assert np.array_equal(c_oracle.downsample(z["out"], d), z["out_ds"])


@pytest.mark.parametrize("r,d", RS)
def test_c_vs_numpy_closed_form(r, d):
    rng = np.random.default_rng(r)
    x = (1 / (1 + np.exp(-rng.standard_normal((3, 4, 6, 3))))).astype(np.float32)
    g = rng.standard_normal((3, 4 * r, 6 * r))
    gd = rng.standard_normal((3, 4 * r // d, 6 * r // d)) if d else None
    np.testing.assert_allclose(c_oracle.lpg_forward_f64(x, r), lpg_closed.forward(x, r), rtol=1e-13)
    a, b = c_oracle.lpg_backward_f64(x, g, r, gd, d), lpg_closed.backward(x, g, r, gd, d)
    assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()


@pytest.mark.parametrize("r,d", RS)
def test_literal_torch_matches_closed_form(r, d):
    torch.manual_seed(r)
    x = torch.sigmoid(torch.randn(2, 3, 5, 3, dtype=torch.float64))
    g = torch.randn(2, 3 * r, 5 * r, 1, dtype=torch.float64)
    gd = torch.randn(2, 3 * r // d, 5 * r // d, 1, dtype=torch.float64) if d else None
    out, ds, gc = lpg_literal.lpg_fwd_bwd(x, g, r, gd, d)
    ref = c_oracle.lpg_forward_f64(x.numpy(), r)
    np.testing.assert_allclose(out.numpy()[..., 0], ref, rtol=5e-7)   # double pi + float32 constant vs fl32(pi) + exact
    gref = c_oracle.lpg_backward_f64(x.numpy(), g.numpy(), r, None if gd is None else gd.numpy(), d)
    assert np.abs(gc.numpy() - gref).max() <= 2e-6 * np.abs(gref).max()


@pytest.mark.parametrize("r", [2, 4, 8])
def test_backward_finite_difference(r):
    """float64 central differences of the closed-form forward vs the analytic backward."""
    rng = np.random.default_rng(10 + r)
    x = rng.uniform(0.1, 0.8, (1, 2, 3, 3))
    g = rng.standard_normal((1, 2 * r, 3 * r))
    ana = c_oracle.lpg_backward_f64(x, g, r)
    eps = 1e-6
    for idx in np.ndindex(x.shape):
        xp, xm = x.copy(), x.copy()
        xp[idx] += eps
        xm[idx] -= eps
        num = ((c_oracle.lpg_forward_f64(xp, r) - c_oracle.lpg_forward_f64(xm, r)) * g).sum() / (2 * eps)
        assert abs(num - ana[idx]) <= 1e-5 * max(1.0, abs(ana[idx])), (idx, num, ana[idx])


def test_axis_pairing_rows_with_n1():
    """meshgrid(width_lin, height_lin) (custom_layers.py:33-35): n1 pairs with ROWS, n2 with COLUMNS."""
    r = 4
    x = np.zeros((1, 1, 1, 3), np.float32)
    x[..., 0], x[..., 1], x[..., 2] = 0.0, 0.5, 1.0           # phi = 0 -> n2 = 0, n1 = sin(pi/6) > 0
    out = c_oracle.lpg_forward_f32(x, r)[0]
    # with n2 = 0 the patch is mirror-symmetric left/right (columns) and NOT top/bottom (rows)
    np.testing.assert_allclose(out, out[:, ::-1], rtol=1e-6)
    assert np.abs(out - out[::-1, :]).max() > 1e-2
    assert (np.diff(out[:, 0]) < 0).all()                      # a_p*n1 grows down the rows -> depth falls
    x[..., 0] = 0.25                                           # phi = pi/2 -> n1 ~ 0, n2 > 0
    out = c_oracle.lpg_forward_f32(x, r)[0]
    np.testing.assert_allclose(out, out[::-1, :], rtol=1e-5)
    assert np.abs(out - out[:, ::-1]).max() > 1e-2
    assert (np.diff(out[0, :]) < 0).all()


def test_epsilon_added_after_the_sum(golden_dir):
    """K.epsilon() is added to the summed denominator (custom_layers.py:55), no clamping: the pole
    fixture reaches den <= 0 and the reference returns negative / huge values there."""
    z = np.load(os.path.join(golden_dir, "lpg_pole_r8.npz"))
    r = int(z["upratio"])
    out64, den = c_oracle.lpg_forward_f64(z["coef"], r, return_den=True)
    assert (den <= 0).any() and (z["out"] < 0).any()
    ok = np.abs(den) > 1e-3
    np.testing.assert_allclose(c_oracle.lpg_forward_f32(z["coef"], r)[ok], z["out"][..., 0][ok], rtol=1e-3)
    np.testing.assert_allclose(out64[ok], z["out64"][..., 0][ok], rtol=1e-4)
    assert np.array_equal(np.sign(out64[ok]), np.sign(z["out"][..., 0][ok]))


def test_head_oracle_matches_torch():
    torch.manual_seed(3)
    C = 16
    f = torch.nn.functional.elu(torch.randn(2, 3, 4, C, dtype=torch.float64)).requires_grad_(True)
    w = ((torch.rand(C, 3, dtype=torch.float64) * 2 - 1) * (6 / (C + 3)) ** 0.5).requires_grad_(True)
    x = lpg_literal.reduction_head(f, w)
    gx = torch.randn_like(x)
    x.backward(gx)
    xo = c_oracle.head_forward_f64(f.detach().numpy(), w.detach().numpy())
    np.testing.assert_allclose(xo, x.detach().numpy(), rtol=1e-12)
    gf, gw = c_oracle.head_backward_f64(f.detach().numpy(), w.detach().numpy(), xo, gx.numpy())
    np.testing.assert_allclose(gf, f.grad.numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(gw, w.grad.numpy(), rtol=1e-10, atol=1e-14)


def test_decoder_fixture_wiring(golden_dir):
    """bts_decoder.py:79-99 wiring as recorded from the reference run: head = sigmoid(1x1 conv), LPG of
    the head output, channel order [phi, theta, dist]."""
    z = np.load(os.path.join(golden_dir, "decoder_small.npz"))
    for r in (8, 4, 2):
        k = z["kernel_%02d" % int(z["head%d_conv_index" % r])]
        assert k.shape[:2] == (1, 1) and k.shape[3] == 3
        coef = c_oracle.head_forward_f64(z["infer_head%d_in" % r], k[0, 0])
        np.testing.assert_allclose(coef, z["infer_head%d_out" % r], rtol=1e-12)
        depth = c_oracle.lpg_forward_f64(coef, r)
        np.testing.assert_allclose(depth, z["infer_depth_%dx%d_scaled" % (r, r)][..., 0], rtol=5e-7)


def full_size_inputs(r, B=2, H=480, W=640):
    """Same seeded inputs as tests/golden/make_golden.py:full_size_inputs (numpy Generator streams are stable)."""
    rng = np.random.default_rng(4000 + r + (0 if (H, W) == (480, 640) else H * 10000 + W))
    z = rng.standard_normal((B, H // r, W // r, 3)).astype(np.float32)
    coef = (1.0 / (1.0 + np.exp(-z.astype(np.float64)))).astype(np.float32)
    idx = np.sort(rng.choice(B * H * W, size=4096, replace=False))
    return coef, idx


@pytest.mark.parametrize("r", [8, 4, 2])
def test_full_size_pin(golden_dir, r):
    """The layer at full size (2 x 480 x 640): the oracle against 4096 sampled outputs of the UNMODIFIED
    reference layer (SURVEY 8(c) pin 4: a full-size fixture without committing the 2.4 MB maps)."""
    z = np.load(os.path.join(golden_dir, "lpg_full_size_samples.npz"))
    coef, idx = full_size_inputs(r)
    assert np.array_equal(idx, z["r%d_idx" % r])                       # same inputs, same sample
    out64, den = c_oracle.lpg_forward_f64(coef, r, return_den=True)
    good = den.reshape(-1)[idx] >= 0.05
    assert good.sum() > 4000
    np.testing.assert_allclose(out64.reshape(-1)[idx][good], z["r%d_out64" % r][good], rtol=5e-7)
    np.testing.assert_allclose(out64.reshape(-1)[idx][good], z["r%d_out" % r][good], rtol=3e-6)
    out32 = c_oracle.lpg_forward_f32(coef, r)
    np.testing.assert_allclose(out32.reshape(-1)[idx][good], z["r%d_out" % r][good], rtol=2e-6)
    assert int((out64 < 0).sum()) == int(z["r%d_n_negative" % r])      # the pole is crossed in the same places


@pytest.mark.parametrize("H,W", [(352, 1216), (416, 544)])
@pytest.mark.parametrize("r", [8, 4, 2])
def test_full_size_pin_other_shapes(golden_dir, r, H, W):
    """The same pin at KITTI Eigen 352 x 1216 and the NYU training crop 416 x 544 (tests/golden/lpg_full_size_samples_shapes.npz)."""
    z = np.load(os.path.join(golden_dir, "lpg_full_size_samples_shapes.npz"))
    key = "h%dw%d_r%d" % (H, W, r)
    coef, idx = full_size_inputs(r, 2, H, W)
    assert np.array_equal(idx, z[key + "_idx"])
    out64, den = c_oracle.lpg_forward_f64(coef, r, return_den=True)
    good = den.reshape(-1)[idx] >= 0.05
    assert good.sum() > 4000
    np.testing.assert_allclose(out64.reshape(-1)[idx][good], z[key + "_out64"][good], rtol=5e-7)
    out32 = c_oracle.lpg_forward_f32(coef, r)
    np.testing.assert_allclose(out32.reshape(-1)[idx][good], z[key + "_out"][good], rtol=2e-6)
    assert int((out64 < 0).sum()) == int(z[key + "_n_negative"])


def test_decoder_f256_fixture_kernels_regenerate(golden_dir):
    """tests/golden/decoder_f256.npz stores no kernels: oracle/decoder_fixture.regen_kernels redraws the glorot_uniform
    kernels of the recorded reference run (make_golden.py checked them bit for bit); their float64 sums pin the draws."""
    import numpy as np
    import os
    from oracle import decoder_fixture
    z = np.load(os.path.join(golden_dir, "decoder_f256.npz"))
    shapes = [tuple(int(v) for v in s) for s in z["kernel_shapes"]]
    assert len(shapes) == 25 and shapes[-1] == (3, 3, 16, 1) and shapes[-2] == (3, 3, 19, 16)       # depth conv, iconv1 (F/16 + 3 inputs)
    assert [s for s in shapes if s[3] == 3 and s[0] == 1] == [(1, 1, 64, 3), (1, 1, 64, 3), (1, 1, 32, 3)]   # reduction_8x8 / 4x4 / 2x2
    kernels = decoder_fixture.regen_kernels(shapes, int(z["seed"]))
    np.testing.assert_allclose([float(k.sum()) for k in kernels], z["kernel_sums"], rtol=0, atol=1e-9)
    for k, s in zip(kernels, shapes):
        limit = (6.0 / (s[0] * s[1] * (s[2] + s[3]))) ** 0.5
        assert float(k.abs().max()) <= limit
