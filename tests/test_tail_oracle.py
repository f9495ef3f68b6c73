"""CPU tests of the decoder-tail oracle (oracle/tail_oracle.py): pinned to the UNMODIFIED reference
files bts.py and custom_eval_metrics.py run over oracle/tf_shim (fixtures tests/golden/tail_*.npz),
plus a finite-difference check of the analytic gradient and host-side checks of the reference-shaped
wrappers.  No GPU, no compute through the C ABI."""
import os
import types

import numpy as np
import pytest

from oracle import tail_oracle as T


@pytest.fixture(scope="module")
def silog(golden_dir):
    return np.load(os.path.join(golden_dir, "tail_silog.npz"))


@pytest.fixture(scope="module")
def metrics(golden_dir):
    return np.load(os.path.join(golden_dir, "tail_metrics.npz"))


@pytest.mark.parametrize("dataset", ["nyu", "kitti"])
def test_silog_oracle_matches_reference_fixture(silog, dataset):
    md, th = float(silog[dataset + "_max_depth"]), T.GT_TH[dataset]
    logit, y_true = silog[dataset + "_logit"], silog[dataset + "_y_true"]
    depth = T.depth_est(logit, md)
    np.testing.assert_allclose(depth, silog[dataset + "_f64_depth_est"], rtol=1e-14)
    np.testing.assert_allclose(T.depth_est(logit, md, np.float32), silog[dataset + "_f32_depth_est"], rtol=3e-7)
    loss, (n, m1, var) = T.si_log_loss(y_true, depth, th)
    assert n == int((y_true > th).sum()) and 0 < n < y_true.size
    np.testing.assert_allclose(loss, silog[dataset + "_f64_loss"], rtol=1e-13)
    np.testing.assert_allclose(loss, silog[dataset + "_f32_loss"], rtol=2e-6)          # the reference's float32 run
    loss32, _ = T.si_log_loss(y_true, silog[dataset + "_f32_depth_est"], th, np.float32)
    np.testing.assert_allclose(loss32, silog[dataset + "_f32_loss"], rtol=2e-6)


@pytest.mark.parametrize("dataset", ["nyu", "kitti"])
def test_silog_gradients_match_reference_autograd(silog, dataset):
    md, th = float(silog[dataset + "_max_depth"]), T.GT_TH[dataset]
    y_true, depth = silog[dataset + "_y_true"], silog[dataset + "_f64_depth_est"]
    g_depth = T.si_log_loss_grad(y_true, depth, th)
    g_logit = T.si_log_loss_grad(y_true, depth, th, max_depth=md)
    ref_gd, ref_gz = silog[dataset + "_f64_g_depth"], silog[dataset + "_f64_g_logit"]
    assert np.abs(g_depth - ref_gd).max() <= 1e-12 * np.abs(ref_gd).max()
    assert np.abs(g_logit - ref_gz).max() <= 1e-12 * np.abs(ref_gz).max()
    assert (g_depth[~(y_true > th)] == 0).all()                                          # masked pixels carry no gradient
    assert np.abs(g_logit - silog[dataset + "_f32_g_logit"]).max() <= 2e-5 * np.abs(ref_gz).max()


def test_silog_gradient_finite_difference():
    rng = np.random.default_rng(0)
    y_true = rng.uniform(0, 10, (1, 6, 7, 1))
    y_true[0, 0, :3] = 0.0
    z = rng.normal(0, 1, y_true.shape)
    md, th = 10.0, 0.1
    g = T.si_log_loss_grad(y_true, T.depth_est(z, md), th, max_depth=md)
    for idx in [(0, 0, 0, 0), (0, 2, 3, 0), (0, 5, 6, 0)]:
        e = np.zeros_like(z)
        e[idx] = 1e-6
        fd = (T.si_log_loss(y_true, T.depth_est(z + e, md), th)[0] - T.si_log_loss(y_true, T.depth_est(z - e, md), th)[0]) / 2e-6
        assert abs(fd - g[idx]) <= 1e-6 * max(1.0, abs(fd))


def test_silog_empty_mask_is_nan_like_the_reference():
    loss, (n, _, _) = T.si_log_loss(np.zeros((1, 4, 4, 1)), np.ones((1, 4, 4, 1)), 0.1)
    assert n == 0 and np.isnan(loss)


def test_metrics_oracle_matches_reference_fixture(metrics):
    lo, hi = float(metrics["min_depth_eval"]), float(metrics["max_depth_eval"])
    assert tuple(metrics["names"]) == T.METRIC_NAMES                                      # list order, custom_eval_metrics.py:88
    got = T.eval_metrics(metrics["y_true"], metrics["y_pred"], lo, hi)
    for name, v64, v32 in zip(metrics["names"], metrics["values_f64"], metrics["values_f32"]):
        # d1..d3: the reference averages a tf.float32 cast of the indicator even in its float64 run (:47,:51,:55)
        np.testing.assert_allclose(got[str(name)], v64, rtol=2e-7 if str(name) in ("d1", "d2", "d3") else 1e-12, err_msg=str(name))
        np.testing.assert_allclose(got[str(name)], v32, rtol=3e-6, err_msg=str(name))
    gt, pred = T.pre_eval(metrics["y_true"], metrics["y_pred"], lo, hi)
    assert np.isfinite(pred).all() and pred.min() >= lo and pred.max() <= hi             # NaN/inf -> max, then clipped
    assert gt.size == got["n_valid"] < metrics["y_true"].size


def test_reference_shaped_wrappers_host_logic():
    """si_log_loss_wrapper / metrics_list_factory keep the reference's names, order and assertion."""
    from bts_fully_tf_b200 import eval_metrics, losses
    assert losses.GT_TH == T.GT_TH
    with pytest.raises(AssertionError):
        losses.si_log_loss_wrapper("sunrgbd")                                            # bts.py:40
    assert losses.si_log_loss_wrapper("kitti").__name__ == "si_log_loss"
    args = types.SimpleNamespace(min_depth_eval=1e-3, max_depth_eval=80.0, garg_crop=True, eigen_crop=False, dataset="kitti")
    fns = eval_metrics.metrics_list_factory(args)
    assert [f.__name__ for f in fns] == list(T.METRIC_NAMES)


def test_depth_tail_oracle_gradient_matches_finite_differences():
    """oracle/tail_oracle.depth_tail_backward (bts_decoder.py:100-102: ELU -> Conv2D(1, 3x3, 'same')) against central
    differences of depth_tail_forward in float64, and the forward against a direct triple loop."""
    import numpy as np
    from oracle import tail_oracle as T
    rng = np.random.default_rng(5)
    B, H, W, C = 1, 4, 5, 3
    x = rng.standard_normal((B, H, W, C)) * 1.5
    w = rng.standard_normal((9, C)) * 0.3
    g = rng.standard_normal((B, H, W, 1))
    y = T.depth_tail_forward(x, w, act_in=True)
    xe = np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
    ref = np.zeros((B, H, W, 1))
    for i in range(H):
        for j in range(W):
            for ky in range(3):
                for kx in range(3):
                    ii, jj = i + ky - 1, j + kx - 1
                    if 0 <= ii < H and 0 <= jj < W:                       # padding='same': zeros of the ACTIVATED map
                        ref[0, i, j, 0] += (xe[0, ii, jj] * w[ky * 3 + kx]).sum()
    np.testing.assert_allclose(y, ref, rtol=1e-12, atol=1e-12)
    d = T.depth_tail_forward(x, w, act_in=True, max_depth=80.0)
    np.testing.assert_allclose(d, 80.0 / (1.0 + np.exp(-ref)), rtol=1e-12)
    gx, gw = T.depth_tail_backward(x, w, g)
    eps = 1e-6
    f = lambda xx, ww: float((T.depth_tail_forward(xx, ww, act_in=True) * g).sum())   # noqa: E731
    for idx in [(0, 0, 0, 0), (0, 1, 2, 1), (0, 3, 4, 2), (0, 2, 0, 1)]:
        xp, xm = x.copy(), x.copy()
        xp[idx] += eps
        xm[idx] -= eps
        assert abs((f(xp, w) - f(xm, w)) / (2 * eps) - gx[idx]) <= 1e-6 * max(1.0, abs(gx[idx]))
    for idx in [(0, 0), (4, 1), (8, 2)]:
        wp, wm = w.copy(), w.copy()
        wp[idx] += eps
        wm[idx] -= eps
        assert abs((f(x, wp) - f(x, wm)) / (2 * eps) - gw[idx]) <= 1e-6 * max(1.0, abs(gw[idx]))
