"""GPU parity of the fused optimizer step (btslpg_adam_step) against the numpy oracle (oracle/optim_oracle.py:
custom_optimizers.py:47-59 over the published Keras / ResourceApplyAdam update, schedule of bts_train.py:125-131), and of the
uint16 depth image of bts_predict.py:140-141.  All calls go through the C ABI."""
import numpy as np
import pytest
import torch

from bts_fully_tf_b200 import ops
from oracle import optim_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# Tolerance: float32 arithmetic in the kernel against the float64 oracle.  One update moves a weight by <= alpha ~ lr, and
# the kernel's relative error on that move is a few float32 ulps, so |p - p_ref| <= 1e-6 * max(|p|, lr) holds with margin.
RTOL = 2e-6


def _buffers(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(n, generator=g)
    grad = torch.randn(n, generator=g) * 0.1
    m = torch.randn(n, generator=g) * 0.01
    v = torch.rand(n, generator=g) * 1e-3
    return p, grad, m, v


@pytest.mark.parametrize("n", [4, 1024, 4099, 1 << 20])
@pytest.mark.parametrize("decay", [(0.0, 0.0), (0.0, 1e-4), (1e-3, 0.0), (1e-3, 1e-4)])
def test_adam_step_matches_oracle(n, decay):
    p, grad, m, v = _buffers(n)
    state = ops.adam_state(DEV)
    cfg = ops.adam_config(8e-4, total_steps=100, l1=decay[0], l2=decay[1], grad_scale=0.125, zero_grad=False)
    dp, dg, dm, dv = (t.to(DEV) for t in (p, grad, m, v))
    rp, rm, rv = p.numpy().astype(np.float64), m.numpy().astype(np.float64), v.numpy().astype(np.float64)
    for step in range(3):
        ops.adam_step(dp, dg, dm, dv, state, cfg)
        rp, rm, rv, lr = optim_oracle.adamw_step(rp, grad.numpy(), rm, rv, step, 8e-4, total_steps=100, l1=decay[0], l2=decay[1], grad_scale=0.125)
        torch.cuda.synchronize()
        assert int(state.view(torch.int32)[0]) == step + 1
        assert float(state[1]) == float(lr)                                  # the schedule, bit for bit (float32 cast of the float64 formula)
        np.testing.assert_allclose(dp.cpu().numpy(), rp, rtol=RTOL, atol=RTOL * 2.5e-3)     # floor: RTOL of the largest single update
        # the moments are sums of terms of the gradient's scale (0.1 * 0.125, squared for v): entries that cancel towards zero
        # keep an absolute error of a float32 ulp of that scale, so the absolute floor is 1e-6 of the largest entry
        np.testing.assert_allclose(dm.cpu().numpy(), rm, rtol=RTOL, atol=1e-6 * np.abs(rm).max())
        np.testing.assert_allclose(dv.cpu().numpy(), rv, rtol=RTOL, atol=1e-6 * np.abs(rv).max())
    assert torch.equal(dg.cpu(), grad)                                       # zero_grad off: the gradient is left alone


def test_adam_step_zeroes_gradient_and_chunks_share_one_step():
    n = 8192
    p, grad, m, v = _buffers(n, seed=1)
    cfg = ops.adam_config(1e-3)
    # whole buffer in one call ...
    a = [t.clone().to(DEV) for t in (p, grad, m, v)]
    sa = ops.adam_state(DEV)
    ops.adam_step(*a, sa, cfg)
    # ... equals two chunk calls of which only the last advances the step counter
    b = [t.clone().to(DEV) for t in (p, grad, m, v)]
    sb = ops.adam_state(DEV)
    ops.adam_step(*[t[:4096] for t in b], sb, cfg, advance=False)
    ops.adam_step(*[t[4096:] for t in b], sb, cfg, advance=True)
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert int(sa.view(torch.int32)[0]) == 1 and int(sb.view(torch.int32)[0]) == 1
    assert float(a[1].abs().sum()) == 0.0                                    # consumed gradient zeroed for the next step


def test_adam_step_in_cuda_graph_advances_on_device():
    n = 4096
    p, grad, m, v = _buffers(n, seed=2)
    dp, dg, dm, dv = (t.to(DEV) for t in (p, grad, m, v))
    state = ops.adam_state(DEV)
    cfg = ops.adam_config(2e-4, total_steps=10, zero_grad=False)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            ops.adam_step(dp, dg, dm, dv, state, cfg)
    torch.cuda.current_stream().wait_stream(side)
    rp, rm, rv = p.numpy().astype(np.float64), m.numpy().astype(np.float64), v.numpy().astype(np.float64)
    for step in range(12):                                                   # past total_steps: the schedule clamps (tf.minimum)
        graph.replay()
        rp, rm, rv, lr = optim_oracle.adamw_step(rp, grad.numpy(), rm, rv, step, 2e-4, total_steps=10)
    torch.cuda.synchronize()
    assert int(state.view(torch.int32)[0]) == 12
    assert float(state[1]) == float(lr) == float(np.float32(2e-5))
    np.testing.assert_allclose(dp.cpu().numpy(), rp, rtol=1e-5, atol=1e-8)


def test_adam_step_rejects_bad_arguments():
    p, grad, m, v = (t.to(DEV) for t in _buffers(64))
    state = ops.adam_state(DEV)
    with pytest.raises(ValueError):
        ops.adam_step(p, grad[:32], m, v, state, ops.adam_config(1e-3))      # element counts differ
    with pytest.raises(ValueError):
        ops.adam_step(p.cpu(), grad, m, v, state, ops.adam_config(1e-3))     # host tensor: no CPU fallback
    with pytest.raises(ValueError):
        ops.adam_step(p[1:33], grad[1:33], m[1:33], v[1:33], state, ops.adam_config(1e-3))   # misaligned slice


@pytest.mark.parametrize("shape", [(2, 48, 64, 1), (1, 7, 9, 1), (3, 480, 640, 1)])
@pytest.mark.parametrize("max_depth", [10.0, 80.0])
def test_png16_matches_reference_line(shape, max_depth):
    g = torch.Generator().manual_seed(5)
    depth = torch.rand(shape, generator=g) * max_depth
    flat = depth.view(-1)
    flat[0], flat[1], flat[2] = max_depth, 0.0, max_depth * (1 - 2 ** -20)   # wraps to 0 like numpy / zero / just below the wrap
    if flat.numel() > 8:
        flat[3], flat[4], flat[5] = float("nan"), -1.0, 2 * max_depth
    png, metrics = ops.eval_metrics_png16(depth.to(DEV), max_depth)
    torch.cuda.synchronize()
    assert metrics is None and png.dtype == torch.uint16 and tuple(png.shape) == shape
    ref = optim_oracle.png16(depth.numpy(), max_depth)
    np.testing.assert_array_equal(png.cpu().numpy(), ref)


def test_png16_fused_with_metrics_equals_separate_calls():
    g = torch.Generator().manual_seed(6)
    shape = (2, 96, 128, 1)
    y_true = (torch.rand(shape, generator=g) * 12.0).to(DEV)
    y_pred = (torch.rand(shape, generator=g) * 10.0).to(DEV)
    png, fused = ops.eval_metrics_png16(y_pred, 10.0, y_true=y_true, min_depth_eval=1e-3, max_depth_eval=10.0)
    alone = ops.eval_metrics(y_true, y_pred, 1e-3, 10.0)
    png_alone, _ = ops.eval_metrics_png16(y_pred, 10.0)
    torch.cuda.synchronize()
    assert torch.equal(fused, alone) and torch.equal(png, png_alone)
    np.testing.assert_array_equal(png.cpu().numpy(), optim_oracle.png16(y_pred.cpu().numpy(), 10.0))
