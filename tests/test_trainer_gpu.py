"""GPU tests of the training step engine (trainer.DataParallelStep): single-GPU parity against the optimizer oracle, CUDA
graph == eager, and (when the box has >= 2 GPUs) the N-rank NCCL run of tests/dp_worker.py."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from bts_fully_tf_b200 import trainer
from bts_fully_tf_b200.decoder import BtsDecoder
from oracle import optim_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _restore_tf32():
    """Some tests here switch the library's TF32 convolutions off; leave the process as it was found."""
    old = torch.backends.cudnn.allow_tf32
    yield
    torch.backends.cudnn.allow_tf32 = old


def _setup(seed=0, F=256, B=2, H=64, W=96):
    chans = [64, 8, 8, 16, 24]
    torch.manual_seed(seed)
    dec = BtsDecoder(chans, 10.0, num_filters=F).to(DEV)
    feats = [torch.relu(torch.randn(B, H // s, W // s, c, device=DEV)) for s, c in zip((32, 2, 4, 8, 16), chans)]
    gt = torch.rand(B, H, W, 1, device=DEV) * 10.0
    return dec, feats, gt


def test_single_gpu_step_matches_autograd_plus_oracle():
    torch.backends.cudnn.allow_tf32 = False
    dec, feats, gt = _setup()
    eng = trainer.DataParallelStep(dec, feats, gt, dataset="nyu", base_lr=1e-3, total_steps=50, use_graph=False)
    flat = eng.flat
    p0 = flat.param.double().cpu().numpy()
    # gradient of the same step by autograd alone
    g, loss = eng.local_gradients()
    g = g.double().cpu().numpy()
    assert np.abs(g).max() > 0
    loss2 = eng.step()
    torch.cuda.synchronize()
    assert float(loss2) == pytest.approx(float(loss), rel=1e-5)
    z = np.zeros_like(p0)
    exp_p, exp_m, exp_v, lr = optim_oracle.adamw_step(p0, g, z, z, 0, 1e-3, total_steps=50, epsilon=1e-3)
    step_size = np.abs(exp_p - p0).max()
    assert np.abs(flat.param.double().cpu().numpy() - exp_p).max() <= 0.02 * step_size
    np.testing.assert_allclose(flat.m.cpu().numpy(), exp_m, rtol=1e-3, atol=1e-3 * np.abs(exp_m).max())
    assert eng.completed_updates() == 1 and eng.learning_rate() == float(lr)
    assert float(flat.grad.abs().sum()) == 0.0                       # consumed and zeroed by the update kernel
    # every parameter still aliases the flat buffers, heads included (they write g_kernel straight into the bucket)
    for p in dec.parameters():
        off, _ = flat.slices[flat.index[id(p)]]
        assert p.data_ptr() == flat.param.data_ptr() + 4 * off and p.grad.data_ptr() == flat.grad.data_ptr() + 4 * off
    eng.close()


def test_cuda_graph_step_equals_eager_step():
    torch.backends.cudnn.allow_tf32 = False
    losses, params = [], []
    for use_graph in (False, True):
        dec, feats, gt = _setup(seed=1)
        eng = trainer.DataParallelStep(dec, feats, gt, dataset="nyu", base_lr=1e-3, use_graph=use_graph)
        if use_graph:
            eng.capture()                                              # capture without moving the weights first
        ls = []
        for _ in range(3):
            ls.append(float(eng.step()))
        torch.cuda.synchronize()
        assert eng.completed_updates() == 3
        losses.append(ls)
        params.append(eng.flat.param.clone())
        eng.close()
    # same trajectory: the loss of step k depends on the weights after k-1 updates
    np.testing.assert_allclose(losses[0], losses[1], rtol=2e-3)
    assert losses[0][0] != losses[0][2]                                # and the weights did move
    scale = params[0].abs().max()
    assert float((params[0] - params[1]).abs().max()) <= 2e-3 * float(scale)


def test_cuda_graph_step_with_tensor_core_weight_gradients():
    """The captured step with torch's TF32 switch on -- the tcgen05 weight gradients (TMA tensor maps as kernel parameters) and the
    training-mode glue inside the graph -- follows the eager trajectory."""
    from bts_fully_tf_b200 import ops
    torch.backends.cudnn.allow_tf32 = True
    calls = []
    real = ops.conv3x3_wgrad
    ops.conv3x3_wgrad = lambda x, g, out=None: (calls.append(tuple(x.shape)), real(x, g, out))[1]
    try:
        losses, params = [], []
        for use_graph in (False, True):
            dec, feats, gt = _setup(seed=3, F=128, B=4, H=128, W=256)
            eng = trainer.DataParallelStep(dec, feats, gt, dataset="nyu", base_lr=1e-3, use_graph=use_graph)
            if use_graph:
                eng.capture()
            ls = [float(eng.step()) for _ in range(3)]
            torch.cuda.synchronize()
            assert eng.completed_updates() == 3
            losses.append(ls)
            params.append(eng.flat.param.clone())
            eng.close()
    finally:
        ops.conv3x3_wgrad = real
    assert len(calls) >= 6, calls                                       # several layers per step, eager steps and the capture
    np.testing.assert_allclose(losses[0], losses[1], rtol=5e-3)
    scale = params[0].abs().max()
    assert float((params[0] - params[1]).abs().max()) <= 5e-3 * float(scale)


def test_new_batches_flow_through_static_buffers():
    dec, feats, gt = _setup(seed=2)
    eng = trainer.DataParallelStep(dec, feats, gt, dataset="nyu", base_lr=1e-4).warmup_and_capture(2)
    l1 = float(eng.step())
    eng.gt.copy_(torch.rand_like(eng.gt) * 10.0)                      # next batch: copy into the graph's input buffers
    for f in eng.feats:
        f.copy_(torch.relu(torch.randn_like(f)))
    l2 = float(eng.step())
    torch.cuda.synchronize()
    assert np.isfinite(l1) and np.isfinite(l2) and l1 != l2
    eng.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_multi_rank_nccl_step_matches_reference_definition():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert out.returncode == 0 and "DP_WORKER_OK" in out.stdout, out.stdout[-4000:]
