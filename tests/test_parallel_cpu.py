"""CPU tests of the data-parallel host logic: batch sharding, the flat gradient
bucket, and a world_size-2 gloo run whose averaged per-rank gradients equal the single-process
result of the same per-rank-loss definition.  Compute on the ranks is the CPU ORACLE (checker
code) -- the product kernels need a GPU; this file tests the plumbing around them."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bts_fully_tf_b200 import parallel
from oracle import c_oracle


def test_shard_range_covers_batch():
    for B in (1, 4, 31, 32, 64):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_range(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.shard_range(64, 8, 3) == (24, 32)          # BASELINE config 3: 8 per GPU at N=8
    assert parallel.shard_range(32, 8, 7) == (28, 32)          # configs 4-5: 4 per GPU, like args/train_*.txt:9


def test_shard_batch_nested():
    x = torch.arange(8 * 3).reshape(8, 3)
    out = parallel.shard_batch([x, (x + 1,)], 4, 2)
    assert torch.equal(out[0], x[4:6]) and torch.equal(out[1][0], x[4:6] + 1)


def test_scaled_learning_rate():
    assert parallel.scaled_learning_rate(1e-4, 8) == pytest.approx(8e-4)      # bts_train.py:125-126


def test_gradient_bucket_layout_and_inplace_accumulation():
    a = torch.nn.Parameter(torch.randn(1, 1, 5, 3))
    b = torch.nn.Parameter(torch.randn(7))
    frozen = torch.nn.Parameter(torch.randn(3), requires_grad=False)
    bucket = parallel.GradientBucket([a, frozen, b])
    assert bucket.numel >= 22 and all(off % 4 == 0 for off, _ in bucket.offsets.values())
    assert a.grad.data_ptr() == bucket.view(a).data_ptr() and b.grad.data_ptr() == bucket.view(b).data_ptr()
    ((a * 2).sum() + (b * 3).sum()).backward()
    assert torch.equal(bucket.view(a), torch.full_like(a, 2.0)) and torch.equal(bucket.view(b), torch.full_like(b, 3.0))
    assert a.grad.data_ptr() == bucket.view(a).data_ptr()              # autograd accumulated in place
    bucket.view(a).copy_(torch.ones_like(a))                            # a kernel writing its slice directly
    assert float(bucket.buffer.sum()) == pytest.approx(15 + 21)
    bucket.zero()
    assert float(a.grad.abs().sum()) == 0.0


def test_bind_heads_targets_bucket_slices():
    from bts_fully_tf_b200 import ReductionLPG
    heads = torch.nn.ModuleList([ReductionLPG(32, 8, 4), ReductionLPG(16, 2)])
    bucket = parallel.GradientBucket(list(heads.parameters()))
    assert bucket.bind_heads(heads) == 2
    for h in heads:
        assert h._grad_view.data_ptr() == bucket.view(h.kernel).data_ptr()
    with pytest.raises(ValueError):
        heads[0].bind_gradient_view(torch.zeros(5))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_grads(coef, feat, kern, g_full, r):
    """Per-rank 'step' on the oracle: head + LPG forward, a per-rank scalar loss
    mean(g_full * depth) (non-linear reductions such as si_log are per rank too, bts.py:37-38),
    gradient w.r.t. the head kernel."""
    x = c_oracle.head_forward_f64(feat, kern)
    n = g_full.size
    g_coef = c_oracle.lpg_backward_f64(x, g_full / n, r)
    _, g_w = c_oracle.head_backward_f64(feat, kern, x, g_coef)
    return g_w


def _worker(rank, world, port, tmp, B, r):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    parallel.init_distributed("gloo")
    data = np.load(os.path.join(tmp, "data.npz"))
    lo, hi = parallel.shard_range(B, world, rank)
    kern = torch.nn.Parameter(torch.from_numpy(data["kern"]).float().reshape(1, 1, -1, 3))
    bucket = parallel.GradientBucket([kern])
    g_w = _rank_grads(None, data["feat"][lo:hi], data["kern"], data["g_full"][lo:hi], r)
    bucket.view(kern).copy_(torch.from_numpy(g_w).float().reshape(kern.shape))     # what the fused backward kernel does on a GPU
    bucket.all_reduce(average=True)
    bucket.wait()
    np.save(os.path.join(tmp, "grad_rank%d.npy" % rank), kern.grad.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_single_process(tmp_path):
    B, h, w, C, r, world = 4, 3, 5, 8, 4, 2
    rng = np.random.default_rng(0)
    feat = rng.standard_normal((B, h, w, C))
    kern = rng.uniform(-0.5, 0.5, (C, 3))
    g_full = rng.standard_normal((B, h * r, w * r))
    np.savez(tmp_path / "data.npz", feat=feat, kern=kern, g_full=g_full)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), B, r), nprocs=world, join=True)
    g0 = np.load(tmp_path / "grad_rank0.npy")
    g1 = np.load(tmp_path / "grad_rank1.npy")
    assert np.array_equal(g0, g1)                                     # every rank holds the same averaged gradient
    # single process, same definition: mean over ranks of the per-rank-loss gradients
    expect = sum(_rank_grads(None, feat[lo:hi], kern, g_full[lo:hi], r)
                 for lo, hi in (parallel.shard_range(B, world, k) for k in range(world))) / world
    np.testing.assert_allclose(g0.reshape(C, 3), expect, rtol=1e-5, atol=1e-7)
    # and, because this loss is a mean over equal shards, it equals the un-sharded gradient too
    np.testing.assert_allclose(expect, _rank_grads(None, feat, kern, g_full, r), rtol=1e-10)
