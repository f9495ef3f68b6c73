"""CPU tests of the training-step host logic: the flat parameter / gradient / moment buffers, the chunked
hook-driven all-reduce under a world_size-2 gloo group, the bucket's zero_grad(set_to_none) safety, and the optimizer
oracle's schedule.  The compute on the ranks is plain torch-CPU autograd + the numpy optimizer ORACLE (checker code);
the product's update kernel needs a GPU (tests/test_trainer_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bts_fully_tf_b200 import parallel, trainer
from oracle import optim_oracle


def _toy():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1, bias=False), torch.nn.BatchNorm2d(8), torch.nn.ELU(),
                               torch.nn.Conv2d(8, 5, 3, padding=1, bias=False), torch.nn.ELU(), torch.nn.Conv2d(5, 1, 1, bias=False))


def _flat_of(model, fractions=(0.2, 0.6, 1.0)):
    params = list(reversed([p for p in model.parameters()]))
    conv_ids = [id(m.weight) for m in model.modules() if isinstance(m, torch.nn.Conv2d)]
    return trainer.FlatState(params, "cpu", fractions, channels_last_ids=conv_ids)


def test_flat_state_layout_views_and_chunks():
    model = _toy()
    before = [p.detach().clone() for p in model.parameters()]
    flat = _flat_of(model)
    for p, b in zip(model.parameters(), before):
        assert torch.equal(p.detach(), b)                                   # values survive the move into the flat buffer
        off, num = flat.slices[flat.index[id(p)]]
        assert off % 4 == 0 and p.data_ptr() == flat.param.data_ptr() + 4 * off
        assert p.grad.data_ptr() == flat.grad.data_ptr() + 4 * off
        if p.dim() == 4:
            assert p.is_contiguous(memory_format=torch.channels_last) and p.grad.stride() == p.stride()
    # reverse creation order: the LAST layer's kernel is the first slice (its gradient is produced first in backward)
    assert flat.params[0] is list(model.parameters())[-1]
    # chunks tile the buffer and respect parameter boundaries
    assert flat.chunks[0][0] == 0 and flat.chunks[-1][1] == flat.numel
    assert all(a[1] == b[0] for a, b in zip(flat.chunks, flat.chunks[1:]))
    assert sorted(flat.chunk_of) == list(range(len(flat.params)))
    # autograd accumulates in place, in the buffer
    x = torch.randn(2, 3, 6, 6)
    model(x).sum().backward()
    assert float(flat.grad.abs().sum()) > 0
    for p in model.parameters():
        assert p.grad.data_ptr() == flat.grad_view(p).data_ptr()


def test_flat_state_survives_zero_grad_set_to_none():
    model = _toy()
    flat = _flat_of(model)
    model.zero_grad(set_to_none=True)                                        # torch's default drops the views ...
    assert all(p.grad is None for p in model.parameters())
    flat.zero()                                                              # ... zero() / attach_grads() restores them
    model(torch.randn(1, 3, 4, 4)).sum().backward()
    for p in model.parameters():
        assert p.grad.data_ptr() == flat.grad_view(p).data_ptr()
    assert float(flat.grad.abs().sum()) > 0


def test_gradient_bucket_detects_detached_views():
    a = torch.nn.Parameter(torch.randn(4, 4))
    bucket = parallel.GradientBucket([a])
    a.grad = None                                                            # what optimizer.zero_grad() does by default
    with pytest.raises(RuntimeError):
        bucket.all_reduce()
    bucket.zero()
    assert a.grad.data_ptr() == bucket.view(a).data_ptr()
    bucket.all_reduce()                                                      # single process: a no-op, but the check passes


def test_poly_lr_matches_reference_formula():
    # bts_train.py:125-131 with 2 replicas: start = 2e-4, end = 0.1 * start
    start, total = 2e-4, 1000
    assert optim_oracle.poly_lr(0, start, start * 0.1, total) == np.float32(start)
    assert optim_oracle.poly_lr(total, start, start * 0.1, total) == np.float32(start * 0.1)
    assert optim_oracle.poly_lr(10 * total, start, start * 0.1, total) == np.float32(start * 0.1)        # tf.minimum(step, total)
    mid = (start - 0.1 * start) * (1 - 0.5) ** 0.9 + 0.1 * start
    assert optim_oracle.poly_lr(500, start, start * 0.1, total) == np.float32(mid)


def test_adamw_oracle_decay_branches_and_first_step():
    p = np.array([1.0, -2.0, 0.0]); g = np.array([0.5, 0.5, -0.25]); z = np.zeros(3)
    # first Adam step with zero moments: m = (1-b1) g, v = (1-b2) g^2, alpha = lr sqrt(1-b2)/(1-b1)  =>  step = lr * g/(|g| + eps*...)
    q, m, v, lr = optim_oracle.adamw_step(p, g, z, z, 0, 1e-3, epsilon=1e-3)
    alpha = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    np.testing.assert_allclose(q, p - alpha * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-3), rtol=1e-5)   # hyper-parameters enter as float32 values
    # decay: custom_optimizers.py:49-54 (l1 and l2 / l1 only / l2 only), applied with lr before the update
    for l1, l2, d in ((0.1, 0.2, 0.1 * np.sign(p) + 0.2 * p), (0.1, 0.0, 0.1 * np.sign(p)), (0.0, 0.2, 0.2 * p)):
        q2, _, _, _ = optim_oracle.adamw_step(p, np.zeros(3), z, z, 0, 1e-3, l1=l1, l2=l2)
        np.testing.assert_allclose(q2, p - np.float32(1e-3) * d, rtol=1e-12)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmp, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    parallel.init_distributed("gloo")
    model = _toy()
    flat = _flat_of(model)
    comm = trainer.ChunkedAllReduce(flat, "cpu", overlap=overlap)
    data = torch.load(os.path.join(tmp, "x.pt"))
    lo, hi = parallel.shard_range(data.shape[0], world, rank)
    order = []
    for step in range(2):                                                    # two steps: begin_step() re-arms the counters
        flat.zero()
        comm.begin_step()
        out = model(data[lo:hi])
        loss = torch.sqrt((out * out).mean())                                # per-rank, non-linear in the batch like si_log_loss
        loss.backward()
        comm.finish()
        order.append(list(comm.launch_order))
        flat.grad.mul_(1.0 / world)
    torch.save({"grad": flat.grad.clone(), "order": order}, os.path.join(tmp, "rank%d_%d.pt" % (rank, int(overlap))))
    comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_two_rank_gloo_chunked_allreduce_matches_single_process(tmp_path, overlap):
    world = 2
    torch.manual_seed(0)
    x = torch.randn(4, 3, 6, 6)
    torch.save(x, tmp_path / "x.pt")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), overlap), nprocs=world, join=True)
    r0 = torch.load(tmp_path / ("rank0_%d.pt" % int(overlap)))
    r1 = torch.load(tmp_path / ("rank1_%d.pt" % int(overlap)))
    assert torch.equal(r0["grad"], r1["grad"])                               # every rank holds the same averaged gradient
    # chunks are exchanged in the order backward completes them: chunk 0 (the tail of the network) first
    n_chunks = len(r0["order"][0])
    assert r0["order"][0] == list(range(n_chunks)) and r0["order"][1] == list(range(n_chunks))
    # single process, same definition: mean over ranks of the per-rank-loss gradients
    expect = None
    for k in range(world):
        model = _toy()
        flat = _flat_of(model)
        lo, hi = parallel.shard_range(4, world, k)
        out = model(x[lo:hi])
        torch.sqrt((out * out).mean()).backward()
        expect = flat.grad.clone() if expect is None else expect + flat.grad
    torch.testing.assert_close(r0["grad"], expect / world, rtol=1e-5, atol=1e-7)


def test_readiness_is_counted_once_per_parameter():
    """A parameter whose gradient a kernel writes into the bucket reports through writer_callback AND autograd may still run
    its (empty) accumulation hook: the second report must not launch the chunk's all-reduce one parameter early
    (the fused LPG heads did exactly that on the first 2-GPU run: chunks holding a head were exchanged too soon)."""
    model = _toy()
    flat = _flat_of(model, fractions=(1.0,))
    comm = trainer.ChunkedAllReduce(flat, "cpu", overlap=True)
    comm.begin_step()
    n = len(flat.params)
    cb = comm.writer_callback(flat.params[0])
    cb()
    cb()                                                   # duplicate report
    comm.ready(0)                                          # and the hook's
    assert comm.pending[0] == n - 1 and not comm.launched[0]
    for k in range(1, n):
        comm.ready(k)
    assert comm.pending[0] == 0 and comm.launched[0] and comm.launch_order == [0]
    comm.close()
