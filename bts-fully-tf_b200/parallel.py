"""Data parallelism for the LPG path -- the reference's only strategy.

Reference: `tf.distribute.MirroredStrategy` (bts_train.py:194-209): the global batch is split over N
replicas, every replica runs the same graph, the gradients of all trainable variables are
all-reduced (NCCL) once per step, un-synced BatchNorm, learning rate scaled by N (bts_train.py:125).

Here: one process per GPU (torchrun), `torch.distributed` for the plumbing.
  * forward / inference: batch shards, NO communication (every hot-path op is per-sample);
  * training: the decoder's gradients live in ONE flat float32 bucket; the fused head backward
    kernel writes its g_kernel straight into its slice of the bucket (no staging copy), torch's
    conv gradients are views of the same bucket, and a single all-reduce(sum) of the bucket runs
    on a side stream as soon as backward has finished producing it; the 1/N of the mean is folded
    into the same pass.
The payload (81 MB for the densenet161 decoder, 33 MB for resnet50) is far too small for a
hand-written NVLink kernel to beat NCCL's NVLS all-reduce, so NCCL is used as is.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_distributed(backend=None):
    """Initialise torch.distributed from the environment when WORLD_SIZE > 1.  Returns (rank, local_rank, world)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kwargs)
    return rank, local_rank, world


def shard_range(global_batch, world, rank):
    """Samples [lo, hi) of rank `rank`: contiguous, sizes differ by at most one (reference: equal
    per-replica batches, bts_train.py:218)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, world, rank):
    """Slice every tensor of a (nested) list along dim 0 to this rank's shard."""
    if isinstance(tensors, (list, tuple)):
        return type(tensors)(shard_batch(t, world, rank) for t in tensors)
    lo, hi = shard_range(tensors.shape[0], world, rank)
    return tensors[lo:hi]


def scaled_learning_rate(base_lr, world):
    """bts_train.py:125-126: learning rate multiplied by the number of replicas."""
    return base_lr * world


class GradientBucket:
    """One flat float32 buffer holding the gradients of `params` (in the given order).

    `.grad` of every parameter is made a view of the buffer, so autograd accumulates in place;
    `view(p)` returns the slice for kernels that write a gradient directly (the fused head's
    g_kernel -- it OVERWRITES its slice, so the bucket supports one backward per `zero()`, not gradient
    accumulation over several).  `all_reduce()` sums the buffer over ranks on a side stream and scales by 1/world
    (Keras averages the per-replica losses), `wait()` joins that stream.  Use `zero()` instead of
    `optimizer.zero_grad()` (whose default set_to_none=True would detach the views).  The overlapped, CUDA-graph
    training step is trainer.DataParallelStep; this class is the simple blocking form.
    """

    def __init__(self, params, device=None):
        self.params = [p for p in params if p.requires_grad]
        device = device or (self.params[0].device if self.params else "cpu")
        self.offsets, n = {}, 0
        for p in self.params:
            n = (n + 3) // 4 * 4                       # 16-byte aligned slices (vector stores in the kernels)
            self.offsets[id(p)] = (n, p.numel())
            n += p.numel()
        self.numel = n
        self.buffer = torch.zeros(max(n, 1), dtype=torch.float32, device=device)
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("GradientBucket holds float32 gradients")
            p.grad = self.view(p)
        self._stream = torch.cuda.Stream(device) if self.buffer.is_cuda else None
        self._work = None

    def view(self, p):
        off, num = self.offsets[id(p)]
        return self.buffer[off:off + num].view(p.shape)

    def bind_heads(self, module):
        """Point every ReductionLPG head under `module` at its slice of the bucket: the fused backward
        kernel then writes g_kernel there itself (the compute -> collective hand-off of SURVEY 8(e))."""
        from .layers import ReductionLPG
        n = 0
        for m in module.modules():
            if isinstance(m, ReductionLPG) and id(m.kernel) in self.offsets:
                m.bind_gradient_view(self.view(m.kernel))
                n += 1
        return n

    def attach(self):
        """(Re-)point every p.grad at its slice of the buffer.  torch's default `zero_grad(set_to_none=True)` drops the
        views, after which autograd would allocate gradients OUTSIDE the bucket and the all-reduce would exchange stale
        zeros: use `bucket.zero()` (which re-attaches) instead of `zero_grad()`, or call this after it."""
        for p in self.params:
            off, _ = self.offsets[id(p)]
            if p.grad is None or p.grad.data_ptr() != self.buffer.data_ptr() + 4 * off:
                p.grad = self.view(p)

    def zero(self):
        self.buffer.zero_()
        self.attach()

    def nbytes(self):
        return self.numel * 4

    def all_reduce(self, average=True):
        """Launch the exchange step; overlappable with whatever the caller enqueues next."""
        for p in self.params:                                   # a detached view means the gradients are not in the buffer
            off, _ = self.offsets[id(p)]
            if p.grad is None or p.grad.data_ptr() != self.buffer.data_ptr() + 4 * off:
                raise RuntimeError("GradientBucket: a parameter's .grad no longer aliases the bucket (zero_grad(set_to_none=True)?); "
                                   "use bucket.zero() / bucket.attach()")
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream(self.buffer.device))
            with torch.cuda.stream(self._stream):
                dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM)
                if average:
                    self.buffer.mul_(1.0 / world)
        else:
            dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM)
            if average:
                self.buffer.mul_(1.0 / world)

    def wait(self):
        if self._stream is not None:
            torch.cuda.current_stream(self.buffer.device).wait_stream(self._stream)


def allreduce_mean_(tensor):
    """In-place mean over ranks (used for scalar metrics; max-over-ranks timing uses ReduceOp.MAX)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
        tensor.mul_(1.0 / dist.get_world_size())
    return tensor
