"""The data-parallel training step of the decoder -- the reference's
`strategy = tf.distribute.MirroredStrategy(...)` + `model.fit` inner loop (bts_train.py:194-209, :150-177):

    per replica:  forward on its batch shard, si_log_loss on the shard (bts.py:27-41), backward
    exchange:     all-reduce(sum) of the gradients of the trainable decoder variables, x 1/N  (the ONLY collective)
    update:       AdamW (custom_optimizers.py:47-59 over Keras Adam), lr = poly_decay(step) x N (bts_train.py:125-131)

Built for one process per GPU (torchrun) with NCCL over NVLink:

  * FlatState: parameters, gradients and both Adam moments live in FOUR flat float32 buffers with one layout, ordered in
    REVERSE creation order -- the order in which backward produces the gradients (the decoder tail first, block5 last).
    `p.data` / `p.grad` are views; convolution kernels are laid out channels_last inside their slice, which is what cuDNN's
    NHWC kernels read and write, so no per-step weight re-layout and a dense in-place gradient accumulation.
  * The flat gradient buffer is split into `chunks` contiguous pieces.  A chunk's all-reduce is launched on a side stream
    the moment its last gradient has been produced (autograd post-accumulate hooks; the fused LPG heads, which write their
    g_kernel straight into the bucket, report through ReductionLPG's on_written callback), so the exchange of the tail
    chunks overlaps the rest of backward; only the last chunk's all-reduce is exposed.
  * When the process group is NCCL the gradient buffer is allocated from NCCL's allocator (ncclMemAlloc) and registered with
    the communicator (ncclCommRegister) through torch's MemPool hooks, which lets NCCL run the NVLS (in-switch) all-reduce
    zero-copy on the user buffer.
  * The update is ONE hand-written kernel per chunk (ops.adam_step -> csrc/optim_kernels.cuh): it reads the summed
    gradient, applies 1/N, the decoupled decay and Adam, and zeroes the gradient for the next step; step counter and
    learning-rate schedule are device-resident.
  * The whole step -- forward, loss, backward, the overlapped all-reduces, the updates -- is captured into ONE CUDA graph
    per rank after `warmup` eager steps (cuDNN autotuning happens there); a replay is a single launch, which is what makes
    small per-GPU batches (4 images per GPU at N = 8, the reference's own setting, args/train_*.txt:9) scale.

BatchNormalization stays per replica (the reference does not sync it, bts_decoder.py:27,33) and the loss is per replica.
"""
import torch
import torch.distributed as dist

from . import ops
from .layers import ReductionLPG
from .parallel import scaled_learning_rate


def _world():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


class FlatState:
    """Flat float32 buffers (param, grad, adam m, adam v) over `params` in the given order, split into chunks."""

    def __init__(self, params, device, chunk_fractions=(0.05, 0.3, 0.5, 1.0), grad_allocator=None, channels_last_ids=()):
        self.params = [p for p in params if p.requires_grad]
        self.channels_last_ids = set(channels_last_ids)       # ids of OIHW convolution kernels (kept channels_last in memory)
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("FlatState holds float32 parameters")
        self.slices, n = [], 0
        for p in self.params:
            n = (n + 3) // 4 * 4                                   # 16-byte aligned slices (vector accesses in the kernels)
            self.slices.append((n, p.numel()))
            n += p.numel()
        self.numel = (n + 3) // 4 * 4
        self.index = {id(p): k for k, p in enumerate(self.params)}
        self.param = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.grad = grad_allocator(self.numel) if grad_allocator is not None else torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.m = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.v = torch.zeros(self.numel, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p in self.params:
                view = self._view(self.param, p)
                view.copy_(p.detach())
                p.data = view
        self.attach_grads()
        # chunk boundaries at parameter boundaries, by cumulative size
        self.chunks, start_k, lo = [], 0, 0                       # (lo, hi, first param index, last param index)
        total = float(max(self.numel, 1))
        fr = list(chunk_fractions)
        for k, (off, num) in enumerate(self.slices):
            hi = self.slices[k + 1][0] if k + 1 < len(self.slices) else self.numel
            if hi / total >= fr[len(self.chunks)] - 1e-12 or k == len(self.slices) - 1:
                self.chunks.append((lo, hi, start_k, k))
                lo, start_k = hi, k + 1
                if len(self.chunks) == len(fr):
                    break
        if self.chunks and self.chunks[-1][1] != self.numel:       # fractions that do not reach 1.0: the rest is the last chunk
            lo_, _, s_, _ = self.chunks[-1]
            self.chunks[-1] = (lo_, self.numel, s_, len(self.slices) - 1)
        self.chunk_of = {}
        for c, (_, _, a, b) in enumerate(self.chunks):
            for k in range(a, b + 1):
                self.chunk_of[k] = c

    def _view(self, flat, p):
        off, num = self.slices[self.index[id(p)]]
        seg = flat[off:off + num]
        if p.dim() == 4 and id(p) in self.channels_last_ids:      # OIHW logical shape, channels_last (O,H,W,I) in memory
            o, i, h, w = p.shape
            return seg.view(o, h, w, i).permute(0, 3, 1, 2)
        return seg.view(p.shape)

    def grad_view(self, p):
        return self._view(self.grad, p)

    def flat_grad_slice(self, p):
        """The parameter's slice of the gradient buffer as a flat contiguous tensor (what the fused heads write)."""
        off, num = self.slices[self.index[id(p)]]
        return self.grad[off:off + num]

    def attach_grads(self):
        """(Re-)point every p.grad at its slice.  `zero_grad(set_to_none=True)` (torch's default) drops these views;
        call this (or zero()) afterwards, or autograd allocates gradients outside the bucket."""
        for p in self.params:
            g = p.grad
            if g is None or g.data_ptr() != self.grad.data_ptr() + 4 * self.slices[self.index[id(p)]][0]:
                p.grad = self.grad_view(p)

    def zero(self):
        self.grad.zero_()
        self.attach_grads()

    def nbytes(self):
        return self.numel * 4


def _nccl_backend(device):
    try:
        return dist.group.WORLD._get_backend(torch.device(device))
    except Exception:  # noqa: BLE001
        return None


class ChunkedAllReduce:
    """Launches the all-reduce(sum) of each chunk of `flat.grad` as soon as the chunk's last gradient has been produced.

    Readiness comes from autograd's post-accumulate hooks on every parameter and, for parameters whose gradient a kernel
    writes straight into the bucket (the fused LPG heads), from the `on_written` callback handed out by `writer_callback`.
    On CUDA the collective runs on a side stream behind an event (overlapping the rest of backward) and `wait(c)` makes the
    current stream wait for chunk c; on CPU (gloo, used by the tests of this logic) it runs inline."""

    def __init__(self, flat, device, overlap=True, producer_stream=None):
        self.flat, self.device = flat, torch.device(device)
        self.producer = producer_stream    # the stream backward runs on (the hooks fire on autograd's worker thread, whose
        self.hook_streams = set()          # "current stream" need not be it); hook_streams: what the hooks saw (diagnostics)
        self.world = _world()
        self.cuda = self.device.type == "cuda"
        self.overlap = bool(overlap)
        self.enabled = True                # False: skip the exchange (measurement of the un-communicated step)
        self.stream = torch.cuda.Stream(self.device) if (self.cuda and self.world > 1) else None
        self.events = [torch.cuda.Event() for _ in flat.chunks] if self.cuda else []
        self.pending = [0] * len(flat.chunks)
        self.launched = [False] * len(flat.chunks)
        self.launch_order = []             # chunk ids in the order their all-reduce was issued (last step)
        self.done = [False] * len(flat.params)
        self.hooks = [p.register_post_accumulate_grad_hook(lambda _p, k=k: self.ready(k)) for k, p in enumerate(flat.params)]

    def writer_callback(self, p):
        k = self.flat.index[id(p)]
        return lambda: self.ready(k)

    def begin_step(self):
        for c, (_, _, a, b) in enumerate(self.flat.chunks):
            self.pending[c] = b - a + 1
            self.launched[c] = False
        self.done = [False] * len(self.flat.params)
        self.launch_order = []

    def ready(self, k):
        """Parameter k's gradient is complete.  Counted once per step: a parameter whose gradient a kernel writes into the
        bucket reports through writer_callback, and autograd may still run its (empty) accumulation hook afterwards."""
        if self.done[k]:
            return
        self.done[k] = True
        c = self.flat.chunk_of[k]
        self.pending[c] -= 1
        if self.pending[c] == 0 and self.overlap:
            self.launch(c)

    def launch(self, c):
        if self.launched[c]:
            return
        self.launched[c] = True
        self.launch_order.append(c)
        if self.world == 1 or not self.enabled:
            return
        lo, hi, _, _ = self.flat.chunks[c]
        if not self.cuda:
            dist.all_reduce(self.flat.grad[lo:hi], op=dist.ReduceOp.SUM)
            return
        cur = torch.cuda.current_stream(self.device)
        self.hook_streams.add(cur.cuda_stream)
        if self.producer is not None:
            self.stream.wait_stream(self.producer)                 # the chunk's gradients are enqueued on the producer stream
        if self.producer is None or cur.cuda_stream != self.producer.cuda_stream:
            self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            work = dist.all_reduce(self.flat.grad[lo:hi], op=dist.ReduceOp.SUM, async_op=True)
            work.wait()                                            # the side stream now depends on the collective
            self.events[c].record(self.stream)

    def finish(self):
        """After backward: issue whatever the hooks did not (overlap off, parameters that received no gradient)."""
        for c in range(len(self.flat.chunks)):
            self.launch(c)

    def wait(self, c):
        if self.cuda and self.world > 1 and self.enabled:
            torch.cuda.current_stream(self.device).wait_event(self.events[c])

    def close(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []


class DataParallelStep:
    """step() = one training iteration of the decoder on this rank's shard (see the module docstring).

        engine = DataParallelStep(decoder, feats, gt, dataset="nyu", base_lr=1e-4, total_steps=...).warmup_and_capture()
        loss = engine.step()                 # static input buffers: engine.feats / engine.gt (copy the next batch into them)

    base_lr is the reference's per-replica --learning_rate; it is multiplied by the world size as in bts_train.py:125.
    """

    def __init__(self, decoder, feats, gt, dataset="nyu", base_lr=1e-4, end_lr=-1.0, total_steps=0, adam_eps=1e-3,
                 chunk_fractions=(0.05, 0.3, 0.5, 1.0), overlap=True, use_graph=True, register_nccl=True, decay=(0.0, 0.0)):
        self.decoder, self.dataset = decoder, dataset
        self.feats = [f.detach() for f in feats]                  # static input buffers (the graph reads these addresses)
        self.gt = gt.detach()
        self.device = self.gt.device
        self.world = _world()
        self.use_graph = bool(use_graph)
        self.registered = False
        self._pool = None
        decoder.train(True)
        params = list(reversed([p for p in decoder.parameters() if p.requires_grad]))      # reverse creation order
        conv_ids = [id(m.weight) for m in decoder.modules() if isinstance(m, torch.nn.Conv2d)]
        self.flat = FlatState(params, self.device, chunk_fractions, grad_allocator=self._alloc_grad if (register_nccl and self.world > 1) else None,
                              channels_last_ids=conv_ids)
        start_lr = scaled_learning_rate(base_lr, self.world)
        lr_end = end_lr * self.world if end_lr > 0 else None
        self.cfg = ops.adam_config(start_lr, lr_end, total_steps, epsilon=adam_eps, l1=decay[0], l2=decay[1], grad_scale=1.0 / self.world,
                                   zero_grad=True)
        self.state = ops.adam_state(self.device)
        self.loss = torch.zeros((), dtype=torch.float32, device=self.device)
        # ONE stream for every step, eager or captured.  Autograd pins a parameter's AccumulateGrad node (and with it the stream
        # it runs on) the first time it is created -- here, when the hooks below are registered -- and keeps it alive for as
        # long as a hook hangs on it; a node born on the default stream would drag the legacy stream into a later capture on
        # another stream and invalidate it.  So: hooks are registered, warm-up steps run and the graph is captured on this stream.
        self.stream = torch.cuda.Stream(self.device)
        with torch.cuda.stream(self.stream):
            self.comm = ChunkedAllReduce(self.flat, self.device, overlap=overlap, producer_stream=self.stream)
        for m in self.decoder.modules():                           # the fused heads write g_kernel straight into the bucket
            if isinstance(m, ReductionLPG) and id(m.kernel) in self.flat.index:
                m.bind_gradient_view(self.flat.flat_grad_slice(m.kernel), on_written=self.comm.writer_callback(m.kernel))
        self.graph = None
        self.steps_done = 0

    # ---- NCCL-registered gradient buffer ------------------------------------------------------
    def _alloc_grad(self, numel):
        be = _nccl_backend(self.device)
        if be is not None and hasattr(be, "mem_allocator") and hasattr(torch.cuda, "MemPool"):
            try:
                pool = torch.cuda.MemPool(be.mem_allocator)
                with torch.cuda.use_mem_pool(pool):
                    buf = torch.zeros(numel, dtype=torch.float32, device=self.device)
                be.register_mem_pool(pool)
                self._pool, self.registered = pool, True
                return buf
            except Exception:  # noqa: BLE001  (no ncclMemAlloc support: a plain buffer works, unregistered)
                self._pool, self.registered = None, False
        return torch.zeros(numel, dtype=torch.float32, device=self.device)

    # ---- one step -----------------------------------------------------------------------------
    def _step_body(self):
        self.comm.begin_step()
        _, loss = self.decoder.forward_loss(self.feats, self.gt, self.dataset)
        loss.backward()
        self.comm.finish()
        last = len(self.flat.chunks) - 1
        for c, (lo, hi, _, _) in enumerate(self.flat.chunks):
            self.comm.wait(c)                                      # chunk c is summed; later chunks may still be in flight
            ops.adam_step(self.flat.param[lo:hi], self.flat.grad[lo:hi], self.flat.m[lo:hi], self.flat.v[lo:hi], self.state, self.cfg,
                          advance=(c == last))
        self.loss.copy_(loss.detach())
        return self.loss

    def _eager_step(self):
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self._step_body()
        cur.wait_stream(self.stream)

    def capture(self):
        """Capture _step_body into one CUDA graph (call after a few eager steps: cuDNN autotuning, workspace allocation)."""
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(self.stream):
            # thread_local: NCCL's watchdog thread and the autograd worker threads stay free to make CUDA calls (event
            # queries, allocator growth) that a "global" capture would turn into capture-invalidating errors
            with torch.cuda.graph(graph, stream=self.stream, capture_error_mode="thread_local"):
                self._step_body()
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        self.graph = graph
        return graph

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._eager_step()
        self.steps_done += 1
        return self.loss

    def warmup_and_capture(self, warmup=3):
        for _ in range(max(1, warmup)):
            self._eager_step()
            self.steps_done += 1
        torch.cuda.synchronize(self.device)
        if self.use_graph:
            self.capture()
        return self

    def local_gradients(self):
        """Diagnostics / tests: this rank's gradient of its per-shard loss, with no exchange and no update (a copy of the flat
        gradient buffer in the bucket's layout; the bucket is left zeroed).  Returns (flat gradient, loss)."""
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            enabled, self.comm.enabled = self.comm.enabled, False
            self.comm.begin_step()
            _, loss = self.decoder.forward_loss(self.feats, self.gt, self.dataset)
            loss.backward()
            self.comm.finish()
            self.comm.enabled = enabled
            g = self.flat.grad.clone()
            self.flat.zero()
            loss = loss.detach().clone()
        cur.wait_stream(self.stream)
        return g, loss

    def reduced_gradients(self):
        """Diagnostics / tests: forward, backward and the (overlapped, chunked) exchange of one step WITHOUT the update.
        Returns (a copy of the summed flat gradient, loss); the bucket is left zeroed."""
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.comm.begin_step()
            _, loss = self.decoder.forward_loss(self.feats, self.gt, self.dataset)
            loss.backward()
            self.comm.finish()
            for c in range(len(self.flat.chunks)):
                self.comm.wait(c)
            g = self.flat.grad.clone()
            self.flat.zero()
            loss = loss.detach().clone()
        cur.wait_stream(self.stream)
        return g, loss

    def learning_rate(self):
        """lr of the last completed update (device -> host read)."""
        return float(self.state[1].item())

    def completed_updates(self):
        return int(self.state.view(torch.int32)[0].item())

    def close(self):
        self.comm.close()
        if self._pool is not None:
            be = _nccl_backend(self.device)
            try:
                be.deregister_mem_pool(self._pool)
            except Exception:  # noqa: BLE001
                pass
            self._pool = None
