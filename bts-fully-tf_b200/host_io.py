"""Host-resident tensors -> LPG kernels -> host-resident results.

The kernels only accept device tensors (no CPU fallback).  Callers whose coefficient maps and
gradients live in host memory (a tf.data / numpy pipeline, or the bench's end-to-end leg) use this
pipeline: pinned host buffers, host->device copies on a copy-in stream, the multi-layer LPG
launches on a compute stream, device->host copies on a copy-out stream, with `slots` device
staging sets so that the copies of step k+1 overlap the kernels of step k.
"""
import torch

from . import ops

# the three LPG layers of one decoder: (upratio, ds_stride)  -- bts_decoder.py:80-81, 87-88, 94
DECODER_SCALES = ((8, 4), (4, 2), (2, 0))


def layer_shapes(B, H, W, scales=DECODER_SCALES):
    """Shapes of every tensor of an LPG forward+backward pass at image size HxW."""
    out = []
    for r, d in scales:
        out.append(dict(upratio=r, ds_stride=d, coef=(B, H // r, W // r, 3), full=(B, H, W, 1),
                        ds=(B, H // d, W // d, 1) if d else None))
    return out


def algorithmic_bytes(B, H, W, elem_bytes, scales=DECODER_SCALES):
    """SURVEY 8(d) / BASELINE.md section 3:  fwd = s*B*(3hw + HW + HW/d^2), bwd = s*B*(HW + HW/d^2 + 3hw + 3hw)."""
    fwd = bwd = 0
    per = []
    for r, d in scales:
        hw, HW = (H // r) * (W // r), H * W
        ds = HW // (d * d) if d else 0
        f = elem_bytes * B * (3 * hw + HW + ds)
        b = elem_bytes * B * (HW + ds + 3 * hw + 3 * hw)
        per.append((r, f, b))
        fwd += f
        bwd += b
    return fwd, bwd, per


class DeviceSet:
    """One set of device tensors for the LPG layers of a decoder (inputs, outputs, gradients)."""

    def __init__(self, B, H, W, dtype, device, scales=DECODER_SCALES, generator=None, fill=True):
        self.layers = []
        for spec in layer_shapes(B, H, W, scales):
            r, d = spec["upratio"], spec["ds_stride"]
            L = dict(upratio=r, ds_stride=d)
            if fill:
                # SURVEY 8(d) synthetic inputs: coef = sigmoid(N(0,1)), gradients ~ N(0,1)
                L["coef"] = torch.sigmoid(torch.randn(spec["coef"], device=device, generator=generator)).to(dtype)
                L["g_full"] = torch.randn(spec["full"], device=device, generator=generator).to(dtype)
                L["g_ds"] = torch.randn(spec["ds"], device=device, generator=generator).to(dtype) if d else None
            else:
                L["coef"] = torch.empty(spec["coef"], device=device, dtype=dtype)
                L["g_full"] = torch.empty(spec["full"], device=device, dtype=dtype)
                L["g_ds"] = torch.empty(spec["ds"], device=device, dtype=dtype) if d else None
            L["out_full"] = torch.empty(spec["full"], device=device, dtype=dtype)
            L["out_ds"] = torch.empty(spec["ds"], device=device, dtype=dtype) if d else None
            L["g_coef"] = torch.empty(spec["coef"], device=device, dtype=dtype)
            self.layers.append(L)

    def forward(self, fused=True):
        if fused:
            ops.lpg_forward_multi(self.layers)
        else:
            for L in self.layers:
                ops.lpg_forward(L["coef"], L["upratio"], L["ds_stride"], out_full=L["out_full"], out_ds=L["out_ds"])

    def backward(self, fused=True):
        if fused:
            ops.lpg_backward_multi(self.layers)
        else:
            for L in reversed(self.layers):
                ops.lpg_backward(L["coef"], L["g_full"], L["g_ds"], L["upratio"], L["ds_stride"], g_coef=L["g_coef"])

    INPUTS = ("coef", "g_full", "g_ds")
    OUTPUTS = ("out_full", "out_ds", "g_coef")


class HostSet:
    """Pinned host mirrors of a DeviceSet's inputs and outputs."""

    def __init__(self, dev_set, copy_inputs=True):
        self.layers = []
        for L in dev_set.layers:
            H = {}
            for k in DeviceSet.INPUTS + DeviceSet.OUTPUTS:
                t = L[k]
                if t is None:
                    H[k] = None
                    continue
                H[k] = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
                if copy_inputs and k in DeviceSet.INPUTS:
                    H[k].copy_(t)
            self.layers.append(H)

    def bytes_in(self):
        return sum(H[k].numel() * H[k].element_size() for H in self.layers for k in DeviceSet.INPUTS if H[k] is not None)

    def bytes_out(self, return_ds=True):
        return sum(H[k].numel() * H[k].element_size() for H in self.layers for k in DeviceSet.OUTPUTS
                   if H[k] is not None and (return_ds or k != "out_ds"))


class HostLpgPipeline:
    """LPG forward+backward of one decoder for HOST tensors, software-pipelined over `slots` device sets."""

    def __init__(self, B, H, W, dtype, device, slots=2, scales=DECODER_SCALES, fused=True, return_ds=True, run_kernels=True):
        """return_ds=False: the strided copies out_ds (= out_full[:, ::d, ::d], bts_decoder.py:81,88) are not copied back to the
        host -- a host consumer can slice them from out_full; saves 12 MB of the 169 MB device->host traffic per step at config 2.
        run_kernels=False: the copy-only leg of the bench (same buffers, streams and events, no kernel launches) that measures
        the platform's host<->device ceiling for this traffic pattern."""
        self.device = torch.device(device)
        self.fused = fused
        self.return_ds, self.run_kernels = bool(return_ds), bool(run_kernels)
        self.slots = [DeviceSet(B, H, W, dtype, device, scales, fill=False) for _ in range(slots)]
        self.s_in, self.s_compute, self.s_out = (torch.cuda.Stream(device) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(slots)]
        self.ev_compute = [torch.cuda.Event() for _ in range(slots)]
        self.ev_out = [torch.cuda.Event() for _ in range(slots)]
        self.count = 0

    def step(self, host):
        """Enqueue one pass: host inputs -> device, kernels, device -> host outputs.  Asynchronous;
        call drain() (or synchronise ev_out) before reading the host outputs."""
        k = self.count % len(self.slots)
        dev = self.slots[k]
        with torch.cuda.stream(self.s_in):
            if self.count >= len(self.slots):
                self.s_in.wait_event(self.ev_out[k])          # slot's previous results have left the device
            for L, Hh in zip(dev.layers, host.layers):
                for name in DeviceSet.INPUTS:
                    if L[name] is not None:
                        L[name].copy_(Hh[name], non_blocking=True)
            self.ev_in[k].record(self.s_in)
        with torch.cuda.stream(self.s_compute):
            self.s_compute.wait_event(self.ev_in[k])
            if self.run_kernels:
                dev.forward(self.fused)
                dev.backward(self.fused)
            self.ev_compute[k].record(self.s_compute)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_compute[k])
            for L, Hh in zip(dev.layers, host.layers):
                for name in DeviceSet.OUTPUTS:
                    if L[name] is not None and (self.return_ds or name != "out_ds"):
                        Hh[name].copy_(L[name], non_blocking=True)
            self.ev_out[k].record(self.s_out)
        self.count += 1
        return self.ev_out[k]

    def join(self, stream=None):
        """Make `stream` (default: current) wait for everything enqueued so far."""
        stream = stream or torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_compute, self.s_out):
            stream.wait_stream(s)

    def drain(self):
        for s in (self.s_in, self.s_compute, self.s_out):
            s.synchronize()
