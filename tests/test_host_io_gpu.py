"""GPU test of the end-to-end product path of the bench (host_io.HostLpgPipeline): host tensors in, host tensors out, device
slots reused round-robin -- results must equal the device-resident ops bit for bit on every step."""
import pytest
import torch

from bts_fully_tf_b200 import ops
from bts_fully_tf_b200.host_io import DeviceSet, HostLpgPipeline, HostSet

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("slots,fused", [(2, True), (3, True), (2, False)])
def test_pipeline_equals_device_resident_ops_across_slot_reuse(dtype, slots, fused):
    B, H, W = 2, 64, 96
    pipe = HostLpgPipeline(B, H, W, dtype, DEV, slots=slots, fused=fused)
    gen = torch.Generator(device=DEV)
    for step in range(2 * slots + 1):                                   # every slot is reused at least once, with NEW data each step
        gen.manual_seed(100 + step)
        ref = DeviceSet(B, H, W, dtype, DEV, generator=gen)             # device-resident inputs
        host = HostSet(ref)                                             # pinned host copies of the inputs, empty host outputs
        pipe.step(host)
        pipe.drain()
        ref.forward(fused)
        ref.backward(fused)
        torch.cuda.synchronize()
        for L, Hh in zip(ref.layers, host.layers):
            for name in DeviceSet.OUTPUTS:
                if L[name] is not None:
                    assert torch.equal(Hh[name], L[name].cpu()), (step, name, L["upratio"])
            if L["ds_stride"]:
                assert torch.equal(Hh["out_ds"], Hh["out_full"][:, ::L["ds_stride"], ::L["ds_stride"]])


def test_pipeline_without_ds_readback_and_copy_only_leg():
    B, H, W = 1, 32, 64
    dtype = torch.float32
    gen = torch.Generator(device=DEV).manual_seed(5)
    ref = DeviceSet(B, H, W, dtype, DEV, generator=gen)
    host = HostSet(ref)
    for L in host.layers:
        if L["out_ds"] is not None:
            L["out_ds"].fill_(-7.0)
    pipe = HostLpgPipeline(B, H, W, dtype, DEV, slots=2, return_ds=False)
    pipe.step(host)
    pipe.drain()
    ref.forward(True)
    torch.cuda.synchronize()
    for L, Hh in zip(ref.layers, host.layers):
        assert torch.equal(Hh["out_full"], L["out_full"].cpu())
        if Hh["out_ds"] is not None:
            assert float(Hh["out_ds"].min()) == -7.0                     # not copied back: the host slices it from out_full
    assert host.bytes_out(False) < host.bytes_out(True)
    # the copy-only leg of the bench launches no kernels
    ops.reset_launch_count()
    idle = HostLpgPipeline(B, H, W, dtype, DEV, slots=2, run_kernels=False)
    idle.step(host)
    idle.drain()
    assert ops.launch_count() == 0
