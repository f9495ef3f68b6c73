"""GPU parity of the tcgen05 implicit-GEMM iconv1 (btslpg_iconv1_forward; bts_decoder.py:98-100 without concat1) against the
float64 oracle, against the library path (ELU + concat kernel + cuDNN convolution), and the reference decoder fixture."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from bts_fully_tf_b200 import ops
from oracle import tail_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# TF32 operands (10-bit mantissa, rounded to nearest: relative error 2^-11 each), float32 accumulation over 9*(NF+3) products:
# the error of a sum is ~ 2^-11 * sqrt(2) * |typical term| * sqrt(K) against a result of ~ |typical term| * sqrt(K): about 7e-4 of
# the output's scale.  Stated tolerance: 3e-3 of the largest output magnitude -- the same class as the library path it replaces
# (cuDNN with TF32 enabled), whose own difference to float64 is reported next to it.
TOL = 3e-3


def _inputs(B, H, W, NF, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, H, W, NF, generator=g)
    planes = [torch.rand(B, H, W, 1, generator=g) * 10.0 for _ in range(3)]
    limit = (6.0 / (9 * (NF + 3) + 9 * NF)) ** 0.5                   # glorot_uniform of a (3,3,NF+3,NF) kernel
    hwio = (torch.rand(3, 3, NF + 3, NF, generator=g) * 2 - 1) * limit
    return a, planes, hwio


@pytest.mark.parametrize("NF", [32, 16])
@pytest.mark.parametrize("B,H,W", [(1, 8, 16), (2, 5, 7), (1, 40, 126), (1, 33, 127), (2, 35, 260), (1, 64, 640)])
@pytest.mark.parametrize("act_out", [False, True])
def test_iconv1_matches_oracle(B, H, W, NF, act_out):
    a, planes, hwio = _inputs(B, H, W, NF, seed=H * 1000 + W)
    out = ops.iconv1_forward(a.to(DEV), [p.to(DEV) for p in planes], hwio.to(DEV), act_out=act_out)
    torch.cuda.synchronize()
    assert ops.last_kernel().startswith("iconv1_fwd_tcgen05"), ops.last_kernel()
    ref = tail_oracle.iconv1_forward(a.numpy(), [p.numpy() for p in planes], hwio.numpy(), act_out=act_out)
    err = np.abs(out.cpu().numpy() - ref).max()
    assert err <= TOL * np.abs(ref).max(), (err, np.abs(ref).max())


def test_iconv1_subpixel_source_and_library_path():
    """a in the sub-pixel layout of the low-res upconv (the decoder's inference path), against the library path on the same GPU:
    ELU + concat kernel + cuDNN convolution in full float32."""
    B, H, W, NF = 2, 48, 160, 32
    g = torch.Generator().manual_seed(3)
    a4 = torch.randn(B, H // 2, W // 2, 4 * NF, generator=g)
    planes = [torch.rand(B, H, W, 1, generator=g) * 10.0 for _ in range(3)]
    hwio = (torch.rand(3, 3, NF + 3, NF, generator=g) * 2 - 1) * 0.1
    out = ops.iconv1_forward(a4.to(DEV), [p.to(DEV) for p in planes], hwio.to(DEV), a_subpixel=True)
    # library path: pixel shuffle -> ELU -> concat -> conv (float32, TF32 off)
    full = a4.view(B, H // 2, W // 2, 2, 2, NF).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, NF)
    ref = tail_oracle.iconv1_forward(full.numpy(), [p.numpy() for p in planes], hwio.numpy())
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        cat = torch.cat([F.elu(full.to(DEV))] + [p.to(DEV) for p in planes], 3).permute(0, 3, 1, 2)
        lib = F.conv2d(cat, hwio.to(DEV).permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    torch.cuda.synchronize()
    scale = np.abs(ref).max()
    assert np.abs(lib.cpu().numpy() - ref).max() <= 1e-5 * scale               # the oracle and cuDNN float32 agree
    assert np.abs(out.cpu().numpy() - ref).max() <= TOL * scale


def test_iconv1_is_bit_reproducible_and_ignores_pad_garbage():
    a, planes, hwio = _inputs(2, 37, 300, 32, seed=9)
    args = (a.to(DEV), [p.to(DEV) for p in planes], hwio.to(DEV))
    out1 = ops.iconv1_forward(*args)
    junk = torch.full((64 << 20,), float("nan"), device=DEV)            # churn memory between the runs
    del junk
    out2 = ops.iconv1_forward(*args)
    torch.cuda.synchronize()
    assert torch.equal(out1, out2) and bool(torch.isfinite(out1).all())


def test_iconv1_rejects_bad_arguments():
    a, planes, hwio = _inputs(1, 8, 16, 32)
    with pytest.raises(ValueError):
        ops.iconv1_forward(a.to(DEV)[..., :24].contiguous(), [p.to(DEV) for p in planes], hwio.to(DEV))      # NF = 24: no fused variant
    with pytest.raises(ValueError):
        ops.iconv1_forward(a, [p.to(DEV) for p in planes], hwio.to(DEV))                                     # host tensor: no CPU fallback
    with pytest.raises(ValueError):
        ops.iconv1_forward(a.to(DEV), [p.to(DEV) for p in planes], hwio.to(DEV)[..., :16].contiguous())      # kernel too small


def test_iconv1_guard_bands_and_repeatability():
    """compute-sanitizer is closed on this GPU pool, so the kernel's memory discipline is checked the hard way: every tensor
    lives inside a larger buffer of NaN (inputs: an out-of-bounds READ poisons the result) or of a sentinel (output: an
    out-of-bounds WRITE is seen), on a ragged shape (three strips, two row segments, odd last strip); and twenty runs of the
    mbarrier / cp.async / tcgen05 pipeline must be bit-identical (a missing fence or a reused ring row shows up as a flicker)."""
    B, H, W, NF = 2, 34, 262, 32
    g = torch.Generator().manual_seed(11)
    pad = 4096

    def guarded(t, fill):
        buf = torch.full((t.numel() + 2 * pad,), fill, device=DEV)
        buf[pad:pad + t.numel()] = t.reshape(-1).to(DEV)
        return buf, buf[pad:pad + t.numel()].view(t.shape)

    a4 = torch.randn(B, H // 2, W // 2, 4 * NF, generator=g)
    planes = [torch.rand(B, H, W, 1, generator=g) * 10 for _ in range(3)]
    hwio = (torch.rand(3, 3, NF + 3, NF, generator=g) * 2 - 1) * 0.1
    _, a_dev = guarded(a4, float("nan"))
    p_dev = [guarded(p, float("nan"))[1] for p in planes]
    _, w_dev = guarded(hwio, float("nan"))
    out_buf, out_dev = guarded(torch.zeros(B, H, W, NF), -12345.0)
    ops.iconv1_forward(a_dev, p_dev, w_dev, a_subpixel=True, out=out_dev)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out_dev).all())
    assert bool((out_buf[:pad] == -12345.0).all()) and bool((out_buf[pad + out_dev.numel():] == -12345.0).all())
    full = a4.view(B, H // 2, W // 2, 2, 2, NF).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, NF)
    ref = tail_oracle.iconv1_forward(full.numpy(), [p.numpy() for p in planes], hwio.numpy())
    assert np.abs(out_dev.cpu().numpy() - ref).max() <= TOL * np.abs(ref).max()
    first = out_dev.clone()
    for _ in range(20):
        out_dev.fill_(0.0)
        ops.iconv1_forward(a_dev, p_dev, w_dev, a_subpixel=True, out=out_dev)
        assert torch.equal(out_dev, first)
