"""CPU tests of the boundary: the shared library loads, exports every symbol include/btslpg.h
declares, validates arguments without touching a GPU, and the Python surface mirrors the
reference layer protocol.  No compute calls here."""
import ctypes
import os
import re

import pytest
import torch

import bts_fully_tf_b200 as pkg
from bts_fully_tf_b200 import _cabi, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _cabi.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "btslpg.h")).read()
    return sorted(set(re.findall(r"BTSLPG_API[^;(]*?\b(btslpg_\w+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "libbtslpg.so does not export %s" % n
    assert sorted(_cabi.SYMBOLS) == names, "ctypes table and header disagree"


def test_version_and_strings(lib):
    assert lib.btslpg_version() == 100
    assert lib.btslpg_status_string(0) == b"ok"
    assert b"CUDA" in lib.btslpg_status_string(-7)


def test_bts_tensor_is_dltensor_layout():
    assert ctypes.sizeof(_cabi.BtsTensor) == 48
    assert _cabi.BtsTensor.data.offset == 0 and _cabi.BtsTensor.device.offset == 8
    assert _cabi.BtsTensor.ndim.offset == 16 and _cabi.BtsTensor.dtype.offset == 20
    assert _cabi.BtsTensor.shape.offset == 24 and _cabi.BtsTensor.strides.offset == 32
    assert _cabi.BtsTensor.byte_offset.offset == 40


def test_dlpack_capsule_is_borrowed_zero_copy():
    t = torch.arange(24, dtype=torch.float32).reshape(1, 2, 4, 3)
    ref = _cabi.from_dlpack_capsule(torch.utils.dlpack.to_dlpack(t))
    assert ref.struct.data == t.data_ptr() and ref.struct.ndim == 4
    assert [ref.struct.shape[k] for k in range(4)] == [1, 2, 4, 3]
    assert (ref.struct.dtype.code, ref.struct.dtype.bits) == (2, 32)
    r2 = _cabi.as_ref(t)
    assert r2.struct.data == t.data_ptr() and [r2.struct.strides[k] for k in range(4)] == [24, 12, 3, 1]


def test_host_tensors_are_rejected_not_computed(lib):
    """No CPU fallback: a host tensor is an error raised by the library itself."""
    coef = torch.rand(1, 2, 2, 3)
    with pytest.raises(ValueError, match="no CPU fallback"):
        ops.lpg_forward(coef, 8)
    with pytest.raises(ValueError, match="not a CUDA tensor"):
        ops.lpg_backward(coef, torch.rand(1, 16, 16, 1), None, 8)
    with pytest.raises(ValueError, match="not a CUDA tensor"):
        ops.reduce_lpg_forward(torch.rand(1, 2, 2, 32), torch.rand(32, 3), 8)


def test_argument_validation_without_gpu(lib):
    def fake_cuda(t):
        ref = _cabi.from_torch(t)
        ref.struct.device.device_type = 2       # pretend; validation fails before any launch
        return ref
    coef = fake_cuda(torch.rand(1, 2, 2, 4))
    out = fake_cuda(torch.rand(1, 16, 16, 1))
    assert lib.btslpg_forward(coef.ptr, 8, out.ptr, None, 0, None) == -3          # last dim must be 3
    assert b"[phi, theta, dist]" in lib.btslpg_last_error()
    coef = fake_cuda(torch.rand(1, 2, 2, 3))
    bad = fake_cuda(torch.rand(1, 16, 15, 1))
    assert lib.btslpg_forward(coef.ptr, 8, bad.ptr, None, 0, None) == -3
    assert lib.btslpg_forward(coef.ptr, 0, out.ptr, None, 0, None) == -1          # upratio
    ds = fake_cuda(torch.rand(1, 4, 4, 1))
    assert lib.btslpg_forward(coef.ptr, 8, out.ptr, ds.ptr, 3, None) == -1        # ds stride must divide r
    half = fake_cuda(torch.rand(1, 2, 2, 3).half())
    assert lib.btslpg_forward(half.ptr, 8, out.ptr, None, 0, None) == -2          # fp16 unsupported
    out_bf = fake_cuda(torch.rand(1, 16, 16, 1).bfloat16())
    assert lib.btslpg_forward(coef.ptr, 8, out_bf.ptr, None, 0, None) == -2       # mixed dtypes
    assert lib.btslpg_forward(None, 8, out.ptr, None, 0, None) == -1
    assert lib.btslpg_backward(coef.ptr, out.ptr, None, 8, 0, None, None) == -1   # g_coef NULL
    assert lib.btslpg_forward_multi(None, 0, None) == -1
    assert lib.btslpg_reduce_backward_workspace_bytes(2457600, 64) >= 256 + 64 * 3 * 4


def test_argument_validation_of_the_decoder_glue_entry_points(lib):
    """tail / concat / up-sampling / slice / depth-conv entry points: validation errors without any launch."""
    def fake_cuda(t):
        ref = _cabi.from_torch(t)
        ref.struct.device.device_type = 2
        return ref
    x = fake_cuda(torch.rand(1, 4, 4, 8))
    out = fake_cuda(torch.rand(1, 4, 4, 9))
    plane = fake_cuda(torch.rand(1, 4, 4, 1))
    arr = (_cabi._TP * 1)(plane.ptr)
    assert lib.btslpg_concat_forward(x.ptr, 0, 2, None, None, None, arr, 1, 0, out.ptr, None) == -1            # act out of range
    assert lib.btslpg_concat_forward(x.ptr, 0, 0, None, None, None, arr, 1, 9, out.ptr, None) == -1            # pad out of range
    assert lib.btslpg_concat_forward(x.ptr, 0, 0, x.ptr, None, None, arr, 1, 0, out.ptr, None) == -1           # scale without shift
    wide = fake_cuda(torch.rand(1, 4, 4, 12))
    assert lib.btslpg_concat_forward(x.ptr, 0, 0, None, None, None, arr, 1, 0, wide.ptr, None) == -3           # out must be CA + planes + pad wide
    big = fake_cuda(torch.rand(1, 8, 9, 8))
    assert lib.btslpg_upsample2x_forward(x.ptr, big.ptr, None) == -3                                         # (B, 2h, 2w, C)
    assert lib.btslpg_affine_act(x.ptr, None, None, 3, x.ptr, None) == -1                                    # act out of range
    assert lib.btslpg_affine_act(x.ptr, None, None, 0, out.ptr, None) == -3                                  # shape differs
    g = fake_cuda(torch.rand(1, 4, 4, 1))
    k = fake_cuda(torch.rand(72))
    assert lib.btslpg_depthconv_backward(x.ptr, k.ptr, g.ptr, 0, x.ptr, None, None, 0, None) == -3           # C must be 16 or 32
    assert b"C = 16 and C = 32" in lib.btslpg_last_error()
    y1 = fake_cuda(torch.rand(1, 4, 4, 1))
    assert lib.btslpg_depthconv_forward(x.ptr, k.ptr, 1, 1, 10.0, y1.ptr, None) == -3                        # C must be 16 or 32
    x32 = fake_cuda(torch.rand(1, 4, 4, 32))
    k288 = fake_cuda(torch.rand(288))
    assert lib.btslpg_depthconv_backward(x32.ptr, k288.ptr, y1.ptr, 2, x32.ptr, None, None, 0, None) == -1   # act_in out of range
    assert lib.btslpg_depthconv_forward(x32.ptr, k288.ptr, 2, 0, 1.0, y1.ptr, None) == -1                    # act_in out of range
    assert lib.btslpg_depthconv_forward(x32.ptr, k288.ptr, 1, 3, 1.0, y1.ptr, None) == -1                    # act_out out of range
    assert lib.btslpg_depthconv_forward(x32.ptr, k288.ptr, 1, 1, 1.0, None, None) == -1                      # y missing
    assert lib.btslpg_depthconv_forward(x32.ptr, k.ptr, 1, 1, 1.0, y1.ptr, None) == -3                       # kernel needs 9*C floats
    host = _cabi.from_torch(torch.rand(1, 4, 4, 32))
    assert lib.btslpg_depthconv_forward(host.ptr, k288.ptr, 1, 1, 1.0, y1.ptr, None) < 0                     # host tensor: no CPU fallback
    yt = fake_cuda(torch.rand(1, 4, 4, 1))
    assert lib.btslpg_silog_forward(None, None, 10.0, 0.1, yt.ptr, None, None, 0, None) == -1                # nothing to compute
    assert lib.btslpg_silog_forward(yt.ptr, yt.ptr, 10.0, 0.1, yt.ptr, None, None, 0, None) == -1            # loss tensor missing
    m = fake_cuda(torch.rand(4))
    ws = ctypes.c_void_p(256)
    assert lib.btslpg_eval_metrics(yt.ptr, yt.ptr, 1e-3, 10.0, m.ptr, ws, 1 << 20, None) == -3               # metrics needs 10 floats
    assert lib.btslpg_tail_workspace_bytes() >= 256 + 148 * 10 * 8
    assert lib.btslpg_depthconv_backward_workspace_bytes(32) >= 256 + 9 * 32 * 4


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libbtslpg.so")
    with pytest.raises(pkg.BtsLpgLibraryMissing, match="no CPU or pure-PyTorch fallback"):
        _cabi.load()


def test_layer_protocol_mirrors_reference():
    """custom_layers.py:25-61: ctor kwargs, build assert, get_config round trip, layer names."""
    layer = pkg.LocalPlanarGuidance(upratio=8, name="depth_8x8_scaled")
    assert layer.name == "depth_8x8_scaled" and layer.upratio == 8
    cfg = layer.get_config()
    assert cfg["upratio"] == 8 and cfg["name"] == "depth_8x8_scaled" and {"trainable", "dtype"} <= set(cfg)
    again = pkg.LocalPlanarGuidance.from_config(cfg)
    assert again.get_config() == cfg
    with pytest.raises(AssertionError):
        layer.build((4, 3))                                   # assert len(input_shape) > 2
    layer.build((2, 60, 80, 3))
    assert layer.built and layer.compute_output_shape((2, 60, 80, 3)) == (2, 480, 640, 1)
    with pytest.raises(TypeError):
        pkg.LocalPlanarGuidance(upratio=2, bogus=1)
    assert list(layer.parameters()) == []                     # the layer has no weights
    head = pkg.ReductionLPG(128, 8, ds_stride=4)
    assert tuple(head.kernel.shape) == (1, 1, 128, 3) and head.name == "reduction_8x8"


def test_golden_fixture_config(golden_dir):
    import numpy as np
    z = np.load(os.path.join(golden_dir, "lpg_r8.npz"))
    layer = pkg.LocalPlanarGuidance(upratio=int(z["config_upratio"]), name=str(z["config_name"]))
    assert layer.get_config()["upratio"] == 8 and layer.get_config()["name"] == "depth_8x8_scaled"
