"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/lpg_oracle.c.

Each function cites the reference lines it restates in lpg_oracle.c.  Arrays are numpy,
contiguous NHWC (the trailing size-1 channel of LPG maps is dropped: (B,H,W))."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblpg_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile lpg_oracle.c with oracle/Makefile (gcc only; no GPU, no reference sources)."""
    src = os.path.join(_HERE, "lpg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_version.restype = ctypes.c_int
    return _lib


def _p32(a):
    return a.ctypes.data_as(_f32p) if a is not None else None


def _p64(a):
    return a.ctypes.data_as(_f64p) if a is not None else None


def _split(a):
    """-> (float32 array or None, float64 array or None), contiguous."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a, None
    return None, np.ascontiguousarray(a, dtype=np.float64)


def pixel_dir_f32(H, W, r):
    """custom_layers.py:30-45 -> (H, W, 3) float32."""
    out = np.empty((H, W, 3), np.float32)
    lib().oracle_lpg_pixel_dir_f32(H, W, r, _p32(out))
    return out


def lpg_forward_f32(coef, r):
    """custom_layers.py:47-56 literal, float32.  coef (B,h,w,3) -> (B,H,W)."""
    coef = np.ascontiguousarray(coef, np.float32)
    B, h, w, _ = coef.shape
    out = np.empty((B, h * r, w * r), np.float32)
    lib().oracle_lpg_forward_f32(_p32(coef), B, h, w, r, _p32(out))
    return out


def lpg_forward_f64(coef, r, return_den=False):
    """Closed form, float64 math on float32 or float64 coef -> (B,H,W) [, den]."""
    c32, c64 = _split(coef)
    B, h, w, _ = (c32 if c32 is not None else c64).shape
    out = np.empty((B, h * r, w * r), np.float64)
    den = np.empty_like(out) if return_den else None
    lib().oracle_lpg_forward_f64(_p32(c32), _p64(c64), B, h, w, r, _p64(out), _p64(den))
    return (out, den) if return_den else out


def downsample(full, d):
    """bts_decoder.py:81,88: full[:, ::d, ::d]."""
    return np.ascontiguousarray(full[:, ::d, ::d])


def lpg_backward_f64(coef, g_full, r, g_ds=None, d=0):
    """SURVEY 8(a) a6/a9 closed-form backward -> (B,h,w,3) float64."""
    c32, c64 = _split(coef)
    B, h, w, _ = (c32 if c32 is not None else c64).shape
    gf = np.ascontiguousarray(g_full, np.float64).reshape(B, h * r, w * r) if g_full is not None else None
    gd = None
    if g_ds is not None:
        assert d > 0 and (h * r) % d == 0 and (w * r) % d == 0
        gd = np.ascontiguousarray(g_ds, np.float64).reshape(B, h * r // d, w * r // d)
    out = np.empty((B, h, w, 3), np.float64)
    lib().oracle_lpg_backward_f64(_p32(c32), _p64(c64), _p64(gf), _p64(gd), int(d), B, h, w, r, _p64(out))
    return out


def head_forward_f64(feat, kernel):
    """bts_decoder.py:79,86,93: sigmoid(feat . kernel), kernel [C][3] -> (..., 3) float64."""
    f32, f64 = _split(feat)
    w32, w64 = _split(kernel)
    f = f32 if f32 is not None else f64
    C = f.shape[-1]
    npix = f.size // C
    out = np.empty(f.shape[:-1] + (3,), np.float64)
    lib().oracle_head_forward_f64(_p32(f32), _p64(f64), _p32(w32), _p64(w64),
                                  ctypes.c_size_t(npix), C, _p64(out))
    return out


def head_backward_f64(feat, kernel, coef, g_coef):
    """SURVEY 8(a) a8 -> (g_feat like feat, g_kernel [C][3]) float64."""
    f32, f64 = _split(feat)
    w32, w64 = _split(kernel)
    f = f32 if f32 is not None else f64
    C = f.shape[-1]
    npix = f.size // C
    coef = np.ascontiguousarray(coef, np.float64)
    g_coef = np.ascontiguousarray(g_coef, np.float64)
    g_feat = np.empty(f.shape, np.float64)
    g_w = np.empty((C, 3), np.float64)
    lib().oracle_head_backward_f64(_p32(f32), _p64(f64), _p32(w32), _p64(w64), _p64(coef), _p64(g_coef),
                                   ctypes.c_size_t(npix), C, _p64(g_feat), _p64(g_w))
    return g_feat, g_w


def lpg_fwdbwd_f32(coef, g_full, r):
    """Scalar float32 closed-form fwd+bwd (single-pass port used for a CPU timing figure)."""
    coef = np.ascontiguousarray(coef, np.float32)
    B, h, w, _ = coef.shape
    out = np.empty((B, h * r, w * r), np.float32)
    g = np.ascontiguousarray(g_full, np.float32) if g_full is not None else None
    gc = np.empty_like(coef) if g is not None else None
    lib().oracle_lpg_fwdbwd_f32(_p32(coef), _p32(g), B, h, w, r, _p32(out), _p32(gc))
    return out, gc
