"""Parity rules shared by the tests (SURVEY section 7 "hard parts" + 8(c)).

The oracle (oracle/, CPU) is the checker; the thing checked is always the CUDA path reached
through the C ABI.  Tolerances:

  float32 forward   |gpu - ref| <= 1e-5 * |ref|            where den_ref >= DEN_OK (0.05)
                    |n4/gpu - den_ref| <= 1e-6              elsewhere (the layer divides by a
                    denominator that can cross zero at r = 8; the reference does not clamp and
                    float32 cannot hold 1e-5 relative there), sign / inf / NaN must agree
  float32 backward  |gpu - ref| <= 1e-5 * scale             scale = sum over the patch of |term|
                    (the sum cancels; tolerance is relative to what was summed, not to the result)
  bfloat16          same rules with 1e-2 (inputs/outputs are rounded to 8 bits of mantissa)
"""
import numpy as np

from oracle import c_oracle, lpg_closed

DEN_OK = 0.05


def check_forward(out_gpu, coef, r, rtol=1e-5, den_atol=1e-6, what="forward"):
    """out_gpu (B,H,W[,1]) float array; coef float32/float64 (B,h,w,3) as seen by the kernel."""
    ref, den = c_oracle.lpg_forward_f64(np.asarray(coef), r, return_den=True)
    got = np.asarray(out_gpu, np.float64).reshape(ref.shape)
    good = den >= DEN_OK
    err = np.abs(got - ref)
    bad = good & ~(err <= rtol * np.abs(ref))
    assert not bad.any(), "%s: %d of %d well-conditioned pixels off by more than %g rel (max rel %.3g)" % (
        what, bad.sum(), good.sum(), rtol, (err[good] / np.maximum(np.abs(ref[good]), 1e-30)).max())
    if (~good).any():
        n4 = np.repeat(np.repeat(np.asarray(coef, np.float64)[..., 2], r, 1), r, 2)
        idx = ~good
        fin = idx & np.isfinite(ref) & (n4 != 0) & (np.abs(den) > 1e-4)
        with np.errstate(divide="ignore", invalid="ignore"):
            den_gpu = n4 / got
        # in the denominator domain the error budget is a few float32 ulps of O(1) quantities
        tol = den_atol * (rtol / 1e-5)
        badd = fin & ~(np.abs(den_gpu - den) <= tol)
        assert not badd.any(), "%s: %d ill-conditioned pixels disagree in the denominator domain (max %.3g)" % (
            what, badd.sum(), np.abs(den_gpu - den)[fin].max())
        assert np.array_equal(np.sign(got[fin]), np.sign(ref[fin])), "%s: sign mismatch near the pole" % what
    return float((err[good] / np.maximum(np.abs(ref[good]), 1e-30)).max()) if good.any() else 0.0


def backward_scale(coef, g_full, r, g_ds=None, d=0):
    """Per coarse pixel, per channel: the sum of |terms| the backward reduction adds up, and the
    smallest denominator of the patch."""
    coef = np.asarray(coef, np.float64)
    sp, cp, st, ct, n4 = lpg_closed.decode(coef)
    B, h, w = n4.shape
    G = np.zeros((B, h * r, w * r))
    if g_full is not None:
        G += np.asarray(g_full, np.float64).reshape(B, h * r, w * r)
    if g_ds is not None:
        G[:, ::d, ::d] += np.asarray(g_ds, np.float64).reshape(B, h * r // d, w * r // d)
    G = np.abs(G).reshape(B, h, r, w, r)
    n = np.stack([st * cp, st * sp, ct], -1)
    dirs = lpg_closed.directions(r)
    den = np.einsum("bijc,pqc->bipjq", n, dirs) + lpg_closed.EPS_F
    s4 = (G / np.abs(den)).sum(axis=(2, 4))
    t = G * np.abs(n4)[:, :, None, :, None] / den ** 2
    s123 = t.sum(axis=(2, 4))            # |dir components| <= 1
    scale = np.stack([2 * np.pi * 2 * s123, (np.pi / 3) * 3 * s123, s4], -1)
    return scale, den.min(axis=(2, 4))


def check_backward(g_gpu, coef, g_full, r, g_ds=None, d=0, rtol=1e-5, what="backward"):
    ref = c_oracle.lpg_backward_f64(np.asarray(coef), None if g_full is None else np.asarray(g_full, np.float64), r,
                                    None if g_ds is None else np.asarray(g_ds, np.float64), d)
    got = np.asarray(g_gpu, np.float64).reshape(ref.shape)
    scale, den_min = backward_scale(coef, g_full, r, g_ds, d)
    good = (den_min >= DEN_OK)[..., None] & np.ones_like(ref, bool)
    err = np.abs(got - ref)
    bad = good & ~(err <= rtol * scale + 1e-30)
    assert not bad.any(), "%s: %d of %d gradient entries off by more than %g of their summed magnitude (max %.3g)" % (
        what, bad.sum(), good.sum(), rtol, (err[good] / np.maximum(scale[good], 1e-30)).max())
    ill = ~good
    if ill.any():
        # conditioning grows like (DEN_OK/den)^2; only require finite agreement to that degree
        fac = (DEN_OK / np.maximum(np.abs(den_min), 1e-4)) ** 2
        bad2 = ill & np.isfinite(ref) & ~(err <= rtol * 10 * scale * fac[..., None] + 1e-30)
        assert not bad2.any(), "%s: %d ill-conditioned gradient entries disagree" % (what, bad2.sum())
    return float((err[good] / np.maximum(scale[good], 1e-30)).max()) if good.any() else 0.0


def bf16_round(a):
    """Round a float array to bfloat16 precision (round to nearest even), returned as float32."""
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).float().numpy()
