#!/usr/bin/env python
"""Derive and validate the float32 sin/cos kernels used by the fast decode path
(csrc/lpg_kernels.cuh: sincos_quadrant).

  sin(r) ~ r + r^3 (S1 + S2 r^2 + S3 r^4)           |r| <= pi/4
  cos(r) ~ 1 + r^2 (C1 + C2 r^2 + C3 r^4 + C4 r^6)

Coefficients: weighted least squares on Chebyshev nodes in float64 (near-minimax), then rounded
to float32.  Validation emulates the kernel's float32 FMA sequence (products of two floats are
exact in float64, so fma(a,b,c) = float32(float64(a)*b + c) up to a negligible double rounding)
including the magic-number quadrant reduction with a 2-term Cody-Waite pi/2, and reports the max
error against float64 sin/cos of the SAME float32 angle.
"""
import numpy as np

f32 = np.float32


def fit():
    n = 4001
    x = np.cos(np.pi * (np.arange(n) + 0.5) / n) * (np.pi / 4)
    x = x[np.abs(x) > 1e-6]
    s = x * x
    # sin: (sin(x) - x)/x^3 = S1 + S2 s + S3 s^2
    ys = (np.sin(x) - x) / x ** 3
    A = np.stack([np.ones_like(s), s, s * s], 1)
    S = np.linalg.lstsq(A, ys, rcond=None)[0]
    yc = (np.cos(x) - 1) / s
    Ac = np.stack([np.ones_like(s), s, s * s, s ** 3], 1)
    C = np.linalg.lstsq(Ac, yc, rcond=None)[0]
    return S.astype(f32), C.astype(f32)


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


TWO_OVER_PI = f32(0.636619772367581343)
PIO2_HI = f32(1.5707963267948966)                       # fl32(pi/2)
PIO2_LO = f32(1.5707963267948966 - float(PIO2_HI))      # fl32(pi/2 - hi)
MAGIC = f32(12582912.0)                                 # 1.5 * 2^23


def sincos_emulated(a, S, C):
    a = a.astype(f32)
    kf = fma(a, np.full_like(a, TWO_OVER_PI), np.full_like(a, MAGIC))
    q = (kf - MAGIC).astype(f32)
    n = kf.view(np.int32)
    r = fma(q, np.full_like(a, -PIO2_HI), a)
    r = fma(q, np.full_like(a, -PIO2_LO), r)
    s = (r * r).astype(f32)
    ps = fma(np.full_like(a, S[2]), s, np.full_like(a, S[1]))
    ps = fma(ps, s, np.full_like(a, S[0]))
    t = (r * s).astype(f32)
    sn = fma(ps, t, r)
    pc = fma(np.full_like(a, C[3]), s, np.full_like(a, C[2]))
    pc = fma(pc, s, np.full_like(a, C[1]))
    pc = fma(pc, s, np.full_like(a, C[0]))
    cs = fma(pc, s, np.ones_like(a))
    swap = (n & 1).astype(bool)
    so = np.where(swap, cs, sn)
    co = np.where(swap, sn, cs)
    so = np.where((n & 2) != 0, -so, so)
    co = np.where(((n + 1) & 2) != 0, -co, co)
    return so.astype(f32), co.astype(f32)


def div3_emulated(t):
    third = f32(1.0 / 3.0)
    q = (t * third).astype(f32)
    r = fma(np.full_like(t, f32(-3.0)), q, t)
    return fma(r, np.full_like(t, third), q)


def main():
    S, C = fit()
    print("S =", [float(v).hex() for v in S], [float(v) for v in S])
    print("C =", [float(v).hex() for v in C], [float(v) for v in C])
    print("PIO2_HI", float(PIO2_HI).hex(), "PIO2_LO", float(PIO2_LO).hex(), "2/pi", float(TWO_OVER_PI).hex())
    rng = np.random.default_rng(0)
    for name, lo, hi in (("phi   [0, 2pi]", 0.0, 6.2832), ("theta [0, pi/3]", 0.0, 1.0472), ("wide  [-1000, 1000]", -1000.0, 1000.0)):
        a = rng.uniform(lo, hi, 4_000_000).astype(f32)
        so, co = sincos_emulated(a, S, C)
        es = np.abs(so.astype(np.float64) - np.sin(a.astype(np.float64))).max()
        ec = np.abs(co.astype(np.float64) - np.cos(a.astype(np.float64))).max()
        print("%-22s max |sin err| %.3e  max |cos err| %.3e   (ulp(1) = 1.19e-7)" % (name, es, ec))
    # x*pi/3 with a correctly rounded division by 3 in three instructions
    PI_F = f32(np.pi)
    x = np.concatenate([rng.uniform(0, 1, 4_000_000), rng.uniform(-50, 50, 1_000_000), rng.standard_normal(1_000_000) * 1e-3]).astype(f32)
    t = (x * PI_F).astype(f32)
    ref = (t / f32(3.0)).astype(f32)
    got = div3_emulated(t)
    print("div-by-3 mismatches vs IEEE float32 division:", int((ref != got).sum()), "of", x.size)


if __name__ == "__main__":
    main()
