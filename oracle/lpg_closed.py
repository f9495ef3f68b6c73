"""TEST INFRASTRUCTURE ONLY -- numpy closed form of the LPG layer, an
independent second statement used to cross-check lpg_oracle.c and lpg_literal.py.

out[b,y,x] = n4[b,i,j] / ((a*n1 + b_*n2 + n3)/sqrt(a^2+b_^2+1) + eps),  i=y//r, j=x//r,
a = ((y%r) - (r-1)/2)/r  pairs with n1 (ROWS), b_ with n2 (COLUMNS)   custom_layers.py:33-43
"""
import numpy as np

PI_F = float(np.float32(np.pi))        # python `pi` meets a float32 tensor -> float32 constant
EPS_F = float(np.float32(1e-7))        # K.epsilon()


def directions(r, dtype=np.float64):
    """(r, r, 3) unit directions [row-offset, col-offset, 1]/norm."""
    k = (np.arange(r, dtype=np.float64) - (r - 1) / 2) / r
    a, b = np.meshgrid(k, k, indexing="ij")
    inv = 1.0 / np.sqrt(a * a + b * b + 1.0)
    return np.stack([a * inv, b * inv, inv], -1).astype(dtype)


def decode(coef):
    coef = np.asarray(coef, np.float64)
    phi = coef[..., 0] * 2.0 * PI_F
    theta = coef[..., 1] * PI_F / 3.0
    return np.sin(phi), np.cos(phi), np.sin(theta), np.cos(theta), coef[..., 2]


def forward(coef, r, return_den=False):
    sp, cp, st, ct, n4 = decode(coef)
    B, h, w = n4.shape
    n = np.stack([st * cp, st * sp, ct], -1)                       # (B,h,w,3)
    d = directions(r)                                              # (r,r,3)
    den = np.einsum("bijc,pqc->bipjq", n, d) + EPS_F               # (B,h,r,w,r)
    out = n4[:, :, None, :, None] / den
    out = out.reshape(B, h * r, w * r)
    return (out, den.reshape(B, h * r, w * r)) if return_den else out


def backward(coef, g_full, r, g_ds=None, d=0):
    sp, cp, st, ct, n4 = decode(coef)
    B, h, w = n4.shape
    G = np.array(g_full, np.float64).reshape(B, h * r, w * r).copy()
    if g_ds is not None:
        G[:, ::d, ::d] += np.asarray(g_ds, np.float64).reshape(B, h * r // d, w * r // d)
    G = G.reshape(B, h, r, w, r)
    n = np.stack([st * cp, st * sp, ct], -1)
    dirs = directions(r)
    den = np.einsum("bijc,pqc->bipjq", n, dirs) + EPS_F
    g4 = (G / den).sum(axis=(2, 4))
    t = -G * n4[:, :, None, :, None] / den ** 2
    gn = np.einsum("bipjq,pqc->bijc", t, dirs)
    g1, g2, g3 = gn[..., 0], gn[..., 1], gn[..., 2]
    gth = g1 * ct * cp + g2 * ct * sp - g3 * st
    gph = -g1 * st * sp + g2 * st * cp
    return np.stack([2.0 * PI_F * gph, (PI_F / 3.0) * gth, g4], -1)
