"""TEST INFRASTRUCTURE (not part of the product path).

Seeded glorot_uniform kernels of the reference decoder for the F = 256 fixture (tests/golden/decoder_f256.npz).  The
fixture was recorded by running the UNMODIFIED reference bts_decoder.decoder_model (tests/golden/make_golden.py) on the
tf stand-in of oracle/tf_shim, whose Conv2D draws its kernel in creation order from one seeded torch CPU generator
(Keras default kernel_initializer='glorot_uniform').  Two million weights are too many to commit, so the fixture keeps
their float64 sums and the tests regenerate them here with the same draws; make_golden.py asserts that this function
reproduces the kernels of the recorded run bit for bit.
"""
import math

import torch


def regen_kernels(shapes_hwio, seed):
    """Kernels (float64, HWIO) in creation order, exactly as oracle/tf_shim's Conv2D.build draws them."""
    gen = torch.Generator()
    gen.manual_seed(int(seed))
    out = []
    for shape in shapes_hwio:
        kh, kw, cin, cout = (int(v) for v in shape)
        limit = math.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
        w = torch.rand((kh, kw, cin, cout), generator=gen, dtype=torch.float64)
        out.append((w * 2 - 1) * limit)
    return out


def sample_index(n, k=512):
    """Indices of the k evenly strided entries a large gradient tensor (n elements, flattened) is sampled at."""
    if n <= k:
        return torch.arange(n)
    return torch.linspace(0, n - 1, k).round().long()
