"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the reference's LPG hot path.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg and
``bench.py --impl reference``.  The product package (bts-fully-tf_b200/) must never import
this package; it has no CPU fallback and fails loudly when its CUDA library is missing.

Modules
  c_oracle      ctypes binding of lpg_oracle.c (literal fp32 + closed-form fp64, fwd and bwd)
  lpg_literal   op-by-op torch-CPU restatement of custom_layers.py:30-56 (autograd backward);
                also the "port" timed as the CPU baseline, since TensorFlow is absent
  decoder_ref   torch restatement of bts_decoder.py:26-105 (whole-decoder parity, next rows)
  tf_shim/      torch-CPU stand-in for the tf symbols the reference imports; lets
                tests/golden/make_golden.py run the UNMODIFIED reference files

Parity status: pinned to the reference's own source run over tf_shim (tests/golden/*.npz);
unpinned with respect to a real TensorFlow run (TF cannot be installed here).
"""
