"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the reference's LPG hot path.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg and
``bench.py --impl reference``.  The product package (bts-fully-tf_b200/) must never import
this package; it has no CPU fallback and fails loudly when its CUDA library is missing.

Modules
  c_oracle      ctypes binding of lpg_oracle.c (literal fp32 + closed-form fp64, fwd and bwd)
  lpg_literal   op-by-op torch-CPU restatement of custom_layers.py:30-56 (autograd backward);
                also the "port" timed as the CPU baseline, since TensorFlow is absent
  lpg_closed    independent numpy closed form (fp64)
  tail_oracle   numpy restatements of bts.py:27-41, custom_eval_metrics.py:24-88 and the decoder glue rows (concat, up-sampling,
                last convolution, iconv1 over the concat's sources)
  optim_oracle  numpy restatement of custom_optimizers.py:47-59 over Keras Adam, the schedule of bts_train.py:125-131 and the
                uint16 image line of bts_predict.py:140-141
  decoder_fixture  regenerates the kernels of tests/golden/decoder_f256.npz from their seed
  tf_shim/      torch-CPU stand-in for the tf symbols the reference imports; lets
                tests/golden/make_golden.py run the UNMODIFIED reference files

Parity status: pinned to the reference's own source run over tf_shim (tests/golden/*.npz);
unpinned with respect to a real TensorFlow run (TF cannot be installed here).
"""
