"""bts-fully-tf_b200 -- B200 (sm_100a) implementation of the BTS decoder's Local-Planar-Guidance
hot path behind the reference's Keras-layer surface.

Only what the path needs lives here:
  csrc/        hand-written CUDA kernels + the C ABI (include/btslpg.h) -> lib/libbtslpg.so:
               lpg_* (LPG fwd/bwd), head_* (fused reduction heads), tail_* (sigmoid*max_depth + si_log_loss, eval metrics),
               concat_* (ELU/BatchNorm + concat), upsample_* (nearest x2), slice_* (DenseASPP glue), depthconv_* (last conv backward)
  _cabi.py     ctypes binding (BtsTensor == DLTensor, zero copy)
  ops.py       functional ops + autograd glue
  layers.py    LocalPlanarGuidance / ReductionLPG with the reference's layer protocol
  losses.py    si_log_loss_wrapper (bts.py:27-41) / fused sigmoid*max_depth + loss
  eval_metrics.py  metrics_list_factory (custom_eval_metrics.py) from one fused pass
  decoder.py   decoder_model(...) wiring of bts_decoder.py around the fused ops (torch/cuDNN glue)
  parallel.py  batch sharding + gradient bucket all-reduce (one process per GPU, NCCL)
  host_io.py   pinned-host staging for callers whose tensors live in host memory
  tf_adapter.py  the same C ABI bound into TensorFlow (import-guarded; TF is absent here)

The directory name carries a hyphen (it mirrors the reference repository's name); import it as
`bts_fully_tf_b200`.
"""
from ._cabi import BtsLpgLibraryMissing, LIB_PATH, load as load_library  # noqa: F401
from .layers import LocalPlanarGuidance, ReductionLPG  # noqa: F401
from . import ops  # noqa: F401

__version__ = "0.1.0"
