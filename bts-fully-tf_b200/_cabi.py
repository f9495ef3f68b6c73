"""ctypes binding of libbtslpg.so (include/btslpg.h) -- the only way Python reaches the kernels.

Tensors cross the boundary as ``BtsTensor`` structs, which are layout-compatible with DLPack's
``DLTensor``: a torch tensor is described in place (pointer, shape, strides -- zero copy), and
any other producer (TensorFlow via ``tf.experimental.dlpack.to_dlpack``, CuPy, ...) is accepted
as a DLPack capsule or an object with ``__dlpack__``.

There is NO CPU fallback: if the shared library is missing, importing the ops raises, and host
tensors are rejected by the library itself (BTSLPG_EDEVICE).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BTSLPG_LIB") or os.path.join(_HERE, "lib", "libbtslpg.so")   # BTSLPG_LIB: experiment builds

BTSLPG_MAX_MULTI = 4


class BtsDevice(ctypes.Structure):
    _fields_ = [("device_type", ctypes.c_int32), ("device_id", ctypes.c_int32)]


class BtsDataType(ctypes.Structure):
    _fields_ = [("code", ctypes.c_uint8), ("bits", ctypes.c_uint8), ("lanes", ctypes.c_uint16)]


class BtsTensor(ctypes.Structure):
    """== DLTensor (dlpack.h)."""
    _fields_ = [("data", ctypes.c_void_p), ("device", BtsDevice), ("ndim", ctypes.c_int32), ("dtype", BtsDataType),
                ("shape", ctypes.POINTER(ctypes.c_int64)), ("strides", ctypes.POINTER(ctypes.c_int64)),
                ("byte_offset", ctypes.c_uint64)]


class _DLManagedTensor(ctypes.Structure):
    _fields_ = [("dl_tensor", BtsTensor), ("manager_ctx", ctypes.c_void_p), ("deleter", ctypes.c_void_p)]


class BtsLpgForwardArgs(ctypes.Structure):
    _fields_ = [("coef", ctypes.POINTER(BtsTensor)), ("upratio", ctypes.c_int32), ("ds_stride", ctypes.c_int32),
                ("out_full", ctypes.POINTER(BtsTensor)), ("out_ds", ctypes.POINTER(BtsTensor))]


class BtsLpgBackwardArgs(ctypes.Structure):
    _fields_ = [("coef", ctypes.POINTER(BtsTensor)), ("g_full", ctypes.POINTER(BtsTensor)), ("g_ds", ctypes.POINTER(BtsTensor)),
                ("upratio", ctypes.c_int32), ("ds_stride", ctypes.c_int32), ("g_coef", ctypes.POINTER(BtsTensor))]


class BtsAdamConfig(ctypes.Structure):
    _fields_ = [("lr_start", ctypes.c_float), ("lr_end", ctypes.c_float), ("total_steps", ctypes.c_int64), ("power", ctypes.c_float),
                ("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("epsilon", ctypes.c_float), ("l1", ctypes.c_float),
                ("l2", ctypes.c_float), ("grad_scale", ctypes.c_float), ("zero_grad", ctypes.c_int32)]


_TP = ctypes.POINTER(BtsTensor)

# name -> (restype, argtypes); every symbol include/btslpg.h declares
SYMBOLS = {
    "btslpg_version": (ctypes.c_int, []),
    "btslpg_last_error": (ctypes.c_char_p, []),
    "btslpg_status_string": (ctypes.c_char_p, [ctypes.c_int]),
    "btslpg_forward": (ctypes.c_int, [_TP, ctypes.c_int, _TP, _TP, ctypes.c_int, ctypes.c_void_p]),
    "btslpg_backward": (ctypes.c_int, [_TP, _TP, _TP, ctypes.c_int, ctypes.c_int, _TP, ctypes.c_void_p]),
    "btslpg_forward_multi": (ctypes.c_int, [ctypes.POINTER(BtsLpgForwardArgs), ctypes.c_int, ctypes.c_void_p]),
    "btslpg_backward_multi": (ctypes.c_int, [ctypes.POINTER(BtsLpgBackwardArgs), ctypes.c_int, ctypes.c_void_p]),
    "btslpg_reduce_forward": (ctypes.c_int, [_TP, _TP, ctypes.c_int, _TP, _TP, _TP, ctypes.c_int, ctypes.c_void_p]),
    "btslpg_reduce_backward_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int]),
    "btslpg_reduce_backward": (ctypes.c_int, [_TP, _TP, _TP, _TP, _TP, ctypes.c_int, ctypes.c_int, _TP, _TP, _TP,
                                              ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "btslpg_tail_workspace_bytes": (ctypes.c_size_t, []),
    "btslpg_silog_forward": (ctypes.c_int, [_TP, _TP, ctypes.c_float, ctypes.c_float, _TP, _TP, ctypes.c_void_p, ctypes.c_size_t,
                                            ctypes.c_void_p]),
    "btslpg_silog_backward": (ctypes.c_int, [_TP, _TP, ctypes.c_float, ctypes.c_float, _TP, ctypes.c_void_p, ctypes.c_size_t,
                                             ctypes.c_int, _TP, ctypes.c_void_p]),
    "btslpg_eval_metrics": (ctypes.c_int, [_TP, _TP, ctypes.c_float, ctypes.c_float, _TP, ctypes.c_void_p, ctypes.c_size_t,
                                           ctypes.c_void_p]),
    "btslpg_eval_metrics_png16": (ctypes.c_int, [_TP, _TP, ctypes.c_float, ctypes.c_float, _TP, ctypes.c_float, _TP, ctypes.c_void_p,
                                                 ctypes.c_size_t, ctypes.c_void_p]),
    "btslpg_adam_step": (ctypes.c_int, [_TP, _TP, _TP, _TP, _TP, ctypes.POINTER(BtsAdamConfig), ctypes.c_int, ctypes.c_void_p]),
    "btslpg_device_synchronize": (ctypes.c_int, [ctypes.c_int]),
    "btslpg_concat_forward": (ctypes.c_int, [_TP, ctypes.c_int, ctypes.c_int, _TP, _TP, _TP, ctypes.POINTER(_TP), ctypes.c_int, ctypes.c_int,
                                             _TP, ctypes.c_void_p]),
    "btslpg_concat_backward": (ctypes.c_int, [_TP, _TP, ctypes.c_int, _TP, ctypes.c_int, _TP, ctypes.POINTER(_TP), ctypes.c_int,
                                              ctypes.c_int, ctypes.c_void_p]),
    "btslpg_iconv1_forward": (ctypes.c_int, [_TP, ctypes.c_int, ctypes.POINTER(_TP), _TP, ctypes.c_int, _TP, ctypes.c_void_p]),
    "btslpg_bn_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "btslpg_bn_elu_stats": (ctypes.c_int, [_TP, ctypes.c_int, _TP, _TP, _TP, _TP, ctypes.c_float, ctypes.c_float, _TP, ctypes.c_void_p,
                                           ctypes.c_size_t, ctypes.c_void_p]),
    "btslpg_bn_elu_backward_stats": (ctypes.c_int, [_TP, _TP, ctypes.c_int, _TP, _TP, _TP, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "btslpg_concat_backward_bn": (ctypes.c_int, [_TP, _TP, ctypes.c_int, _TP, _TP, ctypes.c_int, _TP, ctypes.POINTER(_TP), ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_void_p]),
    "btslpg_bn_moments": (ctypes.c_int, [_TP, _TP, _TP, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "btslpg_bn_fold": (ctypes.c_int, [_TP, _TP, _TP, _TP, _TP, _TP, ctypes.c_float, ctypes.c_float, ctypes.c_int64, _TP, _TP, _TP, ctypes.c_void_p]),
    "btslpg_bn_act_backward_stats": (ctypes.c_int, [_TP, _TP, _TP, _TP, _TP, _TP, _TP, ctypes.c_int, _TP, _TP, ctypes.c_void_p, ctypes.c_size_t,
                                                    ctypes.c_void_p]),
    "btslpg_bn_act_backward": (ctypes.c_int, [_TP, _TP, _TP, _TP, _TP, _TP, _TP, _TP, _TP, ctypes.c_int, _TP, ctypes.c_int, _TP, ctypes.c_void_p]),
    "btslpg_conv3x3_wgrad_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "btslpg_conv3x3_wgrad": (ctypes.c_int, [_TP, _TP, _TP, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "btslpg_upsample2x_forward": (ctypes.c_int, [_TP, _TP, ctypes.c_void_p]),
    "btslpg_upsample2x_backward": (ctypes.c_int, [_TP, _TP, ctypes.c_void_p]),
    "btslpg_affine_act": (ctypes.c_int, [_TP, _TP, _TP, ctypes.c_int, _TP, ctypes.c_void_p]),
    "btslpg_affine_act_split": (ctypes.c_int, [_TP, _TP, _TP, ctypes.c_int, _TP, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "btslpg_depthconv_backward_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "btslpg_depthconv_backward": (ctypes.c_int, [_TP, _TP, _TP, ctypes.c_int, _TP, _TP, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "btslpg_depthconv_forward": (ctypes.c_int, [_TP, _TP, ctypes.c_int, ctypes.c_int, ctypes.c_float, _TP, ctypes.c_void_p]),
    "btslpg_launch_count": (ctypes.c_uint64, []),
    "btslpg_reset_launch_count": (None, []),
    "btslpg_last_kernel": (ctypes.c_char_p, []),
    "btslpg_set_block_threads": (None, [ctypes.c_int, ctypes.c_int]),
    "btslpg_set_tuning": (None, [ctypes.c_int, ctypes.c_int]),
}

_lib = None
_lock = threading.Lock()


class BtsLpgLibraryMissing(ImportError):
    pass


def load():
    """dlopen libbtslpg.so (built by bts-fully-tf_b200/build.py or __graft_entry__.build)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise BtsLpgLibraryMissing(
                        "%s not found. Build it with `python bts-fully-tf_b200/build.py` (needs nvcc). "
                        "There is no CPU or pure-PyTorch fallback for the LPG hot path." % LIB_PATH)
                lib = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SYMBOLS.items():
                    fn = getattr(lib, name)      # AttributeError if the ABI is incomplete
                    fn.restype = res
                    fn.argtypes = args
                if os.environ.get("BTSLPG_NVTX"):
                    lib = _NvtxProxy(lib)
                _lib = lib
    return _lib


class _NvtxProxy:
    """BTSLPG_NVTX=1: every launch through the ABI sits in an NVTX range named after its entry point (nsys / ncu --nvtx
    timelines; the reference's counterpart is the TensorBoard profiler of custom_callbacks.py).  Off by default: no overhead."""
    _QUIET = ("btslpg_version", "btslpg_last_error", "btslpg_status_string", "btslpg_launch_count", "btslpg_reset_launch_count",
              "btslpg_last_kernel")

    def __init__(self, lib):
        self._lib = lib
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._lib, name)
            if name in self._QUIET or name.endswith("_workspace_bytes"):
                fn = raw
            else:
                import torch

                def fn(*a, _raw=raw, _name=name):
                    torch.cuda.nvtx.range_push(_name)
                    try:
                        return _raw(*a)
                    finally:
                        torch.cuda.nvtx.range_pop()
            self._cache[name] = fn
        return fn


_VALUE_ERRORS = (-1, -2, -3, -4, -5)


def check(status):
    if status == 0:
        return
    lib = load()
    msg = lib.btslpg_last_error().decode("utf-8", "replace")
    kind = lib.btslpg_status_string(status).decode()
    text = "libbtslpg: %s (%d): %s" % (kind, status, msg)
    if status in _VALUE_ERRORS:
        raise ValueError(text)
    raise RuntimeError(text)


# ---------------------------------------------------------------------------------------------
# tensor description
# ---------------------------------------------------------------------------------------------
_KDL_CPU, _KDL_CUDA = 1, 2


class TensorRef:
    """Keeps a BtsTensor struct and everything it points to alive for the duration of a call."""
    __slots__ = ("struct", "_shape", "_strides", "_owner")

    def __init__(self, struct, shape, strides, owner):
        self.struct, self._shape, self._strides, self._owner = struct, shape, strides, owner

    @property
    def ptr(self):
        return ctypes.pointer(self.struct)


def _torch_dtype_code(t):
    import torch
    if t.dtype == torch.float32:
        return 2, 32
    if t.dtype == torch.bfloat16:
        return 4, 16
    if t.dtype == torch.float16:
        return 2, 16
    if t.dtype == torch.float64:
        return 2, 64
    if t.dtype == torch.uint16:
        return 1, 16
    if t.dtype == torch.int16:         # carries the uint16 PNG image on hosts without an unsigned 16-bit dtype
        return 1, 16
    raise ValueError("unsupported tensor dtype %s" % (t.dtype,))


def from_torch(t):
    """Describe a torch tensor in place (zero copy)."""
    nd = t.dim()
    shape = (ctypes.c_int64 * max(nd, 1))(*t.shape)
    strides = (ctypes.c_int64 * max(nd, 1))(*t.stride())
    code, bits = _torch_dtype_code(t)
    if t.device.type == "cuda":
        dev = BtsDevice(_KDL_CUDA, t.device.index if t.device.index is not None else 0)
    else:
        dev = BtsDevice(_KDL_CPU, 0)
    s = BtsTensor(ctypes.c_void_p(t.data_ptr()), dev, nd, BtsDataType(code, bits, 1),
                  ctypes.cast(shape, ctypes.POINTER(ctypes.c_int64)), ctypes.cast(strides, ctypes.POINTER(ctypes.c_int64)), 0)
    return TensorRef(s, shape, strides, t)


_PyCapsule_GetPointer = ctypes.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = ctypes.c_void_p
_PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
_PyCapsule_IsValid = ctypes.pythonapi.PyCapsule_IsValid
_PyCapsule_IsValid.restype = ctypes.c_int
_PyCapsule_IsValid.argtypes = [ctypes.py_object, ctypes.c_char_p]


def from_dlpack_capsule(capsule, owner=None):
    """Borrow the DLTensor inside a DLPack capsule (zero copy).  The capsule is NOT consumed: it
    stays named "dltensor", so its own deleter releases the producer's memory when it is collected."""
    if not _PyCapsule_IsValid(capsule, b"dltensor"):
        raise ValueError("expected an unconsumed DLPack capsule named 'dltensor'")
    addr = _PyCapsule_GetPointer(capsule, b"dltensor")
    managed = ctypes.cast(addr, ctypes.POINTER(_DLManagedTensor)).contents
    src = managed.dl_tensor
    s = BtsTensor(src.data, src.device, src.ndim, src.dtype, src.shape, src.strides, src.byte_offset)
    return TensorRef(s, None, None, (capsule, owner))


def as_ref(obj):
    """torch.Tensor | DLPack capsule | object with __dlpack__ | None  ->  TensorRef | None."""
    if obj is None:
        return None
    if isinstance(obj, TensorRef):
        return obj
    try:
        import torch
        if isinstance(obj, torch.Tensor):
            return from_torch(obj)
    except ImportError:  # pragma: no cover
        pass
    if type(obj).__name__ == "PyCapsule":
        return from_dlpack_capsule(obj)
    if hasattr(obj, "__dlpack__"):
        return from_dlpack_capsule(obj.__dlpack__(), obj)
    raise TypeError("cannot describe %r as a BtsTensor" % (type(obj),))


def ptr_or_null(ref):
    return ref.ptr if ref is not None else None


def current_stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device`; NULL for a non-CUDA device, in which case
    the library itself rejects the host tensor (BTSLPG_EDEVICE) -- nothing is computed on the CPU."""
    import torch
    if device is not None and torch.device(device).type != "cuda":
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
