"""TEST INFRASTRUCTURE ONLY -- torch-CPU stand-ins for the tf.keras.backend
functions used by /root/reference/custom_layers.py and bts.py.  Semantics follow
the TensorFlow documentation for each function (cited inline)."""
import torch as _torch


def epsilon():
    # K.epsilon() default fuzz factor
    return 1e-7


def sin(x):
    return _torch.sin(x)


def cos(x):
    return _torch.cos(x)


def log(x):
    return _torch.log(x)


def sqrt(x):
    return _torch.sqrt(x)


def square(x):
    return x * x


def mean(x):
    return _torch.mean(x)


def greater(x, y):
    return x > y


def concatenate(tensors, axis=-1):
    return _torch.cat(list(tensors), dim=axis)


def stack(tensors, axis=0):
    return _torch.stack(list(tensors), dim=axis)


def expand_dims(x, axis=-1):
    return _torch.unsqueeze(x, axis)


def ones_like(x):
    return _torch.ones_like(x)


def repeat_elements(x, rep, axis):
    # K.repeat_elements == np.repeat along `axis` (each element repeated `rep` times in place)
    return _torch.repeat_interleave(x, rep, dim=axis)


def sum(x, axis=None, keepdims=False):  # noqa: A001 - mirrors the Keras name
    if axis is None:
        return _torch.sum(x)
    return _torch.sum(x, dim=axis, keepdim=keepdims)


def l2_normalize(x, axis=None):
    # tf.math.l2_normalize: x * rsqrt(max(sum(x**2, axis, keepdims=True), 1e-12))
    square_sum = _torch.sum(x * x, dim=axis, keepdim=True)
    x_inv_norm = _torch.rsqrt(_torch.clamp_min(square_sum, 1e-12))
    return x * x_inv_norm
