"""TEST INFRASTRUCTURE ONLY -- torch-CPU stand-ins for the tf.keras.layers classes
used by /root/reference/custom_layers.py and bts_decoder.py.

Tensors are NHWC torch tensors (Keras' default `channels_last`).  Every layer
created is appended to `CREATED` in creation order so a test can read the random
weights the reference graph was built with and hand the same weights to the
product code."""
import math as _math

import torch as _torch
import torch.nn.functional as _F

CREATED = []            # every layer instance, in creation order
_GEN = _torch.Generator().manual_seed(0)
DTYPE = _torch.float32  # weights dtype; tests may switch to float64 before building a graph


def reset(seed=0, dtype=_torch.float32):
    global DTYPE
    CREATED.clear()
    _GEN.manual_seed(seed)
    DTYPE = dtype


class Layer:
    """Minimal keras.layers.Layer protocol: name kwarg, lazy build() on first call,
    get_config() with the base keys."""
    _uid = {}

    def __init__(self, name=None, trainable=True, dtype=None, **kwargs):
        if kwargs:
            raise TypeError("unexpected Layer kwargs: %r" % (sorted(kwargs),))
        if name is None:
            base = _snake(type(self).__name__)
            n = Layer._uid.get(base, 0)
            Layer._uid[base] = n + 1
            name = base if n == 0 else "%s_%d" % (base, n)
        self.name = name
        self.trainable = trainable
        self._dtype = dtype or "float32"
        self.built = False
        CREATED.append(self)

    def build(self, input_shape):
        self.built = True

    def call(self, inputs, **kwargs):
        return inputs

    def __call__(self, inputs, **kwargs):
        if not self.built:
            if isinstance(inputs, (list, tuple)):
                shape = [tuple(t.shape) for t in inputs]
            else:
                shape = tuple(inputs.shape)
            self.build(shape)
            self.built = True
        self.last_input = inputs
        self.last_output = self.call(inputs, **kwargs)   # kept so a test can read intermediates
        return self.last_output

    def get_config(self):
        return {"name": self.name, "trainable": self.trainable, "dtype": self._dtype}


def _snake(s):
    out = []
    for i, ch in enumerate(s):
        if ch.isupper() and i and not s[i - 1].isupper():
            out.append("_")
        out.append(ch.lower())
    return "".join(out)


def _activation(name):
    if name is None:
        return lambda x: x
    if name == "elu":
        return _F.elu
    if name == "sigmoid":
        return _torch.sigmoid
    if name == "relu":
        return _F.relu
    raise ValueError(name)


class Conv2D(Layer):
    """Conv2D, NHWC, kernel HWIO, stride 1, padding='same', glorot_uniform init."""

    def __init__(self, filters, kernel_size, strides=1, padding="valid", dilation_rate=1,
                 activation=None, use_bias=True, **kwargs):
        super().__init__(**kwargs)
        assert strides == 1 and padding == "same" and not use_bias, "only what bts_decoder.py uses"
        self.filters, self.k, self.dil = filters, kernel_size, dilation_rate
        self.act = _activation(activation)
        self.kernel = None

    def build(self, input_shape):
        cin = input_shape[-1]
        fan_in, fan_out = self.k * self.k * cin, self.k * self.k * self.filters
        limit = _math.sqrt(6.0 / (fan_in + fan_out))
        w = _torch.rand((self.k, self.k, cin, self.filters), generator=_GEN, dtype=_torch.float64)
        self.kernel = ((w * 2 - 1) * limit).to(DTYPE).requires_grad_(True)

    def call(self, x):
        pad = self.dil * (self.k - 1) // 2
        y = _F.conv2d(x.permute(0, 3, 1, 2).contiguous(), self.kernel.permute(3, 2, 0, 1).contiguous(),
                      padding=pad, dilation=self.dil)
        return self.act(y.permute(0, 2, 3, 1))


class BatchNormalization(Layer):
    def __init__(self, momentum=0.99, epsilon=1e-3, fused=None, **kwargs):
        super().__init__(**kwargs)
        self.momentum, self.eps = momentum, epsilon

    def build(self, input_shape):
        c = input_shape[-1]
        self.gamma = _torch.ones(c, dtype=DTYPE, requires_grad=True)
        self.beta = _torch.zeros(c, dtype=DTYPE, requires_grad=True)
        self.moving_mean = _torch.zeros(c, dtype=DTYPE)
        self.moving_variance = _torch.ones(c, dtype=DTYPE)

    def call(self, x, training=False):
        if training:
            mean = x.mean(dim=(0, 1, 2))
            var = x.var(dim=(0, 1, 2), unbiased=False)
        else:
            mean, var = self.moving_mean, self.moving_variance
        return (x - mean) * _torch.rsqrt(var + self.eps) * self.gamma + self.beta


class UpSampling2D(Layer):
    def __init__(self, size=2, interpolation="nearest", **kwargs):
        super().__init__(**kwargs)
        assert interpolation == "nearest"
        self.size = size

    def call(self, x):
        x = _torch.repeat_interleave(x, self.size, dim=1)
        return _torch.repeat_interleave(x, self.size, dim=2)


class Concatenate(Layer):
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def call(self, xs):
        return _torch.cat(list(xs), dim=self.axis)


class ReLU(Layer):
    def call(self, x):
        return _F.relu(x)


class Lambda(Layer):
    def __init__(self, function, **kwargs):
        super().__init__(**kwargs)
        self.function = function

    def call(self, x):
        return self.function(x)
