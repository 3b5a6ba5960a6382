"""An eager numpy stand-in for the slice of `tf.keras` that the reference's two CNN builders use
(badwinner2.build_model, badwinner2.py:212-324; resnet/wr_resnet_bird.WRResNet, resnet/wr_resnet_bird.py:7-178)
--  TEST INFRASTRUCTURE ONLY (golden-vector generation for the consumers).

Keras cannot be installed here, so `gen_consumer_golden.py` executes the reference's own model-building functions against
this module.  "Functional" Keras is run eagerly: `Input(shape)` returns the actual batch (NHWC, float64), every layer
call computes its result at once, and every variable a layer creates is appended to REGISTRY in creation order -- which
is the order of `model.weights` in Keras (layers in graph-creation order; per layer: kernel, bias / gamma, beta,
moving_mean, moving_variance).  Values are drawn from a seeded generator (`draw`), not from the Keras initialisers: what
is pinned is the layer graph (kernel sizes, padding, pooling, axes, the `filters=X.shape[1]` quirk, which tensors are
added), in inference mode (Dropout = identity, BatchNormalization with its moving statistics).

Semantics restated from the Keras / TensorFlow documentation:
  Conv2D            NHWC, kernel HWIO, stride 1, "valid" or "same" (total pad k - 1: floor before, the rest after)
  MaxPool2D         pool_size, strides = pool_size, "valid"
  AveragePooling2D  "same": windows clipped at the border, padding excluded from the mean
  BatchNormalization(axis, scale, center), epsilon 1e-3:  gamma (x - mean) / sqrt(var + eps) + beta
  LeakyReLU(alpha)  default alpha 0.3 (the keyword the reference passes is `alpha`)
  Dense             on the last axis;  GlobalAveragePooling2D: mean over H and W
  tfp.math.reduce_logmeanexp(x, axis) = log(mean(exp(x), axis))
"""
from __future__ import annotations

import types

import numpy as np

REGISTRY = []          # [(kind, layer name, variable name, array)] in creation order
_RNG = [None]
_COUNTS = {}


def reset(seed):
    REGISTRY.clear()
    _COUNTS.clear()
    _RNG[0] = np.random.default_rng(seed)


def draw(rng, var, shape):
    """The seeded value of a variable: the golden generator and the consumer test both call this, in creation order."""
    shape = tuple(int(s) for s in shape)
    if var == "kernel":
        fan_in = int(np.prod(shape[:-1]))
        lim = np.sqrt(3.0 / fan_in)                       # unit-gain uniform: activations keep their scale through the stack
        return rng.uniform(-lim, lim, shape)
    if var == "bias":
        return rng.uniform(-0.1, 0.1, shape)
    if var in ("gamma", "moving_variance"):
        return rng.uniform(0.5, 1.5, shape)
    if var in ("beta", "moving_mean"):
        return rng.uniform(-0.2, 0.2, shape)
    raise KeyError(var)


def _auto_name(kind):
    n = _COUNTS.get(kind, 0)
    _COUNTS[kind] = n + 1
    return kind if n == 0 else f"{kind}_{n}"


def _new(kind, layer, var, shape):
    value = draw(_RNG[0], var, shape)
    REGISTRY.append((kind, layer, var, value))
    return value


# ---- layers ------------------------------------------------------------------------------------------------------
def _pair(v):
    return (int(v), int(v)) if np.isscalar(v) else (int(v[0]), int(v[1]))


def _windows(x, kh, kw, sh=1, sw=1):
    """[B, H, W, C] -> view [B, Ho, Wo, C, kh, kw]"""
    v = np.lib.stride_tricks.sliding_window_view(x, (kh, kw), axis=(1, 2))
    return v[:, ::sh, ::sw]


class Conv2D:
    def __init__(self, filters, kernel_size, strides=1, padding="valid", name=None, kernel_initializer=None, **_kw):
        assert _pair(strides) == (1, 1)
        self.filters, self.k, self.padding = int(filters), _pair(kernel_size), padding
        self.name = name or _auto_name("conv2d")

    def __call__(self, x):
        kh, kw = self.k
        kernel = _new("conv2d", self.name, "kernel", (kh, kw, x.shape[-1], self.filters))
        bias = _new("conv2d", self.name, "bias", (self.filters,))
        if self.padding == "same":
            ph, pw = kh - 1, kw - 1
            x = np.pad(x, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)))
        else:
            assert self.padding == "valid"
        win = _windows(x, kh, kw)                                   # [B, Ho, Wo, C, kh, kw]
        return np.einsum("bhwcij,ijcf->bhwf", win, kernel, optimize=True) + bias


class BatchNormalization:
    def __init__(self, axis=-1, scale=True, center=True, epsilon=1e-3, name=None, **_kw):
        self.axis, self.scale, self.center, self.eps = axis, scale, center, epsilon
        self.name = name or _auto_name("batch_normalization")

    def __call__(self, x):
        ax = self.axis % x.ndim
        shape = (x.shape[ax],)
        gamma = _new("bn", self.name, "gamma", shape) if self.scale else 1.0
        beta = _new("bn", self.name, "beta", shape) if self.center else 0.0
        mean = _new("bn", self.name, "moving_mean", shape)
        var = _new("bn", self.name, "moving_variance", shape)
        bshape = [1] * x.ndim
        bshape[ax] = -1
        r = lambda v: np.reshape(v, bshape) if isinstance(v, np.ndarray) else v
        return r(gamma) * (x - r(mean)) / np.sqrt(r(var) + self.eps) + r(beta)


class LeakyReLU:
    def __init__(self, alpha=0.3, negative_slope=None, **_kw):
        self.alpha = alpha if negative_slope is None else negative_slope

    def __call__(self, x):
        return np.where(x >= 0, x, self.alpha * x)


class MaxPool2D:
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", **_kw):
        self.p = _pair(pool_size)
        self.s = self.p if strides is None else _pair(strides)
        assert padding == "valid"

    def __call__(self, x):
        return _windows(x, *self.p, *self.s).max(axis=(-2, -1))


class AveragePooling2D:
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", **_kw):
        self.p = _pair(pool_size)
        self.s = self.p if strides is None else _pair(strides)
        self.padding = padding

    def __call__(self, x):
        (ph, pw), (sh, sw) = self.p, self.s
        if self.padding == "valid":
            return _windows(x, ph, pw, sh, sw).mean(axis=(-2, -1))
        B, H, W, C = x.shape
        Ho, Wo = -(-H // sh), -(-W // sw)
        th, tw = max((Ho - 1) * sh + ph - H, 0), max((Wo - 1) * sw + pw - W, 0)
        pad = ((0, 0), (th // 2, th - th // 2), (tw // 2, tw - tw // 2), (0, 0))
        num = _windows(np.pad(x, pad), ph, pw, sh, sw).sum(axis=(-2, -1))
        cnt = _windows(np.pad(np.ones((1, H, W, 1)), pad), ph, pw, sh, sw).sum(axis=(-2, -1))
        return num / cnt


class Dropout:
    def __init__(self, rate=0.5, **_kw):
        pass

    def __call__(self, x, training=False):
        return x


class Activation:
    def __init__(self, fn, **_kw):
        assert fn == "relu"

    def __call__(self, x):
        return np.maximum(x, 0.0)


class Identity:
    def __call__(self, x):
        return x


class Add:
    def __call__(self, xs):
        assert xs[0].shape == xs[1].shape, (xs[0].shape, xs[1].shape)
        return xs[0] + xs[1]


class GlobalAveragePooling2D:
    def __call__(self, x):
        return x.mean(axis=(1, 2))


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _softmax(x, axis=-1):
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


class Dense:
    def __init__(self, units, activation=None, name=None, **_kw):
        self.units, self.activation = int(units), activation
        self.name = name or _auto_name("dense")

    def __call__(self, x):
        kernel = _new("dense", self.name, "kernel", (x.shape[-1], self.units))
        bias = _new("dense", self.name, "bias", (self.units,))
        y = x @ kernel + bias
        return _sigmoid(y) if self.activation == "sigmoid" else y


class Layer:
    """keras.layers.Layer for the reference's own subclasses (MagTransform): add_weight + __call__ -> call."""

    def __init__(self, name=None, **_kw):
        self.name = name or _auto_name(type(self).__name__.lower())

    def add_weight(self, name=None, shape=None, initializer=None, **_kw):
        value = np.full(tuple(shape), float(initializer.value), dtype=np.float64)
        REGISTRY.append(("custom", self.name, name, value))
        return value

    def __call__(self, *a, **k):
        return self.call(*a, **k)


class _Constant:
    def __init__(self, value=0.0):
        self.value = value


class Model:
    def __init__(self, inputs=None, outputs=None, name=None):
        self.inputs, self.outputs, self.name = inputs, outputs, name

    def summary(self):
        pass


def make_tf(batch):
    """The `tf` (and `tfp`) objects the model builders see; `Input(...)` hands out `batch`."""
    def Input(shape=None, name=None, **_kw):
        assert tuple(batch.shape[1:]) == tuple(shape), (batch.shape, shape)
        return batch

    layers = types.SimpleNamespace(Conv2D=Conv2D, BatchNormalization=BatchNormalization, LeakyReLU=LeakyReLU, MaxPool2D=MaxPool2D,
                                   MaxPooling2D=MaxPool2D, AveragePooling2D=AveragePooling2D, Dropout=Dropout, Activation=Activation,
                                   Identity=Identity, Add=Add, GlobalAveragePooling2D=GlobalAveragePooling2D, Dense=Dense, Layer=Layer)
    init = lambda *a, **k: None
    keras = types.SimpleNamespace(
        Input=Input, layers=layers, Layer=Layer, Model=Model, models=types.SimpleNamespace(Model=Model),
        activations=types.SimpleNamespace(sigmoid=_sigmoid, softmax=_softmax),
        initializers=types.SimpleNamespace(Orthogonal=init, GlorotUniform=init, Constant=_Constant),
        constraints=types.SimpleNamespace(MinMaxNorm=init),
        utils=types.SimpleNamespace(register_keras_serializable=lambda package="Custom", name=None: (lambda cls: cls)))
    math = types.SimpleNamespace(pow=lambda x, y: np.power(x, y), sigmoid=_sigmoid)
    tf = types.SimpleNamespace(keras=keras, math=math)
    tfp = types.SimpleNamespace(math=types.SimpleNamespace(
        reduce_logmeanexp=lambda x, axis=None, keepdims=False: np.log(np.mean(np.exp(x), axis=axis, keepdims=keepdims))))
    return tf, tfp
