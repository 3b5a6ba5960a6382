"""A numpy stand-in for the slice of the `tf` namespace that the reference's *feature
functions* touch  --  TEST INFRASTRUCTURE ONLY (golden-vector generation).

TensorFlow is not installable in the build container (no wheel, no network), so
`gen_golden.py` executes the reference's own Python source (tfpcen.py, and the feature
functions of tfdataset.py / badwinner2.py / predict_utils.py) against this module instead.
That pins the reference's Python-level arithmetic (operation order, constants, axis
choices).  What is NOT pinned is TensorFlow's own internals, restated here from its
documented behaviour:

  tf.signal.frame(pad_end=True)  -> ceil(N/step) frames, zero padded at the end
  tf.signal.hann_window          -> periodic Hann, evaluated in float32
  tf.signal.stft                 -> rfft(frame * window), complex64
  tf.scan(fn, elems, init)       -> acc = fn(acc, elems[t]) sequentially, all accs stacked

Every op keeps float32 inputs in float32 (numpy NEP-50 promotion), like TF does.
"""
from __future__ import annotations

import types

import numpy as np
import scipy.fft as sfft


def _f32(x):
    return np.asarray(x, dtype=np.float32) if not np.iscomplexobj(x) else np.asarray(x)


# ---- tf.signal ------------------------------------------------------------------------
def hann_window(window_length, periodic=True, dtype=np.float32, name=None):
    """tf.signal.hann_window (window_ops._raised_cosine_window): every step in `dtype` -- the sample index, 2 pi, the
    division and the cosine are float32 operations, which is what TF executes (round 1 evaluated in float64 and rounded
    once: up to 1 ulp of float32 apart from this)."""
    dt = np.dtype(dtype).type
    denom = window_length if (periodic and window_length % 2 == 0) else window_length - 1
    count = np.arange(window_length).astype(dt)
    cos_arg = dt(2.0 * np.pi) * count / dt(denom)
    return (dt(0.5) - dt(0.5) * np.cos(cos_arg)).astype(dt)


def frame(signal, frame_length, frame_step, pad_end=False, pad_value=0, axis=-1, name=None):
    signal = np.asarray(signal)
    n = signal.shape[-1]
    if pad_end:
        t = -(-n // frame_step)
        need = max(0, frame_length + frame_step * (t - 1) - n)
        signal = np.concatenate(
            [signal, np.full(signal.shape[:-1] + (need,), pad_value, dtype=signal.dtype)], axis=-1)
    else:
        t = 0 if n < frame_length else 1 + (n - frame_length) // frame_step
    idx = frame_step * np.arange(t)[:, None] + np.arange(frame_length)[None, :]
    return signal[..., idx]


def stft(signals, frame_length, frame_step, fft_length=None, window_fn=hann_window,
         pad_end=False, name=None):
    fft_length = fft_length or frame_length
    frames = frame(_f32(signals), frame_length, frame_step, pad_end)
    if window_fn is not None:
        frames = frames * window_fn(frame_length, dtype=np.float32)
    return sfft.rfft(frames.astype(np.float32), n=fft_length, axis=-1).astype(np.complex64)


signal = types.SimpleNamespace(stft=stft, frame=frame, hann_window=hann_window)


# ---- tf.math / top level ---------------------------------------------------------------
def _reduce(fn):
    def op(x, axis=None, keepdims=False, name=None):
        return fn(np.asarray(x), axis=axis, keepdims=keepdims)
    return op


reduce_min = _reduce(np.min)
reduce_max = _reduce(np.max)
reduce_mean = _reduce(np.mean)


def reduce_std(x, axis=None, keepdims=False, name=None):
    return np.std(np.asarray(x), axis=axis, keepdims=keepdims)


def _pow(x, y, name=None):
    return np.power(x, np.asarray(y, dtype=np.asarray(x).real.dtype) if np.isscalar(y) else y)


def sigmoid(x, name=None):
    x = _f32(x)
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x))).astype(np.float32)


math = types.SimpleNamespace(
    reduce_min=reduce_min, reduce_max=reduce_max, reduce_mean=reduce_mean, reduce_std=reduce_std,
    subtract=lambda a, b, name=None: np.subtract(a, b),
    divide=lambda a, b, name=None: np.divide(a, b),
    multiply=lambda a, b, name=None: np.multiply(a, b),
    pow=_pow, abs=lambda x, name=None: np.abs(x),
    maximum=lambda a, b, name=None: np.maximum(a, np.asarray(b, dtype=np.asarray(a).dtype) if np.isscalar(b) else b),
    minimum=lambda a, b, name=None: np.minimum(a, np.asarray(b, dtype=np.asarray(a).dtype) if np.isscalar(b) else b),
    sigmoid=sigmoid, log=lambda x, name=None: np.log(x),
    equal=lambda a, b, name=None: np.equal(a, b), reduce_all=lambda x, name=None: np.all(x),
)


def transpose(a, perm=None, name=None):
    return np.transpose(a, perm)


def expand_dims(x, axis, name=None):
    return np.expand_dims(np.asarray(x), axis)


def repeat(x, repeats, axis=None, name=None):
    return np.repeat(np.asarray(x), repeats, axis=axis)


def reduce_mean(x, axis=None, keepdims=False, name=None):
    x = np.asarray(x)
    return np.mean(x, axis=axis, keepdims=keepdims, dtype=x.dtype)   # TF keeps the input dtype (numpy: pairwise f32 sum)


def constant(value, dtype=None, name=None):
    return np.asarray(value) if dtype is None else np.asarray(value, dtype=dtype)


def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    t = np.asarray(t)
    return np.clip(t, t.dtype.type(clip_value_min), t.dtype.type(clip_value_max))


def gather(params, indices, axis=0, name=None):
    return np.take(np.asarray(params), indices, axis=axis)


def scan(fn, elems, initializer=None, name=None):
    elems = np.asarray(elems)
    acc = initializer
    outs = []
    for t in range(elems.shape[0]):
        acc = fn(acc, elems[t])
        outs.append(acc)
    return np.stack(outs, axis=0)


def tensordot(a, b, axes, name=None):
    return np.tensordot(np.asarray(a), np.asarray(b), axes)


def reshape(x, shape, name=None):
    return np.reshape(np.asarray(x), shape)


def concat(values, axis, name=None):
    return np.concatenate(values, axis=axis)


def squeeze(x, axis=None, name=None):
    return np.squeeze(x, axis=axis)


def numpy_function(func, inp, Tout, name=None):
    """tf.numpy_function: call a Python function on the (numpy) values and cast the result to Tout."""
    return np.asarray(func(*[np.asarray(v) if isinstance(v, np.ndarray) else v for v in inp]), dtype=Tout)


def function(fn=None, **_kw):  # @tf.function -> eager
    if fn is None:
        return lambda f: f
    return fn


Tensor = np.ndarray
float32 = np.float32


# ---- tf.keras ---------------------------------------------------------------------------
class _Constant:
    def __init__(self, value=0.0):
        self.value = value


class _MinMaxNorm:
    def __init__(self, min_value=0.0, max_value=1.0, rate=1.0, axis=0):
        self.min_value, self.max_value, self.rate, self.axis = min_value, max_value, rate, axis


class Layer:
    """Enough of keras.layers.Layer for add_weight(...) + __call__ -> call."""

    def __init__(self, name=None, **kwargs):
        self.name = name
        self._added = []  # creation order, like Keras' weights list

    def add_weight(self, name=None, shape=None, initializer=None, trainable=True, dtype="float32",
                   constraint=None, **_kw):
        value = np.full(shape, initializer.value if isinstance(initializer, _Constant) else 0.0,
                        dtype=np.float32)
        self._added.append((name, value))
        return value

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


def _register_keras_serializable(package="Custom", name=None):
    def deco(cls):
        cls._serial_key = f"{package}>{name or cls.__name__}"
        return cls
    return deco


def _batch_dot(x, y, axes=None):
    # keras.backend.batch_dot on [B,M,K] x [B,K,T]: per-batch matmul (float32 sgemm on CPU)
    return np.matmul(np.asarray(x), np.asarray(y))


keras = types.SimpleNamespace(
    layers=types.SimpleNamespace(Layer=Layer),
    Layer=Layer,
    initializers=types.SimpleNamespace(Constant=_Constant),
    constraints=types.SimpleNamespace(MinMaxNorm=_MinMaxNorm),
    utils=types.SimpleNamespace(register_keras_serializable=_register_keras_serializable),
    backend=types.SimpleNamespace(batch_dot=_batch_dot, epsilon=lambda: 1e-7),
    ops=types.SimpleNamespace(
        shape=lambda x: np.asarray(x).shape,
        log10=lambda x: np.log10(x),
    ),
)

compat = types.SimpleNamespace(v1=types.SimpleNamespace(identity=lambda x: np.array(x, copy=True)))
data = types.SimpleNamespace(AUTOTUNE=-1)
