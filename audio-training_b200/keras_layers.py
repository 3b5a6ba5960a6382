"""The reference's Keras layers as real `tf.keras.layers.Layer` subclasses over the C ABI (import-guarded).

The reference attaches `tfpcen.PCEN` inside the Keras model (audiomodel.py:789-793) and `badwinner2.MagTransform` as the first
layer of `build_model` (badwinner2.py:230): to drop in "unchanged" those call sites need objects Keras accepts as layers --
same class names, same weights (names, shapes, initial values, creation order: tfpcen.py:15-19, 48-87; badwinner2.py:36-45),
`get_config` for serialisation -- whose `call` runs the CUDA kernels.  This image has no TensorFlow, so nothing here imports
it at module load:

    import tensorflow as tf
    from audio_training_b200.keras_layers import make_layers
    L = make_layers(tf)              # -> namespace with ExponentialMovingAverage, PCEN, MagTransform
    x = L.PCEN()(x)                  # audiomodel.py:793

`make_layers(tf)` takes the `tf` module it should build on.  With real TensorFlow the tensors cross into the library
through DLPack inside a `tf.py_function` (`tf.experimental.dlpack.to_dlpack` -> `torch.utils.dlpack.from_dlpack`: the same
device memory, no copy) and PCEN's gradient is `cacfe_pcen_backward` behind `tf.custom_gradient`.  With the numpy stand-in
of the test tree (whose "tensors" are numpy arrays and whose `Layer` records `add_weight` calls) the same classes are
exercised end to end here -- which is what tests/ does; on a TensorFlow box the committed test additionally runs against
Keras itself when `import tensorflow` succeeds.
"""
from __future__ import annotations

import types

import numpy as np

from . import _runtime as rt


def _bridge(tf):
    """(to_torch, from_torch, wrap) for the tensor type of this `tf`."""
    import torch
    dl = getattr(getattr(tf, "experimental", None), "dlpack", None)
    if dl is not None and hasattr(tf, "py_function"):
        def to_torch(x):
            return torch.utils.dlpack.from_dlpack(dl.to_dlpack(x))

        def from_torch(t):
            return dl.from_dlpack(torch.utils.dlpack.to_dlpack(t.contiguous()))

        def wrap(fn, inputs, n_out=1):
            """run `fn` (torch CUDA tensors in / out) as an eager island of the graph"""
            def island(*xs):
                outs = fn(*[to_torch(x) for x in xs])
                outs = outs if isinstance(outs, (tuple, list)) else (outs,)
                return [from_torch(o) for o in outs]
            res = tf.py_function(island, inputs, [tf.float32] * n_out)
            return res[0] if n_out == 1 else res
        return to_torch, from_torch, wrap

    def to_torch(x):       # the numpy stand-in
        return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).cuda()

    def from_torch(t):
        return t.cpu().numpy()

    def wrap(fn, inputs, n_out=1):
        outs = fn(*[to_torch(x) for x in inputs])
        outs = outs if isinstance(outs, (tuple, list)) else (outs,)
        outs = [from_torch(o) for o in outs]
        return outs[0] if n_out == 1 else outs
    return to_torch, from_torch, wrap


def _scalar(v):
    return float(np.asarray(v).reshape(-1)[0])


def make_layers(tf):
    """Build the three layer classes on `tf` (real TensorFlow, or the numpy stand-in the tests use)."""
    Layer = tf.keras.layers.Layer
    Constant = tf.keras.initializers.Constant
    register = getattr(getattr(tf.keras, "utils", None), "register_keras_serializable", None) or (lambda **_k: (lambda c: c))
    _, _, wrap = _bridge(tf)
    custom_gradient = getattr(tf, "custom_gradient", None)

    def plan_for(t):
        return rt.get_plan(rt.FrontendConfig(), t.device.index)

    @register(package="CacfeLayers", name="ExponentialMovingAverage")
    class ExponentialMovingAverage(Layer):
        """tfpcen.py:8-39: weight `smooth` [1] = coeff_init; call(inputs [batch, seq, filters], initial_state)."""

        def __init__(self, coeff_init, trainable=False, **kwargs):
            kwargs.pop("name", None)
            super().__init__(name="EMA", **kwargs)
            self._coeff_init = coeff_init
            self._trainable = trainable
            self._weights_var = self.add_weight(name="smooth", shape=[1], initializer=Constant(self._coeff_init),
                                                trainable=self._trainable)

        def call(self, inputs, initial_state=None):
            w = _scalar(self._weights_var)
            if initial_state is None:   # the reference's only use passes inputs[:, 0, :] (tfpcen.py:92): the kernel's default
                return wrap(lambda x: plan_for(x).ema(x.contiguous(), w, 1), [inputs])
            return wrap(lambda x, s0: plan_for(x).ema(x.contiguous(), w, 1, initial_state=s0.contiguous()), [inputs, initial_state])

        def get_config(self):
            return {"coeff_init": self._coeff_init, "trainable": self._trainable}

    @register(package="CacfeLayers", name="PCEN")
    class PCEN(Layer):
        """tfpcen.py:42-99: weights gain 0.98, bias 2.0, root 2.0, EMA/smooth 0.04, a-power -1.0 (declared, unused: Q12) in that
        creation order.  Rank 3 [batch, time, filters] is the reference's contract; the rank-4 image (audiomodel.py:793) runs the
        smoother along axis 2 (our documented extension, Q13).  The reference registers this class under MagTransform's
        serialisation key (Q12); here it gets its own."""

        def __init__(self, norm_scope="tensor", **kwargs):
            super().__init__(**kwargs)
            self.gain = self.add_weight(initializer=Constant(value=0.98), name="gain", dtype="float32", shape=[1], trainable=True)
            self.bias = self.add_weight(initializer=Constant(value=2.0), name="bias", dtype="float32", shape=[1], trainable=True)
            self.root = self.add_weight(initializer=Constant(value=2.0), name="root", dtype="float32", shape=[1], trainable=True)
            self.eps = 1e-6
            self.ema = ExponentialMovingAverage(coeff_init=0.04, trainable=True)
            self.a = self.add_weight(initializer=Constant(value=-1.0), name="a-power", dtype="float32", shape=[1], trainable=True)
            self.norm_scope = norm_scope

        def _params(self, gain, bias, root, smooth):
            return rt.pcen_params(gain, bias, root, smooth, self.eps, self.norm_scope)

        def call(self, inputs):
            rank = len(inputs.shape)
            if rank not in (3, 4):
                raise ValueError("PCEN: expected [batch, time, filters] (or the rank-4 image extension)")
            axis = 1 if rank == 3 else 2

            def forward(x, gain, bias, root, smooth):
                p = self._params(_scalar(gain.cpu()), _scalar(bias.cpu()), _scalar(root.cpu()), _scalar(smooth.cpu()))
                return plan_for(x).pcen(x.contiguous(), p, axis)

            def backward(x, gain, bias, root, smooth, dy):
                p = self._params(_scalar(gain.cpu()), _scalar(bias.cpu()), _scalar(root.cpu()), _scalar(smooth.cpu()))
                dx, dp = plan_for(x).pcen_backward(x.contiguous(), dy.contiguous(), p, axis)
                return dx, dp[0:1].clone(), dp[1:2].clone(), dp[2:3].clone(), dp[3:4].clone()

            args = [inputs, self.gain, self.bias, self.root, self.ema._weights_var]
            if custom_gradient is None:                       # the numpy stand-in: forward only
                return wrap(forward, args)

            @custom_gradient
            def op(x, gain, bias, root, smooth):
                y = wrap(forward, [x, gain, bias, root, smooth])
                y.set_shape(x.shape)

                def grad(dy):
                    return wrap(backward, [x, gain, bias, root, smooth, dy], 5)
                return y, grad
            return op(*[tf.convert_to_tensor(a) for a in args])

        def get_config(self):
            base = super().get_config() if hasattr(Layer, "get_config") else {}
            return {**base, "norm_scope": self.norm_scope}

    @register(package="CacfeLayers", name="MagTransform")
    class MagTransform(Layer):
        """badwinner2.py:32-49: x ** sigmoid(a), weight `a-power` [1] = -1 constrained to [-2, 1]."""

        def __init__(self, **kwargs):
            super().__init__(**kwargs)
            constraint = None
            mm = getattr(getattr(tf.keras, "constraints", None), "MinMaxNorm", None)
            if mm is not None:
                constraint = mm(min_value=-2.0, max_value=1.0, rate=1.0, axis=-1)
            self.a = self.add_weight(initializer=Constant(value=-1.0), name="a-power", dtype="float32", shape=[1], trainable=True,
                                     constraint=constraint)

        def call(self, inputs):
            def exponent(a):
                return float(np.float32(1.0) / (np.float32(1.0) + np.exp(-np.float32(_scalar(a.cpu())), dtype=np.float32)))

            def forward(x, a):
                return plan_for(x).compress(x.contiguous(), "mag_pow", exponent(a))

            def backward(x, a, dy):
                # y = x^e, e = sigmoid(a):  dy/dx = e x^(e-1),  dy/da = x^e ln(x) e (1 - e).  Plain elementwise torch ops: the
                # gradient of a one-parameter power is not on the measured path
                import torch
                e = exponent(a)
                xe = torch.pow(x, e)
                dx = dy * e * torch.pow(x, e - 1.0)
                lnx = torch.where(x > 0, torch.log(x), torch.zeros_like(x))
                da = (dy * xe * lnx).sum().reshape(1) * (e * (1.0 - e))
                return dx, da

            if custom_gradient is None:
                return wrap(forward, [inputs, self.a])

            @custom_gradient
            def op(x, a):
                y = wrap(forward, [x, a])
                y.set_shape(x.shape)
                return y, (lambda dy: wrap(backward, [x, a, dy], 2))
            return op(tf.convert_to_tensor(inputs), tf.convert_to_tensor(self.a))

    return types.SimpleNamespace(ExponentialMovingAverage=ExponentialMovingAverage, PCEN=PCEN, MagTransform=MagTransform)
