"""Drop-in for the reference's `tfpcen` module (tfpcen.py:8-110): ExponentialMovingAverage, PCEN, normalize_minmax.

The layers keep the reference's constructor arguments, weight names, creation order and initial values so
checkpoints map one to one; `__call__` runs the CUDA kernels.  `PCENTrainable` is the same layer as a torch module whose
backward also runs on our kernels (cacfe_pcen_backward) -- for callers that train the front-end's four scalars.
"""
from __future__ import annotations

import numpy as np

from . import _runtime as rt


def _plan(device):
    return rt.get_plan(rt.FrontendConfig(), device)


class ExponentialMovingAverage:
    """tfpcen.py:8-39.  call(inputs[batch, seq, filters], initial_state): M[t] = w x[t] + (1-w) M[t-1] with
    w = clip(smooth, 0, 1), M[-1] = initial_state (tf.scan's initializer); None = inputs[:, 0, :], what the reference's only
    call site passes (tfpcen.py:92)."""

    def __init__(self, coeff_init, trainable=False):
        self.name = "EMA"
        self._coeff_init = coeff_init
        self._trainable = trainable
        self._weights = np.full([1], coeff_init, dtype=np.float32)  # weight name 'smooth'

    @property
    def weights(self):
        return {"smooth": self._weights}

    def call(self, inputs, initial_state=None, time_axis=1):
        t, restore = rt.to_device(inputs)
        first = None
        if initial_state is not None:   # tf.scan's initializer; PCEN.call passes inputs[:, 0, :] (tfpcen.py:92)
            first, _ = rt.to_device(initial_state)
        return restore(_plan(t.device.index).ema(t, float(self._weights[0]), time_axis, initial_state=first))

    __call__ = call


class PCEN:
    """tfpcen.py:42-99.  Weights in creation order: gain 0.98, bias 2.0, root 2.0, EMA/smooth 0.04, a-power -1.0
    (declared and never used by call, Q12).  Input contract: rank 3 [batch, time, filters]; a rank-4 image
    [batch, mels, time, channels] (how audiomodel.py:793 attaches the layer, Q13) is handled as our documented
    extension: the smoother runs along the time axis of every (mel, channel) row.

    norm_scope: "tensor" = the reference's tensor-global min-max (tfpcen.py:105-110); "clip" / "none" are ours."""

    serial_key = "MyLayers>PCEN"  # the reference registers this class under MagTransform's key (Q12); we do not

    def __init__(self, norm_scope="tensor", **kwargs):
        self.name = kwargs.get("name", "pcen")
        self.gain = np.full([1], 0.98, dtype=np.float32)
        self.bias = np.full([1], 2.0, dtype=np.float32)
        self.root = np.full([1], 2.0, dtype=np.float32)
        self.eps = 1e-6
        self.ema = ExponentialMovingAverage(coeff_init=0.04, trainable=True)
        self.a = np.full([1], -1.0, dtype=np.float32)  # 'a-power'
        self.norm_scope = norm_scope

    def state_dict(self):
        return {"gain": self.gain.copy(), "bias": self.bias.copy(), "root": self.root.copy(),
                "EMA/smooth": self.ema._weights.copy(), "a-power": self.a.copy()}

    def load_state_dict(self, state):
        self.gain[:] = state["gain"]
        self.bias[:] = state["bias"]
        self.root[:] = state["root"]
        self.ema._weights[:] = state["EMA/smooth"]
        if "a-power" in state:
            self.a[:] = state["a-power"]

    def params(self):
        return rt.pcen_params(self.gain[0], self.bias[0], self.root[0], self.ema._weights[0], self.eps, self.norm_scope)

    def call(self, inputs):
        t, restore = rt.to_device(inputs)
        if t.dim() == 3:
            axis = 1
        elif t.dim() == 4:
            axis = 2
        else:
            raise ValueError("PCEN: expected [batch, time, filters] (or the rank-4 image extension)")
        return restore(_plan(t.device.index).pcen(t, self.params(), axis))

    __call__ = call


class _PCENFunction:
    """torch.autograd.Function built lazily (torch is imported by _runtime already; kept out of module import order)."""
    _fn = None

    @classmethod
    def get(cls):
        if cls._fn is not None:
            return cls._fn
        import torch

        class PCENFunction(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, gain, bias, root, smooth, eps, norm_scope, time_axis):
                plan = _plan(x.device.index)
                p = rt.pcen_params(float(gain), float(bias), float(root), float(smooth), eps, norm_scope)
                ctx.save_for_backward(x)
                ctx.cfg = (p, time_axis, plan)
                return plan.pcen(x, p, time_axis)

            @staticmethod
            def backward(ctx, grad_out):
                (x,) = ctx.saved_tensors
                p, time_axis, plan = ctx.cfg
                dx, dp = plan.pcen_backward(x, grad_out.contiguous(), p, time_axis)
                return dx, dp[0:1], dp[1:2], dp[2:3], dp[3:4], None, None, None

        cls._fn = PCENFunction
        return PCENFunction


def PCENTrainable(norm_scope="tensor"):
    """tfpcen.PCEN as a torch.nn.Module: parameters gain / bias / root / smooth (+ the unused a_power, Q12) with the
    reference's initial values; forward and backward both on the CUDA kernels."""
    import torch

    class _PCENTrainable(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gain = torch.nn.Parameter(torch.full((1,), 0.98))
            self.bias = torch.nn.Parameter(torch.full((1,), 2.0))
            self.root = torch.nn.Parameter(torch.full((1,), 2.0))
            self.smooth = torch.nn.Parameter(torch.full((1,), 0.04))
            self.a_power = torch.nn.Parameter(torch.full((1,), -1.0))
            self.eps = 1e-6
            self.norm_scope = norm_scope

        def forward(self, x):
            if x.dim() not in (3, 4):
                raise ValueError("PCEN: expected [batch, time, filters] (or the rank-4 image extension)")
            return _PCENFunction.get().apply(x.contiguous(), self.gain, self.bias, self.root, self.smooth, self.eps,
                                             self.norm_scope, 1 if x.dim() == 3 else 2)

    return _PCENTrainable()


def normalize_minmax(data):
    """tfpcen.py:105-110: 2 * ((x - min) / (max - min)) - 1 with min/max over the whole tensor."""
    t, restore = rt.to_device(data)
    return restore(_plan(t.device.index).compress(t, "minmax"))
