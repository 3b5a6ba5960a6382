"""Drop-in for `badwinner2.MagTransform` (badwinner2.py:32-49): the learned root compression x ** sigmoid(a) that
badwinner2 models use instead of PCEN (badwinner2.py:142,230)."""
from __future__ import annotations

import math

import numpy as np

from . import _runtime as rt


class MagTransform:
    serial_key = "MyLayers>MagTransform"

    def __init__(self, **kwargs):
        self.name = kwargs.get("name", "mag_transform")
        self.a = np.full([1], -1.0, dtype=np.float32)  # weight 'a-power', constrained to [-2, 1] in training

    def state_dict(self):
        return {"a-power": self.a.copy()}

    def load_state_dict(self, state):
        self.a[:] = state["a-power"]

    def exponent(self):
        return float(np.float32(1.0) / (np.float32(1.0) + np.exp(-self.a[0], dtype=np.float32)))

    def call(self, inputs):
        t, restore = rt.to_device(inputs)
        plan = rt.get_plan(rt.FrontendConfig(), t.device.index)
        return restore(plan.compress(t, "mag_pow", self.exponent()))

    __call__ = call
