#!/usr/bin/env python3
"""Seed sweep of the PCEN-backward parity check (tests/test_gpu_parity.py::test_pcen_backward): which case / which term comes
closest to its tolerance.  Prints the worst ratio error / allowance per case over the seeds, for dx and the four parameters."""
import os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from audio_training_b200 import _runtime as rt
from oracle import frontend_oracle as oracle

CASES = [((3, 70, 40), 1, "tensor", {}), ((3, 70, 40), 1, "clip", {}), ((3, 70, 40), 1, "none", {}),
         ((2, 513, 160), 1, "tensor", {}), ((2, 6, 45, 3), 2, "tensor", {}),
         ((2, 33, 17), 1, "tensor", dict(gain=0.7, bias=1.5, root=3.0, smooth=0.3)),
         ((2, 33, 17), 1, "none", dict(gain=1.3, root=0.5, smooth=1.5))]
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
plan = rt.get_plan(rt.FrontendConfig(), 0)
for shape, axis, scope, kw in CASES:
    worst = {"dx": (0, -1), "gain": (0, -1), "bias": (0, -1), "root": (0, -1), "smooth": (0, -1)}
    for seed in range(n_seeds if shape != (2, 513, 160) else max(4, n_seeds // 8)):
        rng = np.random.default_rng(seed)
        x = (rng.random(shape) ** 3 * 5.0 + 1e-3).astype(np.float32)
        g = rng.standard_normal(shape).astype(np.float32)
        p = rt.pcen_params(norm_scope=scope, **kw)
        dx, dp = plan.pcen_backward(torch.from_numpy(x).cuda(), torch.from_numpy(g).cuda(), p, axis)
        want_dx, want_dp = oracle.pcen_backward(x, g, scope=scope, axis=axis, **kw)
        dx, dp = dx.cpu().numpy(), dp.cpu().numpy()
        scale = np.abs(want_dx).max()
        r = float((np.abs(dx - want_dx) / (2e-4 * np.abs(want_dx) + 2e-5 * scale)).max())
        if r > worst["dx"][0]:
            worst["dx"] = (r, seed)
        for name, a, b in zip(("gain", "bias", "root", "smooth"), dp, want_dp):
            allow = 5e-4 * abs(b) + 1e-4 * np.abs(want_dp).max() + 2e-7 * np.abs(g).sum()
            r = abs(a - b) / allow
            if r > worst[name][0]:
                worst[name] = (float(r), seed)
    print(shape, scope, kw, {k: (round(v[0], 3), v[1]) for k, v in worst.items()}, flush=True)
