#!/usr/bin/env python3
"""cacfe_stft (stored-spectrogram producer): clips/s and written bandwidth, CUDA-event timed."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_training_b200 import _runtime as rt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.rand((B, 144000), device="cuda") - 0.5
plan = rt.Plan(rt.FrontendConfig(framing="center_zero", power=1, channels=1, normalize=True), 0)
for _ in range(3):
    out = plan.stft(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
n = 5
for _ in range(n):
    plan.stft(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"op": "cacfe_stft", "B": B, "ms": ms, "clips_per_s": B / ms * 1e3,
                  "GBps_written": B * 2049 * 513 * 4 / ms / 1e6, "GBps_algorithmic": B * (144000 * 4 + 2049 * 513 * 4) / ms / 1e6}))
