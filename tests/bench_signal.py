#!/usr/bin/env python3
"""(Lives under tests/ because it times the oracle next to the device path.)  identifytracks.signal_noise on one recording: device path (FP32 STFT + cacfe_signal_components) against the numpy +
OpenCV restatement on the host cores.  CUDA-event timed after warm-up; prints one JSON line.
    python tests/bench_signal.py [seconds]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_training_b200 import _runtime as rt, identifytracks as it
from oracle import frontend_oracle as fo

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
frames = fo.synth_recording(seconds, seed=3)
x = torch.from_numpy(frames).cuda()
for _ in range(2):
    sig, spec = it.signal_noise(x, 48000)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 5
for _ in range(n):
    sig, spec = it.signal_noise(x, 48000)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / n * 1e3
# the array part alone, device resident
d = it._spectrogram(x, 2048, 281)
plan = rt.get_plan(rt.FrontendConfig(), 0)
for _ in range(2):
    plan.signal_components(d, 4, (6, 42), (3, 3))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    plan.signal_components(d, 4, (6, 42), (3, 3))
e1.record()
torch.cuda.synchronize()
comp_ms = e0.elapsed_time(e1) / n
t0 = time.perf_counter()
want, _ = fo.signal_noise(frames)
cpu_ms = (time.perf_counter() - t0) * 1e3
K, T = d.shape
print(json.dumps({"op": "identifytracks.signal_noise", "recording_s": seconds, "spectrogram": [K, T], "signals": len(sig),
                  "signals_cpu": len(want), "device_ms_whole_call": round(wall, 3), "device_ms_components_only": round(comp_ms, 3),
                  "cpu_ms_numpy_opencv": round(cpu_ms, 1), "speedup_whole_call": round(cpu_ms / wall, 1),
                  "components_GBps_of_spectrogram": round(K * T * 4 / comp_ms / 1e6, 1)}))
