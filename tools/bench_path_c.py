#!/usr/bin/env python3
"""Path C (stored magnitude spectrogram [B, 2049, 513] -> mel [B, 160, 513, 1]): banded FP32 kernel against the tcgen05
banded 3xTF32 GEMM, CUDA-event timed, against the HBM roofline (algorithmic bytes: bins 9..938 in, 160 x 513 out)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_training_b200 import _runtime as rt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 384
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
spec = torch.rand((B, 2049, 513), device="cuda") * 10
alg = B * (930 * 513 * 4 + 160 * 513 * 4)
full = B * (2049 * 513 * 4 + 160 * 513 * 4)
res = {}
for impl in ("banded_fp32", "tc_3xtf32"):
    plan = rt.Plan(rt.FrontendConfig(power=1, channels=1, mel_impl=impl), 0)
    out = plan.mel_from_spectrogram(spec)
    for _ in range(3):
        plan.mel_from_spectrogram(spec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    n = 10
    for _ in range(n):
        plan.mel_from_spectrogram(spec)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[impl] = out
    print(json.dumps({"impl": impl, "B": B, "ms": ms, "clips_per_s": B / ms * 1e3, "algorithmic_GBps": alg / ms / 1e6,
                      "frac_of_hbm": alg / ms / 1e6 / peak, "if_whole_array_were_read_GBps": full / ms / 1e6}))
d = (res["tc_3xtf32"] - res["banded_fp32"]).abs().max().item()
print("max |tc - banded| =", d, " max |banded| =", res["banded_fp32"].abs().max().item())
