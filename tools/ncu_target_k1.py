#!/usr/bin/env python3
"""Target of a K1-only ncu capture (one launch of the fused kernel at 1024 clips between profiler start / stop):

    python tools/ncu_target_k1.py && \
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r3_k1 \
        -k regex:stft_mel_v3 python tools/ncu_target_k1.py
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_training_b200 import _runtime as rt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.rand((B, 144000), device="cuda") - 0.5
plan = rt.Plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0)
out = torch.empty((B, plan.n_frames, 160), dtype=torch.float32, device="cuda")
plan.frontend(x, out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
plan.frontend(x, out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ncu target ran")
