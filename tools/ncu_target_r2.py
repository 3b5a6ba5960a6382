#!/usr/bin/env python3
"""Target of the round-2 ncu captures (one launch of each kernel of interest between profiler start / stop):

    python tools/ncu_target_r2.py && \
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_kernels \
        -k regex:'stft_mel_v3|melspec_tc|melspec_stream|sosfilt_scan' python tools/ncu_target_r2.py
    python tools/ncu_summary_all.py gpurun_out/r2_kernels.ncu-rep > profiles/r02_kernels_ncu_summary.txt

K1 at 1024 clips (the size of every earlier capture), the tcgen05 and the streaming stored-spectrogram kernels at 384
clips, the parallel-scan sosfilt at 1024 clips (order-2 low-pass = 1 section, and a 2-section band-pass).
"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_training_b200 import _runtime as rt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.rand((B, 144000), device="cuda") - 0.5
plan = rt.Plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0)
out = torch.empty((B, plan.n_frames, 160), dtype=torch.float32, device="cuda")
spec = torch.rand((384, 2049, 513), device="cuda")
tc = rt.Plan(rt.FrontendConfig(power=1, channels=1, mel_impl="tc_3xtf32"), 0)
st = rt.Plan(rt.FrontendConfig(power=1, channels=1), 0)
from scipy.signal import butter
lp = butter(2, 3000 / 24000, btype="lowpass", output="sos")
bp = butter(2, [800 / 24000, 5000 / 24000], btype="bandpass", output="sos")


def everything():
    plan.frontend(x, out)
    tc.mel_from_spectrogram(spec)
    st.mel_from_spectrogram(spec)
    plan.sosfilt(lp, x)
    plan.sosfilt(bp, x)


everything()          # warm-up (tables, workspaces)
torch.cuda.synchronize()
torch.cuda.profiler.start()
everything()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ncu target ran")
