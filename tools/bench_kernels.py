#!/usr/bin/env python3
"""Per-entry-point timing (CUDA events, inputs larger than L2) against the HBM roofline: one JSON line per op.
    python tools/bench_kernels.py [B]          (B clips for the raw-audio ops; default 2048)"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_training_b200 import _runtime as rt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
T, M, N = 513, 160, 144000


def timed(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def report(op, ms, bytes_, units, unit_name="clips"):
    print(json.dumps({"op": op, "ms": round(ms, 4), f"{unit_name}_per_s": round(units / ms * 1e3, 1),
                      "algorithmic_GBps": round(bytes_ / ms / 1e6, 1), "frac_of_hbm": round(bytes_ / ms / 1e6 / peak, 3)}))


x = torch.rand((B, N), device="cuda") - 0.5
plan = rt.Plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0)
report("normalize (cluster of 8 CTAs, clip in DSMEM)", timed(lambda: plan.normalize(x)), B * N * 4 * 2, B)
report("frontend raw -> mel [B,T,M] (K0 + K1)", timed(lambda: plan.frontend(x)), B * (N * 4 + T * M * 4), B)
img = rt.Plan(rt.FrontendConfig(normalize=True, channels=3, out_layout="bmtc"), 0)
report("frontend raw -> image [B,M,T,3] (raw_to_mel)", timed(lambda: img.frontend(x)), B * (N * 4 + 3 * T * M * 4), B)
mel = plan.frontend(x)
for scope in ("tensor", "clip", "none"):
    p = rt.pcen_params(norm_scope=scope)
    passes = 1 if scope == "none" else 1.5          # reduce pass reads, apply pass reads and writes
    report(f"pcen [B,T,M] scope={scope}", timed(lambda: plan.pcen(mel, p)), B * T * M * 4 * 2 * passes, B)
report("ema [B,T,M]", timed(lambda: plan.ema(mel, 0.04)), B * T * M * 4 * 2, B)
gout = torch.randn_like(mel)
for scope in ("tensor", "none"):   # algorithmic: read x, read dL/dout, write dL/dx
    p = rt.pcen_params(norm_scope=scope)
    report(f"pcen_backward [B,T,M] scope={scope}", timed(lambda: plan.pcen_backward(mel, gout, p)), B * T * M * 4 * 3, B)
del gout
for mode, reads in (("mag_pow", 1), ("power_to_db", 2), ("minmax", 2), ("std", 2)):
    report(f"compress {mode} (tensor-wide statistic)", timed(lambda: plan.compress(mel, mode, 0.27)), B * T * M * 4 * (reads + 1), B)
report("frontend + pcen (bench.py step)", timed(lambda: plan.frontend_pcen(x)), B * (N * 4 + T * M * 4), B)
del mel
nb = min(B, 384)
spec = torch.rand((nb, 2049, T), device="cuda")
for impl in ("banded_fp32", "tc_3xtf32"):
    pc = rt.Plan(rt.FrontendConfig(power=1, channels=1, mel_impl=impl), 0)
    report(f"mel_from_spectrogram {impl}", timed(lambda: pc.mel_from_spectrogram(spec)), nb * (930 * T * 4 + M * T * 4), nb)
del spec
st = rt.Plan(rt.FrontendConfig(framing="center_zero", power=1, channels=1, normalize=True), 0)
ns = min(B, 1024)
report("stft -> stored spectrogram [B,2049,T]", timed(lambda: st.stft(x[:ns]), 5), ns * (N * 4 + 2049 * T * 4), ns)
import numpy as np
sos = np.array([[0.02995458, 0.05990916, 0.02995458, 1.0, -1.45424359, 0.57406192]])
report("sosfilt (order-2 low-pass: 1 section, FP64 parallel scan)", timed(lambda: plan.sosfilt(sos, x), 5), B * N * 8, B)
from scipy.signal import butter
bp = butter(2, [800 / 24000, 5000 / 24000], btype="bandpass", output="sos")
report("sosfilt (order-2 band-pass: 2 sections)", timed(lambda: plan.sosfilt(bp, x), 5), B * N * 8, B)
