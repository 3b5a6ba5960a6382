#!/usr/bin/env python3
"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck): paths A (both layouts), B (zero and
reflect padding), C (banded and tensor core), PCEN scopes, compress modes, host pipe.  Checks results against the oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_training_b200 as atb
from audio_training_b200 import _runtime as rt
from oracle import frontend_oracle as fo

x = fo.synth_clips(np.arange(5))
xn = fo.normalize(x, np.float32)
w = fo.mel_f(48000, 160, 100, 11000, 4096, 1000)
dev = torch.from_numpy(x).cuda()
for layout, ch in (("btm", 1), ("bmtc", 3)):
    plan = rt.Plan(rt.FrontendConfig(normalize=True, channels=ch, out_layout=layout), 0, w)
    out = plan.frontend(dev).cpu().numpy()
    truth = fo.raw_to_mel(xn, w, channels=0, dtype=np.float64)
    got = np.swapaxes(out, 1, 2) if layout == "btm" else out[..., 0]
    ok, worst = fo.within_tolerance(got, truth)
    print("path A", layout, "ok" if ok else "FAIL", f"{worst:.3f}")
    assert ok
for mode in ("center_zero", "center_reflect"):
    plan = rt.Plan(rt.FrontendConfig(normalize=False, channels=1, framing=mode), 0, w)
    out = plan.frontend(torch.from_numpy(xn).cuda()).cpu().numpy()
    truth = np.stack([fo.get_spect(xn[i], pad_mode="constant" if mode == "center_zero" else "reflect") for i in range(len(xn))])
    ok, worst = fo.within_tolerance(out, truth)
    print("path B", mode, "ok" if ok else "FAIL", f"{worst:.3f}")
    assert ok
mag = np.stack([np.abs(fo.stft_librosa(xn[i], dtype=np.float32)).astype(np.float32) for i in range(2)])
for impl in ("banded_fp32", "tc_3xtf32"):
    plan = rt.Plan(rt.FrontendConfig(power=1, channels=1, mel_impl=impl), 0, w)
    out = plan.mel_from_spectrogram(torch.from_numpy(mag).cuda()).cpu().numpy()
    truth = np.stack([fo.mel_from_spectrogram(m, w) for m in mag])
    ok, worst = fo.within_tolerance(out, truth)
    print("path C", impl, "ok" if ok else "FAIL", f"{worst:.3f}")
    assert ok
plan = rt.Plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0, w)
for scope in ("tensor", "clip", "none"):
    out = plan.frontend_pcen(dev, rt.pcen_params(norm_scope=scope))
    assert torch.isfinite(out).all()
mel = plan.frontend(dev)
for mode in ("mag_pow", "power_to_db", "minmax", "std"):
    assert torch.isfinite(plan.compress(mel, mode, 0.27)).all()
pipe = rt.HostPipe(plan, max_B=8, chunk=2)
host = pipe.run(torch.from_numpy(x).pin_memory(), params=rt.pcen_params())
assert torch.equal(host, plan.frontend_pcen(dev).cpu())
torch.cuda.synchronize()
print("sanitize_small: all paths ran")
