#!/usr/bin/env python3
"""K1 on clips the caller has normalised (plan with normalize=False: raw_to_mel / get_spect): CUDA-event time per 4096 clips.
    python tools/probe_nonorm.py [LIB]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_training_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
from audio_training_b200 import _runtime as rt
B = 4096
x = torch.rand((B, 144000), device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)) - 0.5
for norm in (False, True):
    plan = rt.Plan(rt.FrontendConfig(normalize=norm, channels=1, out_layout="btm"), 0)
    out = torch.empty((B, plan.n_frames, 160), dtype=torch.float32, device="cuda")
    for _ in range(3):
        plan.frontend(x, out)
    torch.cuda.synchronize()
    plan.profile(True); plan.profile_read()
    for _ in range(10):
        plan.frontend(x, out)
    torch.cuda.synchronize()
    ms, n = plan.profile_read()
    print(json.dumps({"lib": os.path.basename(_lib.LIB_PATH), "normalize": norm, "k1_ms": ms / n, "checksum": float(out.double().sum())}))
