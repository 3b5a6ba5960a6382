#!/usr/bin/env python3
"""Residency sweep of the lane-per-row PCEN kernels: CACFE_PCEN_DYN_SMEM caps the resident blocks per SM (unused dynamic shared
memory), the library under test is chosen with --lib (builds with -DCACFE_PCEN_UNROLL=8 / 16).  One process per point."""
import json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

def child(lib, B):
    import torch
    from audio_training_b200 import _lib
    _lib.LIB_PATH = lib
    from audio_training_b200 import _runtime as rt
    plan = rt.Plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0)
    mel = torch.rand((B, 513, 160), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * 4 + 0.01
    res = {}
    for scope in ("none", "tensor"):
        prm = rt.pcen_params(norm_scope=scope)
        for _ in range(3):
            plan.pcen(mel, prm)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            out = plan.pcen(mel, prm)
        e1.record(); torch.cuda.synchronize()
        res[scope] = round(e0.elapsed_time(e1) / 20, 4)
    res["checksum"] = float(out.double().sum())
    print(json.dumps({"lib": os.path.basename(lib), "dyn": os.environ.get("CACFE_PCEN_DYN_SMEM", "0"), **res}), flush=True)

if len(sys.argv) > 2 and sys.argv[1] == "--lib":
    child(os.path.abspath(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 4096)
else:
    import glob
    libs = [os.path.join(REPO, "audio-training_b200", "libcacfe.so")] + sorted(glob.glob(os.path.join(REPO, "tools", "variants", "*.so")))
    sweep = (0, 10, 8, 7, 6, 5, 4) if "--residency" in sys.argv else (0,)
    for _ in range(2):
        for lib in libs:
            for R in sweep:
                env = dict(os.environ)
                env["CACFE_PCEN_DYN_SMEM"] = str(0 if R == 0 else (233472 // R - 1280) // 128 * 128)
                subprocess.run([sys.executable, os.path.abspath(__file__), "--lib", lib], env=env, check=False)
