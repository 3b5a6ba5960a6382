#!/usr/bin/env python3
"""A/B timing of builds of libcacfe.so that differ in compile-time switches of the fused kernel (K1).

    python tools/ab_k1.py                      # parent: runs every library under tools/variants/ and the in-tree one
    python tools/ab_k1.py --lib PATH [--batch] # child: K1's own CUDA-event time (plan.profile) for that library

Each library is timed in its own process (the loader binds one library per process), twice round-robin so that a clock or
thermal drift shows up as a difference between the two visits.  Prints one JSON line per visit: ms per launch of K1 at
4096 clips and a checksum of the features (variants must agree bit for bit unless they change arithmetic).
"""
import argparse, glob, json, os, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def child(lib, B, steps):
    import torch
    from audio_training_b200 import _lib
    _lib.LIB_PATH = lib
    from audio_training_b200 import _runtime as rt
    x = torch.rand((B, 144000), device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)) - 0.5
    plan = rt.Plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0)
    out = torch.empty((B, plan.n_frames, 160), dtype=torch.float32, device="cuda")
    for _ in range(3):
        plan.frontend(x, out)
    torch.cuda.synchronize()
    plan.profile(True)
    plan.profile_read()
    for _ in range(steps):
        plan.frontend(x, out)
    torch.cuda.synchronize()
    ms, n = plan.profile_read()
    plan.profile(False)
    print(json.dumps({"lib": os.path.relpath(lib, REPO), "k1_ms": ms / max(n, 1), "launches": n,
                      "checksum": float(out.double().sum()), "probe": float(out[B // 2, 100, 40])}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    if a.lib:
        return child(os.path.abspath(a.lib), a.batch, a.steps)
    libs = [os.path.join(REPO, "audio-training_b200", "libcacfe.so")] + sorted(glob.glob(os.path.join(REPO, "tools", "variants", "*.so")))
    for _ in range(2):
        for lib in libs:
            subprocess.run([sys.executable, os.path.abspath(__file__), "--lib", lib, "--batch", str(a.batch), "--steps", str(a.steps)], check=True)


if __name__ == "__main__":
    main()
