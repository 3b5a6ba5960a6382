#!/usr/bin/env python3
"""EMA alone (tfpcen.ExponentialMovingAverage) on [B, 513, 160]: ms per call and a checksum, for the in-tree library or --lib.
    python tools/probe_ema.py [--lib tools/variants/x.so] [--batch 4096]"""
import argparse, json, os, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    import torch
    from audio_training_b200 import _lib
    if a.lib:
        _lib.LIB_PATH = os.path.abspath(a.lib)
    from audio_training_b200 import _runtime as rt
    x = torch.rand((a.batch, 513, 160), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) ** 4 * 50.0
    plan = rt.Plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0)
    for _ in range(3):
        out = plan.ema(x, 0.04)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = plan.ema(x, 0.04)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"lib": os.path.relpath(_lib.LIB_PATH, REPO), "batch": a.batch, "ema_ms": e0.elapsed_time(e1) / a.steps,
                      "crc": int(out.view(torch.int32).to(torch.int64).sum())}), flush=True)


if __name__ == "__main__":
    main()
