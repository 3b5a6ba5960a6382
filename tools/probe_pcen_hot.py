#!/usr/bin/env python3
"""PCEN passes on the reference layer's shape [B, 513, 160]: compile-time instantiation (inner = 160, root 2) against the generic
kernel of the same library (plan.force_generic(2)) and, with --lib, against another build.  Prints ms per call and whether the
outputs agree bit for bit.   python tools/probe_pcen_hot.py [--lib tools/variants/prev.so] [--batch 4096]"""
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
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand((a.batch, 513, 160), device="cuda", generator=g) ** 4 * 50.0     # mel-power-like: non-negative, wide range
    x[:, 100:110, 5] = 0.0                                                            # a silent patch
    plan = rt.Plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0)
    res = {"lib": os.path.relpath(_lib.LIB_PATH, REPO), "batch": a.batch}
    outs = {}
    for mode in (0, 2):
        plan.force_generic(mode)
        for scope in ("tensor", "none"):
            p = rt.pcen_params(norm_scope=scope)
            for _ in range(3):
                out = plan.pcen(x, p)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                out = plan.pcen(x, p)
            e1.record()
            torch.cuda.synchronize()
            res[f"{'hot' if mode == 0 else 'generic'}_{scope}_ms"] = e0.elapsed_time(e1) / a.steps
            outs[(mode, scope)] = out.clone()
    for scope in ("tensor", "none"):
        res[f"bit_identical_{scope}"] = bool(torch.equal(outs[(0, scope)].view(torch.int32), outs[(2, scope)].view(torch.int32)))
        res[f"checksum_{scope}"] = float(outs[(0, scope)].double().sum())
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
