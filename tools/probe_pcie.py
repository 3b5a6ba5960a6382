#!/usr/bin/env python3
"""Host link ceiling for the end-to-end number: pinned-memory H2D and D2H rates of this box, each alone and both at once.

bench.py's `e2e` moves 576 000 B per clip to the device and 328 320 B back (f32 clip in, f32 features out: the
reference's contract).  This probe measures what the link gives for exactly those two buffer sizes (one bench batch:
4096 clips), so that "e2e is PCIe-bound" is a measured statement: e2e clips/s <= min(H2D rate / 576 000,
D2H rate / 328 320) with both directions busy.

    python tools/probe_pcie.py [--clips 4096] [--reps 5]

Prints one JSON line.  No kernels of this repository are involved (plumbing only: torch pinned tensors and streams).
"""
import argparse
import json
import time

import torch

CLIP_IN = 144000 * 4
CLIP_OUT = 513 * 160 * 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n_in, n_out = args.clips * CLIP_IN // 4, args.clips * CLIP_OUT // 4
    h_in = torch.empty(n_in, dtype=torch.float32, pin_memory=True).fill_(1.0)
    h_out = torch.empty(n_out, dtype=torch.float32, pin_memory=True)
    d_in = torch.empty(n_in, dtype=torch.float32, device=dev)
    d_out = torch.ones(n_out, dtype=torch.float32, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def up():
        with torch.cuda.stream(s_up):
            d_in.copy_(h_in, non_blocking=True)

    def down():
        with torch.cuda.stream(s_dn):
            h_out.copy_(d_out, non_blocking=True)

    def timed(fns):
        for f in fns:
            f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            for f in fns:
                f()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / args.reps

    t_up, t_dn, t_both = timed([up]), timed([down]), timed([up, down])
    gb_in, gb_out = n_in * 4 / 1e9, n_out * 4 / 1e9
    res = {
        "clips": args.clips,
        "h2d_alone_GBps": gb_in / t_up,
        "d2h_alone_GBps": gb_out / t_dn,
        "both_ms": t_both * 1e3,
        "both_h2d_GBps_if_link_bound": gb_in / t_both,
        "both_d2h_GBps_if_link_bound": gb_out / t_both,
        "e2e_ceiling_clips_per_s": args.clips / t_both,
        "e2e_ceiling_h2d_only_clips_per_s": args.clips / t_up,
        "note": "ceiling = one batch's upload and download issued together on two streams; the longer of the two sets the time",
    }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
