#!/usr/bin/env python3
"""Headline benchmark: clips/s for 3 s @ 48 kHz clips -> normalise -> STFT -> mel -> PCEN (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A step = one pass of the hot path over one batch of B synthetic clips (BASELINE.json configs[1]: B = 4096 on one
B200).  With N > 1 (torchrun, one rank per GPU) every rank processes its own B clips -- clips are independent, no
collective on the data path -- and `value` is all ranks' clips over the slowest rank's time ("weak" scaling).

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel (fused STFT/power/mel), algorithmic bytes per launch / its CUDA-event duration
                measured inside the timed region, against MEASURED_PEAKS.json hbm_gbs; `fp32` gives the roof
                that actually binds it (the FFT is FP32-issue bound, not HBM bound)
  cpu_baseline  the oracle's port of the reference's own CPU op order (f32, scipy pocketfft, all host threads) on
                a bounded sample of the same workload
  e2e           same metric through the host-buffer API (cacfe_hostpipe: pinned host in -> H2D -> kernels -> D2H)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The reference arm uses every host thread it can (torchrun exports OMP_NUM_THREADS=1, which would cripple the CPU path):
# the thread-pool sizes are read when numpy / scipy load, so they are set before the import.
if "reference" in sys.argv:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

CLIP = 144000
T_FRAMES, N_MELS = 513, 160
BYTES_PER_CLIP = CLIP * 4 + T_FRAMES * N_MELS * 4          # 904 320 B  (path A fused, 1 channel / BTM)
FLOPS_PER_CLIP = 68e6                                       # FFT 63.0 M + window/power 5.25 M + banded mel
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12           # 74.4
METRIC = "clips/sec (3 s @48 kHz -> mel+PCEN)"
WORKLOAD = ("fused normalize -> STFT 4096/281 pad_end -> |z|^2 -> mel 160 -> PCEN (tensor-global min-max), "
            "batch {B} clips x 3 s @ 48 kHz per GPU")


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons while the timed region runs (B200_PROFILING.md recipe), sampled through NVML every
    20 ms from a host thread (nvidia-smi -lms needs ~0.2 s to deliver its first line, longer than a short bench)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.samples, self.first, self.run, self.thread, self.nvml = index, [], 0, False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def start(self):
        if self.nvml is None:
            return
        self.run = True
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        n = self.nvml
        while self.run:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, mask))
            except Exception:
                pass
            time.sleep(0.02)

    def mark(self):
        """Samples before this call (warm-up) are dropped."""
        self.first = len(self.samples)

    def stop(self):
        if self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"]}
        self.run = False
        self.thread.join(timeout=2)
        rows = self.samples[self.first:] or self.samples[-1:]
        sm = [r[0] for r in rows]
        reasons = sorted(name for name, bit in self.REASONS.items() if any(r[1] & bit for r in rows))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                "reasons": reasons}


def numa_node_of(addr):
    """NUMA node holding the page at `addr` (move_pages(2) with nodes = NULL only queries); -1 when unknown."""
    try:
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        page = ctypes.c_void_p(addr & ~0xFFF)
        status = ctypes.c_int(-1)
        rc = libc.syscall(279, 0, ctypes.c_ulong(1), ctypes.byref(page), None, ctypes.byref(status), 0)   # __NR_move_pages, x86-64
        return int(status.value) if rc == 0 else -1
    except Exception:
        return -1


def gpu_numa_node(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(ClockSampler._physical_index(index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        with open(f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node") as fh:
            return int(fh.read().strip())
    except Exception:
        return -1


def synth_on_device(torch, B, device, seed):
    """SURVEY 8d generator shape (noise + 3 linear chirps + DC) evaluated on the device, in chunks."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((B, CLIP), dtype=torch.float32, device=device)
    t = torch.arange(CLIP, device=device, dtype=torch.float64) / 48000.0
    for b0 in range(0, B, 256):
        nb = min(256, B - b0)
        x = 0.3 * (2.0 * torch.rand((nb, CLIP), generator=g, device=device) - 1.0)
        p = torch.rand((nb, 3, 3), generator=g, device=device, dtype=torch.float64)
        for j in range(3):
            a = 0.05 + 0.45 * p[:, j, 0:1]
            f0 = 300.0 + 10700.0 * p[:, j, 1:2]
            f1 = 300.0 + 10700.0 * p[:, j, 2:3]
            c = (f1 - f0) / 3.0
            phase = 2.0 * np.pi * (f0 * t[None] + 0.5 * c * t[None] ** 2)
            x += (a * torch.sin(phase)).float()
        x += (0.2 * torch.rand((nb, 1), generator=g, device=device) - 0.1)
        out[b0:b0 + nb] = x
    return out


def cpu_reference_run(n_clips, budget_s, min_batches=3, steps=None):
    """The reference's CPU op order (oracle ports, f32) on batches of `n_clips` synthetic clips, all host threads.
    Two faithful ports are timed -- torch's CPU kernels (FFT + batched sgemm, intra-op threads) and numpy / scipy pocketfft --
    and the FASTER one is the baseline (round-1 review: the numpy form alone flatters the GPU ratio 2-4x).
    -> {"value", "kind_detail", "batches", "torch_port", "numpy_port", "threads"}"""
    import torch
    from oracle import frontend_oracle as fo
    torch.set_num_threads(os.cpu_count() or 1)
    w = fo.mel_f(48000, 160, 100, 11000, 4096, 1000)
    x = fo.synth_clips(np.arange(n_clips))
    res = {}
    for name, fn in (("torch_port", fo.reference_cpu_path_torch), ("numpy_port", fo.reference_cpu_path)):
        fn(x, w)  # warm-up
        times = []
        t_end = time.perf_counter() + budget_s / 2
        while (len(times) < (steps or min_batches)) if steps else (len(times) < min_batches or (time.perf_counter() < t_end and len(times) < 50)):
            t0 = time.perf_counter()
            fn(x, w)
            times.append(time.perf_counter() - t0)
        res[name] = n_clips / float(np.median(times))
        res[name + "_batches"] = len(times)
    best = "torch_port" if res["torch_port"] >= res["numpy_port"] else "numpy_port"
    res.update(value=res[best], best=best, batches=res[best + "_batches"], threads=torch.get_num_threads())
    return res


CPU_SAMPLE = ("{n} batches of 32 synthetic clips (BASELINE.json configs[0]) per port; the reference's f32 op order "
              "(normalize -> stft 4096/281 pad_end -> z**2, abs -> replicated-weight batch matmul -> x3 channels -> PCEN) as two "
              "faithful CPU ports, value = the faster one ({best}): torch CPU kernels {tp:.1f} clips/s, numpy + scipy pocketfft "
              "{npp:.1f} clips/s, {th} threads; TensorFlow / librosa are not installable offline")


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n = 32
    steps = max(1, args.steps)
    r = cpu_reference_run(n, budget_s=0.0, steps=min(steps, 20))
    value = r["value"]
    dt = n / value
    sample = CPU_SAMPLE.format(n=r["batches"], best=r["best"], tp=r["torch_port"], npp=r["numpy_port"], th=r["threads"])
    line = {"metric": METRIC, "value": value, "unit": "clips/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(B=4096),
                       "sample": "each step is one batch of 32 clips of that workload on the host cores (BASELINE.json configs[0])"},
            "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample,
                             "torch_port": r["torch_port"], "numpy_port": r["numpy_port"]},
            "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    args.emit(line)
    return 0


def run_ours(args):
    import torch
    import audio_training_b200 as atb
    from audio_training_b200 import _runtime as rt
    from audio_training_b200 import distributed as dist_

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    rank, world, local = dist_.init()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    B = args.batch
    peaks, peak_kind = measured_peaks()

    cfg = rt.FrontendConfig(normalize=True, channels=1, out_layout="btm")
    plan = rt.get_plan(cfg, local)
    params = rt.pcen_params()
    x = synth_on_device(torch, B, device, 20240 + rank)
    out = torch.empty((B, plan.n_frames, cfg.n_mels), dtype=torch.float32, device=device)
    plan.workspace_for(B)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~0.1 s to deliver its first line: start it before the warm-up
    for _ in range(max(args.warmup, 3)):
        plan.frontend_pcen(x, params, out)
    barrier()
    if rank == 0:
        sampler.mark()           # only samples from here on count
    plan.profile(True)
    plan.profile_read()
    launches0 = plan.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        plan.frontend_pcen(x, params, out)   # inputs (2.36 GB at B=4096) far exceed the 126 MB L2: no flush needed
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = plan.launch_count() - launches0
    k1_ms, k1_n = plan.profile_read()
    plan.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = dist_.max_over_ranks(ms_total / args.steps, device)
    value = world * B / (ms_step * 1e-3)

    # ---- every rank's result, checked: an exact checksum of all output bits (sum of the int32 views, order independent) per
    # rank; rank 0 then regenerates each rank's batch from its seed, runs it on its own GPU and compares.  Catches a rank
    # that computed something else (wrong device, stale buffer, a kernel that misbehaves on one GPU only).
    def bits_checksum(t):
        return int(t.view(torch.int32).sum(dtype=torch.int64).item())

    my_sum = bits_checksum(out)
    rank_sums = [my_sum]
    if world > 1:
        gathered = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
        torch.distributed.all_gather(gathered, torch.tensor([my_sum], dtype=torch.int64, device=device))
        rank_sums = [int(g.item()) for g in gathered]
    rank_check = None
    if rank == 0:
        recomputed = []
        for r in range(world):
            xr = x if r == 0 else synth_on_device(torch, B, device, 20240 + r)
            recomputed.append(bits_checksum(plan.frontend_pcen(xr, params, out)))
            del xr
        rank_check = {"per_rank_checksum": rank_sums, "recomputed_on_rank0": recomputed, "match": recomputed == rank_sums}
        if recomputed != rank_sums:
            raise SystemExit(f"bench.py: per-rank results differ from their single-GPU recomputation: {rank_check}")
        plan.frontend_pcen(x, params, out)      # leave rank 0's own result in `out`

    # ---- DRAM traffic of the dominant kernel: from the committed ncu --set full capture (1024 clips per launch),
    # scaled to this launch's clip count; None when the summary is not there
    traffic = None
    try:
        import re
        # round-2 capture (tools/ncu_target_r2.py): the first kernel block of the digest is stft_mel_v3_kernel at 1024 clips
        txt = open(os.path.join(REPO, "profiles", "r02_kernels_ncu_summary.txt")).read()
        txt = txt[txt.index("stft_mel_v3_kernel"):]
        txt = txt[:txt.index("\n== ")]
        rd = float(re.search(r"dram__bytes_read\.sum \('([0-9.]+)', 'Mbyte'\)", txt).group(1))
        wr = float(re.search(r"dram__bytes_write\.sum \('([0-9.]+)', 'Mbyte'\)", txt).group(1))
        traffic = (rd + wr) * 1e6 / 1024.0 * B
    except Exception:
        traffic = None

    # ---- roofline of the dominant kernel, from the in-region CUDA events --------------------------------------
    k1_avg_ms = k1_ms / max(k1_n, 1)
    achieved = BYTES_PER_CLIP * B / (k1_avg_ms * 1e-3) / 1e9
    fp32_achieved = FLOPS_PER_CLIP * B / (k1_avg_ms * 1e-3) / 1e12
    # `bound` names the roof that binds this kernel.  The contract's vocabulary is "hbm" | "tensor"; neither is true here: the
    # fused kernel performs 75 flop per algorithmic byte on the CUDA cores (ridge 11.5), so the top-level achieved / peak /
    # frac are FP32 TFLOP/s against the nominal 148 SM x 128 lanes x 2 x 1.965 GHz, and `hbm` carries the fraction of the
    # measured HBM peak that BASELINE.json's metric asks for beside it.
    roofline = {"kernel": "stft_mel_v3_kernel", "bound": "fp32", "achieved": fp32_achieved, "peak": FP32_PEAK_TFLOPS,
                "unit": "TFLOP/s", "frac": fp32_achieved / FP32_PEAK_TFLOPS,
                "note": "68 MFLOP/clip (SURVEY 8d: 2.5 N log2 N FFT convention + window, power, banded mel) over the kernel's "
                        "CUDA-event time; peak is nominal FP32 FMA issue (no measured FP32 figure in MEASURED_PEAKS.json)",
                "traffic": traffic,
                "traffic_note": "bytes per launch: dram__bytes_read.sum + dram__bytes_write.sum of the ncu --set full capture "
                                "profiles/r02_kernels_ncu_summary.txt (1024 clips per launch) scaled to this batch; a profiler "
                                "counter cannot be read in an un-profiled run", "peak_kind": peak_kind,
                "ms_per_launch": k1_avg_ms, "share_of_step": k1_avg_ms / (ms_total / args.steps),
                "algorithmic_bytes_per_launch": BYTES_PER_CLIP * B, "algorithmic_flops_per_launch": FLOPS_PER_CLIP * B,
                "hbm": {"achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                        "note": "904 320 algorithmic bytes per clip over the same time: what BASELINE.json's '% HBM roofline' "
                                "asks; the FFT's arithmetic, not memory, bounds this kernel (SURVEY 8d: <= 15 % reachable)"},
                # kept under its round-1 name for readers of older lines
                "fp32": {"achieved": fp32_achieved, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "frac": fp32_achieved / FP32_PEAK_TFLOPS}}

    # ---- e2e: host buffers through the public host API, copies inside the timed region --------------------------
    e2e = None
    if not args.no_e2e:
        # Two or three HostPipes, one host thread each: a step of one pipe is H2D -> kernels -> D2H of ITS batch; with several
        # in flight the D2H of one batch overlaps the H2D of the next (PCIe is full duplex), which is how a data loader
        # would call it.  Every step's copies are inside the timed region.
        chunk = min(args.chunk, B)
        n_pipes = args.pipes if args.pipes > 0 else (3 if world <= 2 else 2)
        pipes = [rt.HostPipe(plan, max_B=B, chunk=chunk) for _ in range(n_pipes)]
        h_in = [torch.empty((B, CLIP), dtype=torch.float32, pin_memory=True) for _ in range(n_pipes)]
        h_out = [torch.empty((B, plan.n_frames, cfg.n_mels), dtype=torch.float32, pin_memory=True) for _ in range(n_pipes)]
        for q in range(n_pipes):
            h_in[q].copy_(x)
            pipes[q].run(h_in[q], h_out[q], params)
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        errors = []

        def timed_e2e(bufs):
            def drive(q):
                try:
                    for _ in range(n_e2e):
                        pipes[q].run(bufs[q], h_out[q], params)
                except Exception as exc:  # surfaced below: a failed pipe must fail the bench
                    errors.append(exc)

            threads = [threading.Thread(target=drive, args=(q,)) for q in range(n_pipes)]
            t0 = time.perf_counter()
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            dt_ = (time.perf_counter() - t0) / (n_e2e * n_pipes)
            if errors:
                raise errors[0]
            return dist_.max_over_ranks(dt_, device)

        dt = timed_e2e(h_in)
        e2e = {"value": world * B / dt, "unit": "clips/s", "h2d_bytes_per_step": B * CLIP * 4,
               "d2h_bytes_per_step": B * plan.n_frames * cfg.n_mels * 4, "steps": n_e2e * n_pipes,
               "api": "HostPipe.run (cacfe_hostpipe_run): pinned host -> H2D -> normalise/STFT/mel/PCEN -> D2H; "
                      f"{n_pipes} pipes on {n_pipes} host threads, each step copies its own batch in and out",
               "checksum": float(h_out[0][0, :4, :4].sum())}
        # What the host link gives for exactly these two buffers (one batch up, one batch of features down, issued together
        # on two streams; plumbing only, no kernels), measured at EVERY N with all ranks copying at the same time: the
        # ceiling of any f32-in / f32-out end-to-end path on this box at this N.  Per-rank rates and the NUMA node of each
        # rank's pinned buffer and GPU go into the line so that the e2e scaling curve is attributed by measurement.
        d_up = torch.empty((B, CLIP), dtype=torch.float32, device=device)
        d_dn = torch.ones((B, plan.n_frames, cfg.n_mels), dtype=torch.float32, device=device)
        s_up, s_dn = torch.cuda.Stream(device), torch.cuda.Stream(device)

        def both():
            with torch.cuda.stream(s_up):
                d_up.copy_(h_in[0], non_blocking=True)
            with torch.cuda.stream(s_dn):
                h_out[0].copy_(d_dn, non_blocking=True)

        both()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            both()
        torch.cuda.synchronize()
        t_mine = (time.perf_counter() - t0) / 3
        t_link = dist_.max_over_ranks(t_mine, device)
        mine = torch.tensor([B * CLIP * 4 / t_mine / 1e9, B * plan.n_frames * cfg.n_mels * 4 / t_mine / 1e9,
                             float(numa_node_of(h_in[0].data_ptr())), float(gpu_numa_node(local))], dtype=torch.float64, device=device)
        per_rank = [mine]
        if world > 1:
            per_rank = [torch.zeros_like(mine) for _ in range(world)]
            torch.distributed.all_gather(per_rank, mine)
        ceiling = world * B / t_link
        e2e["link_ceiling"] = {"clips_per_s": ceiling, "frac": e2e["value"] / ceiling, "n_gpus": world,
                               "h2d_GBps": world * B * CLIP * 4 / t_link / 1e9,
                               "per_rank": [{"h2d_GBps": round(float(v[0]), 2), "d2h_GBps": round(float(v[1]), 2),
                                             "pinned_numa_node": int(v[2]), "gpu_numa_node": int(v[3])} for v in per_rank],
                               "host_threads": os.cpu_count(),
                               "note": "every rank's H2D and D2H of one batch issued together from pinned memory, all ranks "
                                       "at once, no kernels: the host side bounds e2e, not the GPU"}
        del d_up, d_dn
        # Extension, reported beside the headline and never instead of it: the same run from 16-bit PCM host samples
        # (cacfe_hostpipe_run_pcm16: half the upload bytes, s / 32768 on the device, features bit-identical to the float32
        # call on the converted samples).  The reference's callables take float32, so `e2e.value` above stays the number.
        del h_in                      # the float32 staging goes first: the pinned footprint per rank does not grow
        try:
            if world != 1:            # single-GPU line only: a leg that may fail must not sit between collectives
                raise RuntimeError("measured at N = 1 only")
            h_pcm = [torch.empty((B, CLIP), dtype=torch.int16, pin_memory=True) for _ in range(n_pipes)]
            for q in range(n_pipes):
                for b0 in range(0, B, 512):
                    h_pcm[q][b0:b0 + 512].copy_((x[b0:b0 + 512] * 16000.0).round().clamp_(-32768, 32767).to(torch.int16))
                pipes[q].run(h_pcm[q], h_out[q], params)
            barrier()
            dt16 = timed_e2e(h_pcm)
            e2e["pcm16_input"] = {"value": world * B / dt16, "unit": "clips/s", "h2d_bytes_per_step": B * CLIP * 2,
                                  "d2h_bytes_per_step": B * plan.n_frames * cfg.n_mels * 4,
                                  "note": "extension (16-bit PCM host samples, converted on the device as soundfile / "
                                          "librosa.load convert a 16-bit file); not the reference's float32 contract, not "
                                          "the headline"}
            del h_pcm
        except Exception as exc:      # an extension leg must never cost the bench line
            e2e["pcm16_input"] = {"skipped": str(exc)}
        del pipes
        del h_out

    # ---- the other HBM-bound rows of the path, timed alone on rank 0: not part of `value` ----------------
    rows = None
    if rank == 0 and world == 1 and not args.no_rows:
        def alone(fn, n=10):
            for _ in range(2):
                fn()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a0.record()
            for _ in range(n):
                fn()
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / n

        def row(ms, clips, alg_bytes_per_clip, moved_bytes_per_clip):
            return {"ms": ms, "clips": clips, "clips_per_s": clips / ms * 1e3,
                    "algorithmic_GBps": clips * alg_bytes_per_clip / ms / 1e6,
                    "frac_of_hbm": clips * alg_bytes_per_clip / ms / 1e6 / peaks["hbm_gbs"],
                    "moved_GBps": clips * moved_bytes_per_clip / ms / 1e6}

        T, M = plan.n_frames, cfg.n_mels
        feat = T * M * 4
        mel = plan.frontend(x)
        rows = {"note": "standalone entry points, CUDA events, inputs resident and larger than L2; algorithmic bytes per clip as in "
                        "SURVEY 8d, moved = what the implementation has to touch (a tensor-global statistic costs one more read)"}
        rows["pcen_tensor_scope"] = row(alone(lambda: plan.pcen(mel, params)), B, 2 * feat, 3 * feat)
        rows["pcen_no_minmax"] = row(alone(lambda: plan.pcen(mel, rt.pcen_params(norm_scope="none"))), B, 2 * feat, 2 * feat)
        rows["ema"] = row(alone(lambda: plan.ema(mel, 0.04)), B, 2 * feat, 2 * feat)
        rows["normalize"] = row(alone(lambda: plan.normalize(x)), B, 2 * CLIP * 4, 2 * CLIP * 4)   # cluster kernel: one read, one write
        rows["minmax_epilogue"] = row(alone(lambda: plan.compress(mel, "minmax")), B, 2 * feat, 3 * feat)
        del mel
        nb = min(B, 384)
        pc = rt.get_plan(rt.FrontendConfig(power=1, channels=1), local)
        spec = torch.rand((nb, pc.n_bins, T), device=device)
        lo, hi = pc.bin_range()
        alg = (hi - lo + 1) * T * 4 + M * T * 4
        rows["mel_from_spectrogram"] = row(alone(lambda: pc.mel_from_spectrogram(spec)), nb, alg, alg)
        del spec

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference_run(32, budget_s=20.0)
        cpu = {"value": r["value"], "unit": "clips/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": CPU_SAMPLE.format(n=r["batches"], best=r["best"], tp=r["torch_port"], npp=r["numpy_port"], th=r["threads"]),
               "torch_port": r["torch_port"], "numpy_port": r["numpy_port"]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD.format(B=B),
                           "batch_per_gpu": B, "l2": "inputs (%.2f GB per step) exceed the 126 MB L2" % (B * CLIP * 4 / 1e9),
                           "parallelism": f"clips sharded over {world} GPU(s), no data-path collective"},
                "roofline": roofline, "other_rows": rows, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "checksum": float(out[0, :4, :4].sum()), "rank_check": rank_check}
        args.emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def synth_keyed(torch, indices, device, seed=20240):
    """SURVEY 8d generator with a counter-based RNG keyed (seed, global clip index): clip i is the same whichever rank and
    whichever chunk it lands in.  indices: int64 tensor [n] on `device` -> float32 [n, CLIP]."""
    def u01(ctr):                                   # splitmix64-style integer hash -> uniform [0, 1) (int64 wrap-around is intended)
        z = ctr * -7046029254386353131              # 0x9E3779B97F4A7C15 as int64
        z = (z ^ (z >> 30 & 0x3FFFFFFFF)) * -4658895280553007687
        z = (z ^ (z >> 27 & 0x1FFFFFFFFF)) * -7723592293110705685
        z = z ^ (z >> 31 & 0x1FFFFFFFF)
        return ((z >> 11) & 0x1FFFFFFFFFFFFF).to(torch.float64) * (1.0 / 9007199254740992.0)

    if indices.numel() > 512:                       # the int64 counters are 8 bytes per sample: generate 512 clips at a time
        return torch.cat([synth_keyed(torch, indices[i:i + 512], device, seed) for i in range(0, indices.numel(), 512)])
    n = indices.numel()
    key = (indices.to(torch.int64) + seed * 1000003).view(n, 1)
    col = torch.arange(CLIP, device=device, dtype=torch.int64).view(1, CLIP)
    x = (0.3 * (2.0 * u01(key * 262144 * 4 + col) - 1.0)).float()
    t = torch.arange(CLIP, device=device, dtype=torch.float64) / 48000.0
    p = u01(key * 262144 * 4 + 1000000 + torch.arange(10, device=device, dtype=torch.int64).view(1, 10))      # 10 parameters per clip
    for j in range(3):
        a = 0.05 + 0.45 * p[:, 3 * j:3 * j + 1]
        f0 = 300.0 + 10700.0 * p[:, 3 * j + 1:3 * j + 2]
        f1 = 300.0 + 10700.0 * p[:, 3 * j + 2:3 * j + 3]
        x += (a * torch.sin(2.0 * np.pi * (f0 * t[None] + 0.5 * ((f1 - f0) / 3.0) * t[None] ** 2))).float()
    x += (0.2 * p[:, 9:10] - 0.1).float()
    return x.contiguous()


def run_corpus(args):
    """BASELINE.json configs[2] as SURVEY 8d writes it: a corpus of `--corpus` distinct synthetic clips, clip i -> rank i mod R,
    processed in chunks of `--batch` clips generated on the device from the keyed generator (generation outside the timed
    region: CUDA events bracket only the front-end steps), per-clip PCEN scope so that a clip's features do not depend on how the
    corpus was sharded.  Every rank folds an exact per-clip checksum (sum of the feature bits) into an order-independent corpus
    checksum; rank 0 all-reduces it, and prints it: the same corpus gives the same number at every N."""
    import torch
    from audio_training_b200 import _runtime as rt
    from audio_training_b200 import distributed as dist_
    rank, world, local = dist_.init()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    B = min(args.batch, 4096)
    mine = torch.arange(rank, args.corpus, world, dtype=torch.int64, device=device)      # clip i -> rank i mod R
    cfg = rt.FrontendConfig(normalize=True, channels=1, out_layout="btm")
    plan = rt.get_plan(cfg, local)
    params = rt.pcen_params(norm_scope="clip")
    out = torch.empty((B, plan.n_frames, cfg.n_mels), dtype=torch.float32, device=device)
    plan.frontend_pcen(synth_keyed(torch, mine[:min(B, mine.numel())], device), params, out[:min(B, mine.numel())])   # warm-up
    total_ms, clips, launches0 = 0.0, 0, plan.launch_count()
    csum = torch.zeros(1, dtype=torch.int64, device=device)
    if world > 1:
        torch.distributed.barrier()
    for c0 in range(0, mine.numel(), B):
        idx = mine[c0:c0 + B]
        x = synth_keyed(torch, idx, device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        plan.frontend_pcen(x, params, out[:idx.numel()])
        e1.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        clips += idx.numel()
        per_clip = out[:idx.numel()].view(torch.int32).view(idx.numel(), -1).sum(dim=1, dtype=torch.int64)
        csum += (per_clip * (2 * idx + 1)).sum()          # weighted by the global index: a swapped pair of clips shows
        del x
    launches = plan.launch_count() - launches0
    ms = dist_.max_over_ranks(total_ms, device)
    if world > 1:
        torch.distributed.all_reduce(csum)
    if rank == 0:
        args.emit({"metric": METRIC + ", 1 M-clip corpus sharded over the GPUs", "value": args.corpus / (ms * 1e-3), "unit": "clips/s",
                   "n_gpus": world, "steps": -(-mine.numel() // B), "warmup": 1, "ms_per_step": ms / max(1, -(-mine.numel() // B)),
                   "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                   "config": {"workload": f"corpus of {args.corpus} distinct keyed synthetic clips, clip i -> rank i mod {world}, chunks of "
                                          f"{B} generated on the device (untimed), fused normalize -> STFT -> mel -> PCEN (per-clip min-max)",
                              "parallelism": f"{world} GPU(s), no data-path collective"},
                   "corpus_checksum": int(csum.item()), "gpu_launches": int(launches)})
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def run_config45(args):
    """BASELINE.json configs[3] / configs[4]: the front-end feeding its consumers on the device
    -- `--config 4`: centred front-end -> badwinner2 inference (predict.py path); `--config 5`: raw_to_mel (C = 3) -> PCEN ->
    wr_resnet_bird training step, one process per GPU (DDP over NCCL under torchrun).  The consumers are the torch
    restatements of audio-training_b200/consumers.py (cuDNN, bf16 autocast; pinned against the executed reference graph by
    tests/test_consumers.py), random weights.  `value` = end-to-end clips/s; the front-end's own rate and its share of the
    step say whether it keeps the model fed.  Not the headline: configs[1] (no --config) is."""
    import types

    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(REPO, "tools"))
    import bench_configs as bc
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --config 4|5: no CUDA device")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))
    a = types.SimpleNamespace(batch=args.batch if args.batch != 4096 else 64, steps=max(2, min(args.steps, 10)), labels=10)
    res = (bc.config4 if args.config == 4 else bc.config5)(a, dev, world)
    if int(os.environ.get("RANK", 0)) == 0:
        ms = a.batch * world * 1e3 / res["end_to_end_clips_per_s"]
        args.emit({"metric": METRIC + f" feeding the consumer of BASELINE config {args.config}", "value": res["end_to_end_clips_per_s"],
                   "unit": "clips/s", "n_gpus": world, "steps": a.steps, "warmup": 2, "ms_per_step": ms, "higher_is_better": True,
                   "scaling": "weak", "vs_baseline": None, "dtype": "f32 features, bf16 consumer", "data": "synthetic",
                   "config": {"workload": res["config"], "batch_per_gpu": a.batch, "model": res["model"]},
                   "frontend_clips_per_s": res["frontend_clips_per_s"], "model_clips_per_s": res["model_clips_per_s"],
                   "frontend_share_of_step": res["frontend_share_of_step"], "frontend_over_model": res["frontend_over_model"]})
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="2: the headline (BASELINE.json configs[1]); 3: a corpus of --corpus distinct clips sharded i mod R; "
                         "4 / 5: the front-end feeding badwinner2 inference / a wr_resnet_bird training step (reported, not the headline)")
    ap.add_argument("--corpus", type=int, default=1000000, help="clips of the --config 3 corpus")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--pipes", type=int, default=0,
                    help="host pipes (host threads) of the end-to-end leg; 0 = 3 at N <= 2 (measured 84.6 k -> 86.0 k clips/s "
                         "against two; four: 86.7 k), 2 beyond (each pipe pins 3.7 GB of host memory per rank)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-rows", action="store_true", help="skip the standalone timing of the other HBM-bound rows")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: while the run lasts, file descriptor 1 points at stderr, so that anything a
    # library prints there (NCCL's "NCCL version ..." banner is a plain printf when the box exports NCCL_DEBUG=VERSION) cannot
    # land next to it; emit() puts the descriptor back for the one line.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)          # teardown chatter (process-group destruction) stays off stdout too

    args.emit = emit
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 3:
        return run_corpus(args)
    if args.config in (4, 5):
        return run_config45(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
