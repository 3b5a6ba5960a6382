#!/usr/bin/env python3
"""Per-kernel digest of an ncu report holding several launches (the single-kernel form is tools/ncu_summary.py):
    python tools/ncu_summary_all.py report.ncu-rep"""
import csv, re, subprocess, sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
EXTRA = re.compile(r"(sm__pipe_tensor|sm__inst_executed_pipe_t(ensor|mem|c)|sm__ops_path_tensor.*tf32.*pct|tmem|utc)", re.I)
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("==", d.get("Kernel Name", ("?", ""))[0])
    for k in KEYS:
        if k in d:
            print("  ", k, d[k])
    for h in sorted(d):     # every tensor-pipe / TMEM metric the report carries that is not zero
        if EXTRA.search(h) and h not in KEYS and ("pct" in h or h.endswith(".sum")):
            try:
                if float(d[h][0].replace(",", "")) != 0.0:
                    print("  ", h, d[h])
            except ValueError:
                pass
    try:
        t = float(d["gpu__time_duration.sum"][0].replace(",", ""))
        tu = d["gpu__time_duration.sum"][1]
        scale = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(tu, 1e-3)
        def b(name):
            v, u = d[name]
            return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        tot = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
        print("   dram GB/s (read + write over the launch):", round(tot / (t * scale) / 1e9, 1))
    except Exception as exc:
        print("   (no dram rate:", exc, ")")
    stalls = []
    for h, (v, u) in d.items():
        if re.search(r"smsp__average_warps_issue_stalled.*per_issue_active", h):
            try:
                if float(v) > 0.04:
                    stalls.append((float(v), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("   stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)))
