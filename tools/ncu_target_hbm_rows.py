#!/usr/bin/env python3
"""Target of the ncu capture of the HBM-bound kernels of the benchmarked step (K0 row min/max, the two PCEN passes, the
extremes fold) and of the standalone min-max epilogue, 1024 clips per launch:

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/hbm_rows \
        -k regex:'row_minmax|pcen_|stats_kernel|compress_kernel' python tools/ncu_target_hbm_rows.py
    python tools/ncu_summary_all.py gpurun_out/hbm_rows.ncu-rep > profiles/r01_hbm_rows_ncu_summary.txt
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_training_b200 import _runtime as rt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.rand((B, 144000), device="cuda") - 0.5
plan = rt.Plan(rt.FrontendConfig(normalize=True, channels=1, out_layout="btm"), 0)
params = rt.pcen_params()
out = torch.empty((B, plan.n_frames, 160), dtype=torch.float32, device="cuda")
plan.frontend_pcen(x, params, out)          # warm-up (tables, workspace)
mel = plan.frontend(x)
plan.compress(mel, "minmax")
torch.cuda.synchronize()
torch.cuda.profiler.start()
plan.frontend_pcen(x, params, out)
plan.compress(mel, "minmax")
torch.cuda.synchronize()
torch.cuda.profiler.stop()
