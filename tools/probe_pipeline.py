import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_training_b200 import _runtime as rt
import bench
B = 4096
dev = torch.device("cuda", 0)
x = bench.synth_on_device(torch, B, dev, 20240)
cfg = rt.FrontendConfig(normalize=True, channels=1, out_layout="btm")
params = rt.pcen_params()
plans = [rt.Plan(cfg, 0) for _ in range(2)]
outs = [torch.empty((B, plans[0].n_frames, cfg.n_mels), device=dev) for _ in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]
for p in plans:
    p.workspace_for(B)
def run(steps, two):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    if two:
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
        for i in range(steps):
            with torch.cuda.stream(streams[i & 1]):
                plans[i & 1].frontend_pcen(x, params, outs[i & 1])
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
    else:
        for i in range(steps):
            plans[0].frontend_pcen(x, params, outs[0])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
for two in (False, True, False, True):
    run(4, two)
    ms = run(20, two)
    print("two_streams" if two else "one_stream ", round(ms, 3), "ms/step", round(B / ms * 1e3), "clips/s", float(outs[0][0, :4, :4].sum()), float(outs[1][0, :4, :4].sum()))
