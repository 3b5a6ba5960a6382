import json, os, sys
sys.path.insert(0, os.getcwd())
import torch
from audio_training_b200 import _lib
if len(sys.argv) > 1: _lib.LIB_PATH = os.path.abspath(sys.argv[1])
from audio_training_b200 import _runtime as rt
B = 1024
x = torch.rand((B, 144000), device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)) - 0.5
plan = rt.Plan(rt.FrontendConfig(framing="center_zero", power=1, channels=1, normalize=True), 0)
for _ in range(3): out = plan.stft(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(8): out = plan.stft(x)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"lib": os.path.basename(_lib.LIB_PATH), "ms": e0.elapsed_time(e1) / 8, "checksum": float(out.double().sum())}))
