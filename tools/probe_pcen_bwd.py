import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_training_b200 import _runtime as rt
plan = rt.get_plan(rt.FrontendConfig(), 0)
x = torch.rand(1024, 513, 160, device="cuda") * 3 + 0.01
g = torch.randn_like(x)
p = rt.pcen_params(norm_scope="none")
for _ in range(3):
    plan.pcen_backward(x, g, p)
torch.cuda.synchronize()
