import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import torch, plb200
from plb200.device import DeviceEngine
plb200.init(0); eng = DeviceEngine()
g = torch.Generator(device="cuda").manual_seed(3)
A = torch.randn((722500, 500), device="cuda", dtype=torch.float64, generator=g)
b = torch.randn(722500, device="cuda", dtype=torch.float64, generator=g)
for _ in range(3):
    eng.syrk_partial(A, None, 2500, 100.0, 1.0, b, None, True)
torch.cuda.synchronize()
print("ok")
