"""ncu driver: a few launches of the SYRK path at config-3 / config-4 shape (usage: syrk_profile_case.py [c3|c3q|c4q])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import torch

import plb200
from plb200.device import DeviceEngine

case = sys.argv[1] if len(sys.argv) > 1 else "c3q"
plb200.init(0)
eng = DeviceEngine()
g = torch.Generator(device="cuda").manual_seed(3)
if case in ("c3", "c3q"):
    R, ne = (2_890_000, 10_000) if case == "c3" else (722_500, 2_500)
    A = torch.randn((R, 500), device="cuda", dtype=torch.float64, generator=g)
    b = torch.randn(R, device="cuda", dtype=torch.float64, generator=g)
    run = lambda: eng.syrk_partial(A, None, ne, 100.0, 1.0, b, None, True)
else:
    A = torch.randn((2_500_000, 1000), device="cuda", dtype=torch.float64, generator=g) + 10.0
    m = eng.scale_div(eng.colsum(A), float(A.shape[0]))
    run = lambda: eng.syrk_partial(A, None, 0, 1.0, 1.0, None, m, False)
for _ in range(3):
    run()
torch.cuda.synchronize()
print("ok")
