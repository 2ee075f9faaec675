"""Times the SYRK path (device-resident, CUDA events) at a given row count; PLB200_SYRK_WAVES is read once per process, so the caller
loops over it: `for w in 6 10 14 20; do PLB200_SYRK_WAVES=$w python scripts/syrk_waves_probe.py 361250 500; done`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import torch

import plb200
from plb200.device import DeviceEngine

R = int(sys.argv[1]) if len(sys.argv) > 1 else 361_250
K = int(sys.argv[2]) if len(sys.argv) > 2 else 500
centred = len(sys.argv) > 3 and sys.argv[3] == "c"
plb200.init(0)
eng = DeviceEngine()
g = torch.Generator(device="cuda").manual_seed(3)
A = torch.randn((R, K), device="cuda", dtype=torch.float64, generator=g)
if centred:
    A += 10.0
    m = eng.scale_div(eng.colsum(A), float(R))
    run = lambda: eng.syrk_partial(A, None, 0, 1.0, 1.0, None, m, False)
else:
    b = torch.randn(R, device="cuda", dtype=torch.float64, generator=g)
    ne = R // 289
    run = lambda: eng.syrk_partial(A, None, ne, 100.0, 1.0, b, None, True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    run()
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print(json.dumps({"R": R, "K": K, "centred": centred, "waves": os.environ.get("PLB200_SYRK_WAVES", "default"), "ms_min": ts[0], "ms_med": ts[len(ts) // 2],
                  "tflops": R * K * (K + 1) / ts[0] / 1e9}))
