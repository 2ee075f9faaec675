"""Developer tool: per-CTA timing of one SYRK launch (PLB200_SYRK_TRACE hook) at config-3 / config-4 shapes.
Prints the mean duration of off-diagonal vs diagonal units and the per-warp finishing skew."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import numpy as np
import torch

import plb200
from plb200.device import DeviceEngine

case = sys.argv[1] if len(sys.argv) > 1 else "c3"
plb200.init(0)
eng = DeviceEngine()
g = torch.Generator(device="cuda").manual_seed(3)
if case == "c3":
    R, K, ne = 2_890_000, 500, 10_000
    A = torch.randn((R, K), device="cuda", dtype=torch.float64, generator=g)
    b = torch.randn(R, device="cuda", dtype=torch.float64, generator=g)
    run = lambda: eng.syrk_partial(A, None, ne, 100.0, 1.0, b, None, True)
elif case == "c3novec":
    R, K, ne = 2_890_000, 500, 10_000
    A = torch.randn((R, K), device="cuda", dtype=torch.float64, generator=g)
    run = lambda: eng.syrk_partial(A, None, ne, 100.0, 1.0, None, None, False)
else:
    R, K = 2_500_000, 1000
    A = torch.randn((R, K), device="cuda", dtype=torch.float64, generator=g) + 10.0
    m = A.mean(0).contiguous()
    run = (lambda: eng.syrk_partial(A, None, 0, 1.0, 1.0, None, m, False)) if case == "c4" else (lambda: eng.syrk_partial(A, None, 0, 1.0, 1.0, None, None, False))
for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(case, "untraced ms", e0.elapsed_time(e1))
path = f"/tmp/syrk_trace_{case}.txt"
os.environ["PLB200_SYRK_TRACE"] = path
run()
torch.cuda.synchronize()
os.environ.pop("PLB200_SYRK_TRACE")
hdr = open(path).readline().split()
units, splits, T = int(hdr[2]), int(hdr[4]), int(hdr[6])
d = np.loadtxt(path, dtype=np.int64)
noff = T * (T - 1) // 2
t0 = d[:, 2].min()
start = (d[:, 2] - t0) / 1e3
ends = (d[:, 4:12] - t0) / 1e3
dur = ends.max(1) - start
off, dg = d[:, 0] < noff, d[:, 0] >= noff
print(f"units {units} CTAs {splits} T {T}; kernel span {ends.max():.0f} us (column 0 = first unit of the CTA's piece list)")
print(f"off-diagonal units: n={off.sum()} mean {dur[off].mean():.1f} us (min {dur[off].min():.1f} max {dur[off].max():.1f})")
if dg.any():
    print(f"diagonal units:     n={dg.sum()} mean {dur[dg].mean():.1f} us (min {dur[dg].min():.1f} max {dur[dg].max():.1f})  ratio {dur[dg].mean() / dur[off].mean():.3f}")
    w = ends[dg] - start[dg][:, None]
    print("diag unit per-warp mean finish (us):", np.round(w.mean(0), 1))
w = ends[off] - start[off][:, None]
print("off unit per-warp mean finish (us):", np.round(w.mean(0), 1))
# SM occupancy: busy time per SM vs span
sm = d[:, 3]
busy = np.array([dur[sm == s].sum() for s in np.unique(sm)])
print(f"per-SM busy: mean {busy.mean():.0f} us min {busy.min():.0f} max {busy.max():.0f}; CTAs per SM min {np.bincount(sm).min()} max {np.bincount(sm).max()}")
