"""Skinny contraction depth (config-1-like d = 26): the Gram kernel is bound by writing the kernel matrix, not by the DMMA pipe.
Reports GB/s of result written (CUDA events, device-resident)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import torch

import plb200
from plb200 import _lib

plb200.init(0)
out = {}
g = torch.Generator(device="cuda").manual_seed(1)
for n, d in ((50000, 26), (50000, 8), (50000, 64), (1000, 26)):
    X = (torch.randn((n, d), device="cuda", dtype=torch.float64, generator=g) / 5.6 + 2.0 * torch.randn(d, device="cuda", dtype=torch.float64, generator=g)).contiguous()
    K = torch.zeros((n, n), device="cuda", dtype=torch.float64)
    for name, kern in (("dot", plb200.DotProduct()), ("rbf", plb200.RBF(plb200.Euclidean(d)))):
        for fname, fill in (("full", _lib.PL_FILL_FULL), ("juliaU", _lib.PL_FILL_JULIA_U)):
            ts = []
            for it in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                plb200.kernel_matrix_dev(X, kern, out=K, fill=fill)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = min(ts[1:])
            written = n * n * 8 if fname == "full" else n * (n + 128) / 2 * 8
            out[f"{name}_{fname}_n{n}_d{d}"] = {"ms": ms, "write_GBs": written / ms / 1e6, "tflops_N(N+1)d": n * (n + 1) * d / ms / 1e9}
    del K
print(json.dumps(out, indent=1))
