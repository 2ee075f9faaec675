"""ncu driver for the HBM-bound kernels: skinny-depth DotProduct Gram (config-1 shape at N = 50 000, d = 26, Symmetric(:U) storage),
the GlobalMean segmented reduction at config-2 shape, and the column-sum pass at config-4 width. Three launches of each."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import torch

import plb200
from plb200 import _lib
from plb200.device import DeviceEngine

plb200.init(0)
eng = DeviceEngine()
g = torch.Generator(device="cuda").manual_seed(1)
n, d = 50000, 26
X = (torch.randn((n, d), device="cuda", dtype=torch.float64, generator=g) / 5.6 + 2.0 * torch.randn(d, device="cuda", dtype=torch.float64, generator=g)).contiguous()
K = torch.zeros((n, n), device="cuda", dtype=torch.float64)
for _ in range(3):
    plb200.kernel_matrix_dev(X, plb200.DotProduct(), out=K, fill=_lib.PL_FILL_JULIA_U)
torch.cuda.synchronize()
del K
N, nat, dd = 20000, 96, 300
L = torch.randn((N * nat, dd), device="cuda", dtype=torch.float64, generator=g)
off = (torch.arange(N + 1, device="cuda", dtype=torch.int64) * nat).contiguous()
Xo = torch.empty((N, dd), device="cuda", dtype=torch.float64)
for _ in range(3):
    plb200.global_features_dev(L, off, True, out=Xo)
torch.cuda.synchronize()
del L
A = torch.randn((2_500_000, 1000), device="cuda", dtype=torch.float64, generator=g)
for _ in range(3):
    eng.colsum(A)
torch.cuda.synchronize()
print("ok")
