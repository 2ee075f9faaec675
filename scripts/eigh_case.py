#!/usr/bin/env python
"""One pl_eigh_dev call of the built-in solver at size argv[1] (after one warm-up call): the ncu target behind profiles/*eigh*launches*.csv."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "potentiallearning.jl_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import plb200
from plb200 import _lib

plb200.init(0)
lib = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
os.environ["PLB200_EIGH"] = sys.argv[2] if len(sys.argv) > 2 else "own"
rng = np.random.default_rng(n)
B = rng.standard_normal((n, n + 3))
A = B @ B.T / (n + 3)
for rep in range(2):
    dA = torch.from_numpy(A).cuda()
    dW = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.check(lib.pl_eigh_dev(dA.data_ptr(), n, n, _lib.PL_FILL_FULL, 1, dW.data_ptr(), torch.cuda.current_stream().cuda_stream), "pl_eigh_dev")
torch.cuda.synchronize()
print("ok", n, float(dW[-1]))
