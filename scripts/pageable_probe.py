"""End-to-end KernelMatrix (config 2) into PAGEABLE host memory (what a plain Julia `zeros(n,n)` is) vs pinned memory."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import numpy as np
import torch

import plb200
from plb200 import _lib

sys.path.insert(0, ROOT)
import bench

plb200.init(0)
lib = _lib.lib()
X = bench.make_features_host()
N, D = X.shape
out = {}


def run(K):
    t0 = time.perf_counter()
    _lib.check(lib.pl_gram_rbf(_lib.dptr(X), N, D, D, 0, None, 1.0, 1.0, _lib.dptr(K.reshape(-1)), N, _lib.PL_FILL_JULIA_U), "pl_gram_rbf")
    return (time.perf_counter() - t0) * 1e3


t0 = time.perf_counter()
Kp = plb200.pinned_zeros((N, N))  # pl_host_alloc
out["pinned_alloc_ms"] = (time.perf_counter() - t0) * 1e3
run(Kp)
out["pinned_ms"] = min(run(Kp) for _ in range(3))
out["pinned_timing"] = plb200.last_timing()
t = []
for _ in range(3):
    K = np.zeros((N, N))  # fresh calloc'ed pages, like Julia's zeros(n, n)
    t.append(run(K))
out["pageable_fresh_zeros_ms"] = t
K = np.zeros((N, N))
K[:] = 0.0  # touched once: pages exist
out["pageable_touched_ms"] = [run(K) for _ in range(3)]
out["pageable_timing"] = plb200.last_timing()
print(json.dumps(out))
