#!/usr/bin/env python
"""What a Julia session sees: ONE process, HOST arrays in, HOST results out, pl_init_all over ndev GPUs (argv[1], 0 = all). End to end
(H2D + kernels + exchange + D2H) for the three contracted operations at their BASELINE shapes (config 3 as a 1/4 slice when the host
cannot pin 11.6 GB), pinned host buffers, wall clock around the C-ABI call, best of 3 after one warm-up. One JSON line."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "potentiallearning.jl_b200")):
    sys.path.insert(0, p)
import numpy as np

from plb200 import _lib

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib = _lib.init_all(ndev)
nd = _lib.device_count()
out = {"devices": nd, "comm": _lib.comm_info()}
rng = np.random.default_rng(7)


def pinned(shape):
    return _lib.pinned_zeros(shape)


def best(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


# config 2: RBF kernel matrix, N = 20 000, d = 300, Symmetric(:U) storage into pinned zeros(n, n)
N, D = 20000, 300
X = pinned((N, D))
X[:] = 2.0 * rng.standard_normal(D) + rng.standard_normal((N, D)) / np.sqrt(96.0)
K = pinned((N, N))
ms = best(lambda: _lib.check(lib.pl_gram_rbf(_lib.dptr(X), N, D, D, 0, None, 1.0, 1.0, _lib.dptr(K.reshape(-1)), N, _lib.PL_FILL_JULIA_U), "pl_gram_rbf"))
out["config2_gram_rbf_e2e_ms"] = ms
out["config2_gflops_e2e"] = N * (N + 1) * D / (ms * 1e-3) / 1e9
chk = float(K[N - 1, : N].sum())
out["config2_last_row_sum"] = chk
del K

# config 3: learn! normal equations from host rows (pl_normal_eq): R x 500, ws = [100, 1], intercept, lambda
ncfg, nat, Kf = 2500, 96, 500  # a quarter of config 3: 2.89 GB of rows
R = ncfg * (1 + 3 * nat)
A = pinned((R, Kf))
for r0 in range(0, R, 1 << 17):
    A[r0:r0 + (1 << 17)] = rng.standard_normal((min(1 << 17, R - r0), Kf))
b = rng.standard_normal(R)
M, v = np.zeros((Kf + 1, Kf + 1)), np.zeros(Kf + 1)
ms = best(lambda: _lib.check(lib.pl_normal_eq(_lib.dptr(A), R, Kf, Kf, None, ncfg, 100.0, 1.0, _lib.dptr(b), 1, 0.01, _lib.dptr(M.reshape(-1)), _lib.dptr(v)), "pl_normal_eq"))
out["config3_quarter_normal_eq_e2e_ms"] = ms
out["config3_rows_per_s_e2e"] = R / (ms * 1e-3)
out["config3_host_GBps"] = R * Kf * 8 / (ms * 1e-3) / 1e9
out["config3_trace"] = float(np.trace(M))

# config 4: PCAState covariance from host rows (pl_cov), the same 2.89 GB seen as n x 500 rows, centred
C, m = np.zeros((Kf, Kf)), np.zeros(Kf)
ms = best(lambda: _lib.check(lib.pl_cov(_lib.dptr(A), R, Kf, Kf, _lib.dptr(m), 0, 1, _lib.dptr(C.reshape(-1))), "pl_cov"))
out["cov_e2e_ms"] = ms
out["cov_host_GBps"] = R * Kf * 8 / (ms * 1e-3) / 1e9
out["cov_trace"] = float(np.trace(C))
_lib.finalize()
print(json.dumps(out))
