#!/usr/bin/env python
"""The built-in symmetric eigensolver (csrc/eigh.cuh) against LAPACK (numpy) and against cuSOLVER Xsyevd on the same box: accuracy on
several matrix classes (random SPD, rank-deficient DotProduct kernel matrices, RBF kernel matrices, covariance matrices, clustered and
graded spectra, identity, Wilkinson) and device time of pl_eigh_dev for both solvers. Writes one JSON document to stdout / argv[1]."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "potentiallearning.jl_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import plb200
from plb200 import _lib

plb200.init(0)
lib = _lib.lib()
dev = torch.device("cuda", 0)


def eigh_dev(A, mode, vectors=True):
    os.environ["PLB200_EIGH"] = mode
    n = A.shape[0]
    dA = torch.from_numpy(A).to(dev).contiguous()
    dW = torch.empty(n, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(lib.pl_eigh_dev(dA.data_ptr(), n, n, _lib.PL_FILL_FULL, 1 if vectors else 0, dW.data_ptr(), st), "pl_eigh_dev")
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    return dW.cpu().numpy(), dA.cpu().numpy().T, ms


def matrices(n, rng):
    B = rng.standard_normal((n, n + 3))
    yield "random_spd", B @ B.T / (n + 3) + np.diag(rng.uniform(0, 2, n))
    X = rng.standard_normal((n, 26)) + 2 * rng.standard_normal(26)
    G = X @ X.T
    nn = np.sqrt(np.diag(G))
    yield "dot_kernel_d26_alpha2", (G / nn[:, None] / nn[None, :]) ** 2
    Xr = rng.standard_normal((n, 30)) / 3
    sq = (Xr * Xr).sum(1)
    yield "rbf_kernel", np.exp(-0.5 * np.maximum(sq[:, None] + sq[None, :] - 2 * Xr @ Xr.T, 0))
    R = rng.standard_normal((4 * n, n)) * np.logspace(0, -6, n)[None, :] + 10 * rng.standard_normal(n)
    Rc = R - R.mean(0)
    yield "covariance_graded", Rc.T @ Rc / (4 * n)
    Qo, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([np.full(n // 3, 1.0), np.full(n // 3, 1.0 + 1e-13), np.linspace(-2, 3, n - 2 * (n // 3))])
    yield "clustered_indefinite", (Qo * lam) @ Qo.T
    yield "identity", np.eye(n)
    W = np.diag(np.abs(np.arange(n) - n // 2).astype(float)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1)
    yield "wilkinson", W


out = {"cases": [], "timing": []}
sizes = [int(s) for s in os.environ.get("EIGH_SIZES", "3,33,64,130,500,1000,2048").split(",")]
for n in sizes:
    rng = np.random.default_rng(n)
    for name, A in matrices(n, rng):
        A = 0.5 * (A + A.T)
        lo = np.linalg.eigvalsh(A)
        scale = max(np.max(np.abs(lo)), 1e-300)
        rec = {"n": n, "matrix": name}
        for mode in ("own", "cusolver"):
            lam, U, ms = eigh_dev(A, mode)
            rec[mode] = {
                "eig_rel": float(np.max(np.abs(lam - lo)) / scale),
                "residual_rel": float(np.max(np.abs(A @ U - U * lam[None, :])) / scale),
                "orth": float(np.max(np.abs(U.T @ U - np.eye(n)))),
                "ascending": bool(np.all(np.diff(lam) >= 0)),
            }
        out["cases"].append(rec)
        print(json.dumps(rec), file=sys.stderr, flush=True)

for n in [int(s) for s in os.environ.get("EIGH_TIME_SIZES", "500,1000,2048,4096").split(",")]:
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n + 3))
    A = B @ B.T / (n + 3)
    rec = {"n": n}
    for mode in ("own", "cusolver"):
        if mode == "own" and n > 4096:
            continue
        ts = []
        for rep in range(4):
            _, _, ms = eigh_dev(A, mode)
            ts.append(ms)
        rec[mode + "_ms"] = float(min(ts[1:]))
        tv = []
        for rep in range(3):
            _, _, ms = eigh_dev(A, mode, vectors=False)
            tv.append(ms)
        rec[mode + "_values_only_ms"] = float(min(tv[1:]))
    out["timing"].append(rec)
    print(json.dumps(rec), file=sys.stderr, flush=True)
os.environ.pop("PLB200_EIGH", None)
s = json.dumps(out, indent=1)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(s)
print(s)
