#!/usr/bin/env python
"""Random stress of the built-in eigensolver (pl_eigh through the host entry point, default routing): random sizes, random spectra
(uniform, geometric over 1e-14..1, clustered, low rank, with exact repeats, indefinite), random orthogonal bases, random scales and
leading dimensions / triangles. Reports the worst eigenvalue error, residual and orthogonality against LAPACK."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "potentiallearning.jl_b200")):
    sys.path.insert(0, p)
import numpy as np

import plb200
from plb200.dpp import eigh

plb200.init(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 7)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 300
worst = {"eig_rel": 0.0, "residual_rel": 0.0, "orth": 0.0}
worst_case = {}
fails = []
for it in range(ncase):
    n = int(rng.choice([rng.integers(1, 40), rng.integers(40, 300), rng.integers(300, 1300), rng.integers(1300, 2049)], p=[0.3, 0.4, 0.25, 0.05]))
    kind = rng.choice(["uniform", "geometric", "clustered", "lowrank", "repeats", "indefinite", "tridiagonal", "diagonal"])
    if kind == "uniform":
        lam = rng.uniform(0, 1, n)
    elif kind == "geometric":
        lam = np.logspace(0, -14, n) * rng.choice([1, -1], n)
    elif kind == "clustered":
        centers = rng.uniform(-1, 1, max(1, n // 20 + 1))
        lam = rng.choice(centers, n) + 1e-13 * rng.standard_normal(n)
    elif kind == "lowrank":
        lam = np.zeros(n)
        k = max(1, n // 10)
        lam[:k] = rng.uniform(0.5, 2, k)
    elif kind == "repeats":
        lam = rng.choice(rng.uniform(-1, 1, 3), n)
    else:
        lam = rng.standard_normal(n)
    if kind == "tridiagonal":
        A = np.diag(rng.standard_normal(n)) + np.diag(rng.standard_normal(n - 1) * (rng.random(n - 1) > 0.2), 1)
        A = A + np.triu(A, 1).T
    elif kind == "diagonal":
        A = np.diag(lam)
    else:
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        A = (Q * lam) @ Q.T
    A = 0.5 * (A + A.T) * float(10.0 ** rng.integers(-30, 30))
    ref = np.linalg.eigvalsh(A)
    scale = max(np.max(np.abs(ref)), 1e-300)
    fill = str(rng.choice(["full", "U", "L"]))
    B = A.copy()
    if fill == "U":
        B = np.triu(A) + np.tril(rng.standard_normal((n, n)), -1)
    elif fill == "L":
        B = np.tril(A) + np.triu(rng.standard_normal((n, n)), 1)
    lamg, U = eigh(B, fill=fill)
    m = {"eig_rel": float(np.max(np.abs(lamg - ref)) / scale), "residual_rel": float(np.max(np.abs(A @ U - U * lamg[None, :])) / scale),
         "orth": float(np.max(np.abs(U.T @ U - np.eye(n))))}
    for k2, v in m.items():
        if not (v <= worst[k2]):
            worst[k2] = v
            worst_case[k2] = {"n": n, "kind": str(kind), "fill": fill}
    if not (m["eig_rel"] < 1e-12 and m["residual_rel"] < 1e-12 and m["orth"] < 1e-12 and np.all(np.diff(lamg) >= 0)):
        fails.append({"n": n, "kind": str(kind), "fill": fill, **m})
out = {"cases": ncase, "worst": worst, "worst_case": worst_case, "failures": fails[:20], "n_failures": len(fails)}
print(json.dumps(out, indent=1))
