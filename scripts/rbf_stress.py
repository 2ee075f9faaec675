"""One-off stress of the RBF path: a few hundred random shapes (n 1..700, d 1..70, cross n2 1..200, the three Cinv kinds, both fills) against
numpy fp64 at the contract's 1e-10. Not part of the test suite (the suite's sweeps cover the same ground with fixed seeds)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import numpy as np

import plb200 as pl

pl.init(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 7)
worst = 0.0
ntr = int(sys.argv[2]) if len(sys.argv) > 2 else 300


def ref(X1, X2, C, ell, beta):
    out = np.empty((len(X1), len(X2)))
    for a in range(len(X1)):
        U = X1[a] - X2
        out[a] = beta * np.exp(-0.5 * np.einsum("jk,jk->j", U @ C, U) / ell ** 2)
    return out


for trial in range(ntr):
    n = int(rng.integers(1, 700))
    d = int(rng.integers(1, 71))
    X = rng.standard_normal((n, d)) * rng.uniform(0.1, 1.0) + rng.standard_normal(d) * rng.uniform(0.0, 5.0)
    ell, beta = float(rng.uniform(0.3, 4.0)), float(rng.choice([1.0, 0.3, 2.5, 40.0]))
    kind = trial % 3
    if kind == 0:
        e, C = pl.Euclidean(d), np.eye(d)
    elif kind == 1:
        cd = rng.uniform(0.3, 3.0, size=d)
        e, C = pl.Euclidean(pl.Diagonal(cd)), np.diag(cd)
    else:
        B = rng.standard_normal((d, d))
        C = B @ B.T / d + np.eye(d)
        e = pl.Euclidean(pl.Symmetric(C))
    K = np.asarray(pl.KernelMatrix(X, pl.RBF(e, ell=ell, beta=beta)))
    R = ref(X, X, C, ell, beta)
    ok = R > 1e-250
    err = float(np.max(np.abs(K[ok] - R[ok]) / R[ok]))
    assert np.array_equal(K, K.T) and np.all(np.diag(K) == beta), (trial, n, d, kind)
    n2 = int(rng.integers(1, 200))
    X2 = rng.standard_normal((n2, d)) * 0.5 + X.mean(0)
    Kx = pl.KernelMatrix(X, X2, pl.RBF(e, ell=ell, beta=beta))
    Rx = ref(X, X2, C, ell, beta)
    okx = Rx > 1e-250
    errx = float(np.max(np.abs(Kx[okx] - Rx[okx]) / Rx[okx])) if okx.any() else 0.0
    worst = max(worst, err, errx)
    assert err < 1e-10 and errx < 1e-10, (trial, n, d, kind, n2, err, errx)
print("trials", ntr, "worst relative error", worst)
