"""The built-in symmetric eigensolver of libplb200 (csrc/eigh.cuh): eigen(Symmetric(Q)) of
dimension_reduction.jl:13 and the eigendecomposition inside EllEnsemble(K) (dpp.jl:26,41). Checker: LAPACK through numpy (the same
syev family the reference reaches); contract 1e-10 (north_star), asserted here at 1e-12 / 1e-11. The CPU half checks the numpy model of
the divide-and-conquer step (scripts/eig_proto/dc_proto.py) that the CUDA kernels restate."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "potentiallearning.jl_b200"), os.path.join(ROOT, "scripts", "eig_proto")):
    if p not in sys.path:
        sys.path.insert(0, p)


def matrix(kind, n, rng):
    if kind == "random_spd":
        B = rng.standard_normal((n, n + 3))
        A = B @ B.T / (n + 3) + np.diag(rng.uniform(0, 2, n))
    elif kind == "dot_kernel":  # config 1: DotProduct(alpha = 2) on d = 26 features: rank <= 351, the rest of the spectrum is one cluster at 0
        X = rng.standard_normal((n, 26)) + 2 * rng.standard_normal(26)
        G = X @ X.T
        nn = np.sqrt(np.diag(G))
        A = (G / nn[:, None] / nn[None, :]) ** 2
    elif kind == "rbf_kernel":
        X = rng.standard_normal((n, 30)) / 3
        sq = (X * X).sum(1)
        A = np.exp(-0.5 * np.maximum(sq[:, None] + sq[None, :] - 2 * X @ X.T, 0))
    elif kind == "covariance":  # config 4 shape: large mean / sigma, graded column scales
        R = rng.standard_normal((4 * n, n)) * np.logspace(0, -6, n)[None, :] + 10 * rng.standard_normal(n)
        Rc = R - R.mean(0)
        A = Rc.T @ Rc / (4 * n)
    elif kind == "clustered_indefinite":
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        lam = np.concatenate([np.full(n // 3, 1.0), np.full(n // 3, 1.0 + 1e-13), np.linspace(-2, 3, n - 2 * (n // 3))])
        A = (Q * lam) @ Q.T
    elif kind == "identity":
        A = np.eye(n)
    elif kind == "wilkinson":
        A = np.diag(np.abs(np.arange(n) - n // 2).astype(float)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1)
    elif kind == "zero":
        A = np.zeros((n, n))
    else:
        raise ValueError(kind)
    return 0.5 * (A + A.T)


def check(A, lam, U, tol_eig=1e-12, tol_res=1e-12, tol_orth=1e-12):
    n = A.shape[0]
    lo = np.linalg.eigvalsh(A)
    scale = max(np.max(np.abs(lo)), 1e-300)
    assert np.all(np.diff(lam) >= 0), "ascending, as eigen() returns"
    assert np.max(np.abs(lam - lo)) / scale < tol_eig
    assert np.max(np.abs(A @ U - U * lam[None, :])) / scale < tol_res, "residual |A U - U diag(lam)|"
    assert np.max(np.abs(U.T @ U - np.eye(n))) < tol_orth, "orthonormal eigenvectors"


# ------------------------------------------------------------------------------------------------------------------ CPU: the numpy model
@pytest.mark.parametrize("kind", ["random_spd", "dot_kernel", "rbf_kernel", "clustered_indefinite", "wilkinson", "identity"])
def test_divide_and_conquer_model(kind):
    import scipy.linalg as sl
    from dc_proto import dc_eigh

    rng = np.random.default_rng(5)
    n = 150
    A = matrix(kind, n, rng)
    H, Qh = sl.hessenberg(A, calc_q=True)  # tridiagonal for a symmetric matrix
    d, e = np.diag(H).copy(), np.diag(H, 1).copy()
    lam, Z, _ = dc_eigh(d, e)
    check(A, lam, Qh @ Z)


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def pl():
    import plb200

    plb200.init(0)
    return plb200


def eigh_with(monkeypatch, mode, A, **kw):
    from plb200.dpp import eigh

    monkeypatch.setenv("PLB200_EIGH", mode)
    return eigh(A, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65, 96, 127, 129, 255, 256, 257, 500, 513, 769, 1000, 1025, 1184])
def test_builtin_eigh_sizes(pl, monkeypatch, n):
    """every template instance of the register-resident tridiagonalisation (8 / 16 / 24 / 32 / 37 chunks), leaf-only and multi-level
    trees, odd sizes (padded leading dimension), against LAPACK"""
    rng = np.random.default_rng(n)
    A = matrix("random_spd", n, rng)
    lam, U = eigh_with(monkeypatch, "own", A)
    check(A, lam, U)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["dot_kernel", "rbf_kernel", "covariance", "clustered_indefinite", "identity", "wilkinson", "zero"])
@pytest.mark.parametrize("n", [130, 1000])
def test_builtin_eigh_matrix_classes(pl, monkeypatch, kind, n):
    """the spectra the reference's callers produce (rank-deficient DotProduct kernel matrices: hundreds of eigenvalues in one cluster at
    zero; RBF; graded covariances) and the classical hard cases for divide and conquer (heavy deflation, clusters 1e-13 apart)"""
    rng = np.random.default_rng(1000 + n)
    A = matrix(kind, n, rng)
    lam, U = eigh_with(monkeypatch, "own", A)
    check(A, lam, U)


@pytest.mark.gpu
def test_builtin_eigh_scaling_triangles_and_values_only(pl, monkeypatch):
    rng = np.random.default_rng(3)
    n = 300
    A = matrix("random_spd", n, rng)
    lo = np.linalg.eigvalsh(A)
    for s in (1e-200, 1e150):  # the matrix is scaled to unit max-norm on the way in
        lam, U = eigh_with(monkeypatch, "own", A * s)
        assert np.max(np.abs(lam / s - lo)) / lo.max() < 1e-12
        assert np.max(np.abs(U.T @ U - np.eye(n))) < 1e-12
    # one triangle is enough (Symmetric(:U) storage: the other triangle holds garbage), and eigenvalues alone
    Au = np.triu(A) + np.tril(rng.standard_normal((n, n)), -1)
    Al = np.tril(A) + np.triu(rng.standard_normal((n, n)), 1)
    assert np.max(np.abs(eigh_with(monkeypatch, "own", Au, fill="U", vectors=False) - lo)) / lo.max() < 1e-12
    assert np.max(np.abs(eigh_with(monkeypatch, "own", Al, fill="L", vectors=False) - lo)) / lo.max() < 1e-12


@pytest.mark.gpu
def test_builtin_eigh_is_deterministic_and_default(pl, monkeypatch):
    """same bits on every call (fixed reduction orders, no atomics on data); the default solver is the built-in one up to n = 2048
    (register-resident rows up to 1184, rows in L2 above) and cuSOLVER beyond (PLB200_EIGH / PLB200_EIGH_OWN_MAX override)"""
    from plb200.dpp import eigh

    rng = np.random.default_rng(11)
    A = matrix("rbf_kernel", 700, rng)
    l1, U1 = eigh_with(monkeypatch, "own", A)
    l2, U2 = eigh_with(monkeypatch, "own", A)
    assert np.array_equal(l1, l2) and np.array_equal(U1, U2)
    monkeypatch.delenv("PLB200_EIGH")
    l3, U3 = eigh(A)
    assert np.array_equal(l1, l3) and np.array_equal(U1, U3), "n = 700 takes the built-in path by default"
    B = matrix("random_spd", 1300, rng)
    lam, U = eigh(B)  # built-in, rows in L2
    check(B, lam, U)
    lam_o, U_o = eigh_with(monkeypatch, "own", B)
    assert np.array_equal(lam, lam_o) and np.array_equal(U, U_o)
    lam_c, U_c = eigh_with(monkeypatch, "cusolver", B)
    check(B, lam_c, U_c)
    assert np.max(np.abs(lam_c - lam)) / lam.max() < 1e-12
    monkeypatch.delenv("PLB200_EIGH")
    C = matrix("random_spd", 2100, rng)
    lam, U = eigh(C)  # cuSOLVER
    check(C, lam, U)


@pytest.mark.gpu
@pytest.mark.parametrize("n,kind", [(1185, "random_spd"), (1777, "rbf_kernel"), (2048, "dot_kernel"), (2049, "random_spd"), (3000, "clustered_indefinite")])
def test_builtin_eigh_rows_in_l2(pl, monkeypatch, n, kind):
    """the L2-resident tridiagonalisation (one and two rows per warp, both template instances) and the back-transformations above
    n = 1024 / 2048"""
    A = matrix(kind, n, np.random.default_rng(n))
    lam, U = eigh_with(monkeypatch, "own", A)
    check(A, lam, U)


@pytest.mark.gpu
def test_builtin_eigh_refuses_nonfinite_input(pl, monkeypatch):
    """Julia's eigen(Symmetric) throws on Inf / NaN (LAPACK chkfinite); the built-in solver returns PL_EINVAL instead of running its
    cooperative kernel on them, and the library stays usable"""
    from plb200 import _lib

    A = matrix("random_spd", 200, np.random.default_rng(1))
    B = A.copy()
    B[5, 7] = B[7, 5] = np.nan
    with pytest.raises(_lib.PLB200Error, match="Infs or NaNs"):
        eigh_with(monkeypatch, "own", B, vectors=False)
    B[5, 7] = B[7, 5] = np.inf
    with pytest.raises(_lib.PLB200Error, match="Infs or NaNs"):
        eigh_with(monkeypatch, "own", B)
    lam, U = eigh_with(monkeypatch, "own", A)
    check(A, lam, U)
