"""oracle.py — CPU restatement (numpy / pure Python) of the reference algorithms on the hot path of
PotentialLearning.jl v0.2.7. Citations are into /root/reference.

TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (potentiallearning.jl_b200/) never does and has no CPU fallback.

PARITY UNPINNED: the reference is pure Julia; Julia is not installed here nor on the GPU box, and the reference's
tests carry no golden values for this path (only properties and types). The functions below
follow the reference line by line in its own operation order; tests/golden/make_golden.py cross-checks them against an
independent extended-precision evaluation of the same formulas, and tests/test_oracle.py re-states every property
the reference's tests assert (test/kernels/kernel_tests.jl, test/learning/linear_tests.jl,
test/dimension_reduction/dimension_reduction_tests.jl).

The C half (oracle/pl_oracle.c -> oracle/_build/libploracle.so) holds the same loops for sizes where Python is too slow.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libploracle.so")


# ----------------------------------------------------------------------------------------------------------------
# C half
# ----------------------------------------------------------------------------------------------------------------
def build_c(force: bool = False) -> str:
    """gcc-compile oracle/pl_oracle.c (strict fp64, no fast-math, no FMA contraction)."""
    src = os.path.join(_HERE, "pl_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off", "-o", _LIB_PATH, src, "-lm"]
        )
    return _LIB_PATH


_clib = None


def clib():
    global _clib
    if _clib is None:
        build_c()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        i64 = ctypes.c_int64
        L.orc_kernel_matrix_sym.argtypes = [ctypes.c_int, dp, i64, i64, i64, ctypes.c_int, ctypes.c_int, dp, ctypes.c_double,
                                            ctypes.c_double, dp, i64, i64, i64]
        L.orc_kernel_matrix_cross.argtypes = [ctypes.c_int, dp, i64, i64, dp, i64, i64, i64, ctypes.c_int, ctypes.c_int, dp,
                                              ctypes.c_double, ctypes.c_double, dp, i64]
        L.orc_outer_sum.argtypes = [dp, i64, i64, i64, dp, dp, dp, dp, dp]
        L.orc_atqa.argtypes = [dp, i64, i64, i64, dp, dp, dp, dp]
        L.orc_atqa_blocked.argtypes = [dp, i64, i64, i64, dp, dp]
        for f in (L.orc_kernel_matrix_sym, L.orc_kernel_matrix_cross, L.orc_outer_sum, L.orc_atqa, L.orc_atqa_blocked):
            f.restype = ctypes.c_int
        _clib = L
    return _clib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _c64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


# ----------------------------------------------------------------------------------------------------------------
# Kernels: features, distances, kernels (src/Kernels)
# ----------------------------------------------------------------------------------------------------------------
def global_mean(local_descriptors) -> np.ndarray:
    """compute_feature(B::LocalDescriptors, ::GlobalMean) = mean(get_values(B))  (features.jl:32-34):
    sequential sum of the per-atom vectors, divided by natoms."""
    vals = [np.asarray(v, dtype=np.float64) for v in local_descriptors]
    s = vals[0].copy()
    for v in vals[1:]:
        s = s + v
    return s / len(vals)


def global_sum(local_descriptors) -> np.ndarray:
    """sum(get_values(B)) (features.jl:19-21; linear-learning-problem.jl:75; pca_state.jl:23)."""
    vals = [np.asarray(v, dtype=np.float64) for v in local_descriptors]
    s = vals[0].copy()
    for v in vals[1:]:
        s = s + v
    return s


def julia_sum(vals, blksize: int = 1024):
    """Base.sum over a Vector of vectors = mapreduce_impl(identity, +, A, 1, n, pairwise_blocksize = 1024) (Julia Base
    reduce.jl, stdlib — not under /root/reference; restated from its published source): fewer than `blksize` elements are
    added left to right; longer ranges split at imid = ifirst + (ilast - ifirst) >> 1 and the halves are summed recursively.
    global_sum / global_mean above are the n < 1024 case of this."""
    vals = [np.asarray(v, dtype=np.float64) for v in vals]

    def impl(lo, hi):  # inclusive, 0-based
        if lo == hi:
            return vals[lo].copy()
        if hi - lo < blksize:
            v = vals[lo] + vals[lo + 1]
            for i in range(lo + 2, hi + 1):
                v = v + vals[i]
            return v
        mid = lo + ((hi - lo) >> 1)
        return impl(lo, mid) + impl(mid + 1, hi)

    return impl(0, len(vals) - 1)


def global_features(local_blocks, mean: bool) -> np.ndarray:
    """compute_features(ds, GlobalMean() / GlobalSum()) (features.jl:76-78 over :19-21 / :32-34): one row per configuration."""
    rows = []
    for b in local_blocks:
        s = julia_sum(list(np.asarray(b, dtype=np.float64)))
        rows.append(s / len(b) if mean else s)
    return np.stack(rows)


@dataclass
class Euclidean:
    """Euclidean(dim) / Euclidean(Cinv) (distances.jl:39-51). cinv: None (identity), 1-D (Diagonal) or 2-D (Symmetric)."""
    dim: int
    cinv: np.ndarray | None = None

    @property
    def kind(self) -> int:
        return 0 if self.cinv is None else (1 if np.ndim(self.cinv) == 1 else 2)


@dataclass
class DotProduct:
    alpha: int = 2  # kernels.jl:21-26


@dataclass
class RBF:
    d: Euclidean
    alpha: float = 1e-8  # stored, never applied (kernels.jl:73-74)
    ell: float = 1.0
    beta: float = 1.0


@dataclass
class InverseMultiquadric:
    d: Euclidean
    c2: float = 1.0  # kernels.jl:139-152
    ell: float = 1.0


def correlation_matrix(local_descriptors) -> np.ndarray:
    """compute_feature(B, ::CorrelationMatrix) = Symmetric(mean(Bi*Bi' for Bi in B)) (features.jl:48-50)."""
    vals = [np.asarray(v, dtype=np.float64) for v in local_descriptors]
    acc = np.zeros((len(vals[0]), len(vals[0])))
    for v in vals:
        acc = acc + np.outer(v, v)
    return acc / len(vals)


def compute_distance(b1, b2, e: Euclidean) -> float:
    """(B1 - B2)' * e.Cinv * (B1 - B2)  (distances.jl:58-60)."""
    if np.ndim(b1) == 2:  # Symmetric features: tr(Csqrt*(C1-C2)'*Cinv*(C1-C2)*Csqrt) (distances.jl:62-68); identity Cinv here
        assert e.cinv is None
        D = np.asarray(b1, dtype=np.float64) - np.asarray(b2, dtype=np.float64)
        return float(np.trace(D.T @ D))
    v = np.asarray(b1, dtype=np.float64) - np.asarray(b2, dtype=np.float64)
    if e.cinv is None:
        c = 1.0 * v  # 1.0*I(dim)
    elif np.ndim(e.cinv) == 1:
        c = np.asarray(e.cinv) * v
    else:
        return float((v @ np.asarray(e.cinv)) @ v)
    return float(np.dot(v, c))


def compute_kernel(a, b, k) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if isinstance(k, DotProduct):
        # (dot(A,B) / sqrt(dot(A,A)*dot(B,B)))^α   (kernels.jl:35)
        # dot() of two matrices is the sum of elementwise products (LinearAlgebra.dot), so Symmetric features work unchanged
        return float((np.sum(a * b) / np.sqrt(np.sum(a * a) * np.sum(b * b))) ** int(k.alpha)) if a.ndim == 2 else \
            float((np.dot(a, b) / np.sqrt(np.dot(a, a) * np.dot(b, b))) ** int(k.alpha))
    if isinstance(k, RBF):
        d2 = compute_distance(a, b, k.d)
        return float(k.beta * np.exp(-0.5 * d2 / k.ell ** 2))  # kernels.jl:73-74
    if isinstance(k, InverseMultiquadric):
        d2 = compute_distance(a, b, k.d)
        return float((k.c2 + d2 / k.ell ** 2) ** (-0.5))  # kernels.jl:162-163
    raise TypeError(k)


def kernel_matrix(F, k, use_c: bool | None = None) -> np.ndarray:
    """KernelMatrix(F,k) (kernels.jl:211-223). Returns the n x n array whose [i, j] entry for i <= j is k(F_i, F_j) and
    whose strict lower triangle is 0 — i.e. `K.data` of the reference's Symmetric(K) (uplo = :U)."""
    X = np.ascontiguousarray(np.asarray(F, dtype=np.float64))
    if X.ndim == 3:  # Vector{Symmetric}: pure-Python pair loop on the matrices themselves
        n = X.shape[0]
        K = np.zeros((n, n))
        for i in range(n):
            for j in range(i, n):
                K[i, j] = compute_kernel(X[i], X[j], k)
        return K
    n, d = X.shape
    if use_c is None:
        use_c = n > 64
    K = np.zeros((n, n))
    if use_c:
        kid, alpha, kind, cinv, ell, beta = _kernel_args(k)
        # C writes Julia column-major K[i + j*ld] for j >= i; viewed row-major that is KT = K.T
        KT = np.zeros((n, n))
        rc = clib().orc_kernel_matrix_sym(kid, _p(X), n, d, d, alpha, kind, _p(cinv), ell, beta, _p(KT), n, 0, n)
        assert rc == 0
        return np.ascontiguousarray(KT.T)
    for i in range(n):
        for j in range(i, n):
            K[i, j] = compute_kernel(X[i], X[j], k)
    return K


def symmetric(Kdata: np.ndarray) -> np.ndarray:
    """Symmetric(K) with uplo=:U (kernels.jl:222): mirror the upper triangle."""
    return np.triu(Kdata) + np.triu(Kdata, 1).T


def kernel_matrix_cross(F1, F2, k, use_c: bool | None = None) -> np.ndarray:
    """KernelMatrix(F1,F2,k) (kernels.jl:229-244): m x n."""
    X1 = np.ascontiguousarray(np.asarray(F1, dtype=np.float64))
    X2 = np.ascontiguousarray(np.asarray(F2, dtype=np.float64))
    m, d = X1.shape
    n = X2.shape[0]
    if use_c is None:
        use_c = m * n > 4096
    if use_c:
        kid, alpha, kind, cinv, ell, beta = _kernel_args(k)
        KT = np.zeros((n, m))  # column-major m x n
        rc = clib().orc_kernel_matrix_cross(kid, _p(X1), m, d, _p(X2), n, d, d, alpha, kind, _p(cinv), ell, beta, _p(KT), m)
        assert rc == 0
        return np.ascontiguousarray(KT.T)
    K = np.zeros((m, n))
    for i in range(m):
        for j in range(n):
            K[i, j] = compute_kernel(X1[i], X2[j], k)
    return K


def _kernel_args(k):
    if isinstance(k, DotProduct):
        return 1, int(k.alpha), 0, None, 1.0, 1.0
    if isinstance(k, RBF):
        cinv = None if k.d.cinv is None else _c64(k.d.cinv)
        return 2, 0, k.d.kind, cinv, float(k.ell), float(k.beta)
    if isinstance(k, InverseMultiquadric):
        cinv = None if k.d.cinv is None else _c64(k.d.cinv)
        return 3, 0, k.d.kind, cinv, float(k.ell), float(k.c2)
    raise TypeError(k)


def kernel_matrix_rows(F, k, row_begin: int, row_end: int) -> int:
    """Bounded sample of the reference pair loop: rows [row_begin, row_end) of the upper triangle (kernels.jl:217-221).
    Returns the number of kernel evaluations done (used by bench.py to time the CPU baseline)."""
    X = np.ascontiguousarray(np.asarray(F, dtype=np.float64))
    n, d = X.shape
    kid, alpha, kind, cinv, ell, beta = _kernel_args(k)
    strip = np.zeros((n, row_end))  # column-major row_end x n: K[i + j*row_end], i < row_end
    rc = clib().orc_kernel_matrix_sym(kid, _p(X), n, d, d, alpha, kind, _p(cinv), ell, beta, _p(strip), row_end, row_begin, row_end)
    assert rc == 0
    return sum(n - i for i in range(row_begin, row_end))


# ----------------------------------------------------------------------------------------------------------------
# Learning: LinearProblem layout and normal-equation assembly (src/Learning)
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class UnivariateLinearProblem:
    iv_data: list
    dv_data: list
    beta: np.ndarray
    beta0: np.ndarray
    sigma: np.ndarray = field(default_factory=lambda: np.array([1.0]))


@dataclass
class CovariateLinearProblem:
    e: list
    f: list
    B: list
    dB: list
    beta: np.ndarray
    beta0: np.ndarray


def linear_problem(local_descriptors=None, energies=None, force_descriptors=None, forces=None):
    """LinearProblem(ds) (linear-learning-problem.jl:71-145) on plain nested lists:
    local_descriptors[c][atom] -> K-vector; force_descriptors[c][atom][xyz] -> K-vector; forces[c][atom] -> 3-vector."""
    d_flag = local_descriptors is not None and energies is not None
    fd_flag = force_descriptors is not None and forces is not None
    if d_flag:
        descriptors = [global_sum(ld) for ld in local_descriptors]  # :75 sum.(get_values.(...))
    if fd_flag:
        fdesc = [[np.asarray(comp, dtype=np.float64) for atom in cfg for comp in atom] for cfg in force_descriptors]  # :81 reduce(vcat, ...)
        fvals = [np.concatenate([np.asarray(fa, dtype=np.float64) for fa in cfg]) for cfg in forces]  # :126 reduce(vcat, fi)
        dB = [np.stack(fi, axis=1) for fi in fdesc]  # :127 reduce(hcat, fi): K x 3natoms
    if d_flag and not fd_flag:
        dim = len(descriptors[0])
        return UnivariateLinearProblem(descriptors, list(energies), np.zeros(dim), np.zeros(1))
    if fd_flag and not d_flag:
        dim = dB[0].shape[0]
        return UnivariateLinearProblem(dB, fvals, np.zeros(dim), np.zeros(1))
    if d_flag and fd_flag:
        dim = len(descriptors[0])
        if dim != dB[0].shape[0]:
            raise ValueError("Descriptors and Force Descriptors have different dimension!")
        return CovariateLinearProblem(list(energies), fvals, descriptors, dB, np.zeros(dim), np.zeros(1))
    raise ValueError("Either no (Energy, Descriptors) or (Forces, Force Descriptors) in DataSet")


def wls_matrices(lp, ws, intercept: bool):
    """A, b, q of learn!(lp, ws, int) (linear-learn.jl:183-197 univariate, :232-249 covariate)."""
    if isinstance(lp, UnivariateLinearProblem):
        iv = lp.iv_data
        if np.ndim(iv[0]) == 1:
            B_train = np.stack(iv, axis=0)  # reduce(hcat, iv)'
            b = np.asarray(lp.dv_data, dtype=np.float64)
        else:  # force-only: iv[i] is K x 3natoms
            B_train = np.concatenate([m.T for m in iv], axis=0)
            b = np.concatenate(lp.dv_data)
        A = np.hstack([np.ones((B_train.shape[0], 1)), B_train]) if intercept else B_train
        q = ws[0] * np.ones(len(b))
        return A, b, q
    B_train = np.stack(lp.B, axis=0)
    dB_train = np.concatenate([m.T for m in lp.dB], axis=0)
    e_train = np.asarray(lp.e, dtype=np.float64)
    f_train = np.concatenate(lp.f)
    A = np.vstack([B_train, dB_train])
    if intercept:
        int_col = np.concatenate([np.ones(B_train.shape[0]), np.zeros(dB_train.shape[0])])  # :240
        A = np.hstack([int_col[:, None], A])
    b = np.concatenate([e_train, f_train])
    q = np.concatenate([ws[0] * np.ones(len(e_train)), ws[1] * np.ones(len(f_train))])
    return A, b, q


def normal_equations(A, b, q, lam: float = 0.0, use_c: bool = False):
    """(A'*Q*A + λ*I, A'*Q*b)  (linear-learn.jl:200, :253)."""
    A = _c64(A)
    q = _c64(q)
    b = _c64(b)
    n = A.shape[1]
    if use_c:
        M = np.zeros((n, n))
        v = np.zeros(n)
        rc = clib().orc_atqa(_p(A), A.shape[0], n, n, _p(q), _p(b), _p(M), _p(v))
        assert rc == 0
    else:
        QA = q[:, None] * A
        M = A.T @ QA
        v = A.T @ (q * b)
    return M + lam * np.eye(n), v


def learn_wls(lp, ws, intercept: bool, lam: float = 0.0):
    """learn!(lp, ws, int; λ) (linear-learn.jl:178-215, 226-268): solve and update lp.β / lp.β0 in place."""
    A, b, q = wls_matrices(lp, ws, intercept)
    M, v = normal_equations(A, b, q, lam if isinstance(lp, CovariateLinearProblem) else 0.0)
    try:
        bs = np.linalg.solve(M, v)
    except np.linalg.LinAlgError:
        bs = np.linalg.pinv(M) @ v
    if intercept:
        lp.beta0[:] = bs[0]
        lp.beta[:] = bs[1:]
    else:
        lp.beta[:] = bs
    return M, v


def ols_sums(iv_data, dv_data):
    """AtA = sum(v*v' for v in iv), Atb = sum(v*b ...) (linear-learn.jl:16-17; also :45-49 per block).
    v is a K-vector (b scalar) or a K x m matrix (b an m-vector)."""
    K = np.asarray(iv_data[0]).shape[0]
    AtA = np.zeros((K, K))
    Atb = np.zeros(K)
    for v, b in zip(iv_data, dv_data):
        v = np.asarray(v, dtype=np.float64)
        if v.ndim == 1:
            AtA = AtA + np.outer(v, v)
            Atb = Atb + v * b
        else:
            AtA = AtA + v @ v.T
            Atb = Atb + v @ np.asarray(b, dtype=np.float64)
    return AtA, Atb


def pinv_julia(M, tol: float):
    """pinv(M, tol) as linear-learn.jl:19 calls it. In Julia's LinearAlgebra [external, stdlib] the positional form is
    pinv(M; rtol = tol): with the SVD M = U S V', singular values S .> max(rtol * maximum(S), atol = 0) are inverted, the others zeroed."""
    M = np.asarray(M, dtype=np.float64)
    U, S, Vt = np.linalg.svd(M)
    cut = max(tol * (S.max() if S.size else 0.0), 0.0)
    Sinv = np.array([1.0 / x if x > cut else 0.0 for x in S])
    return (Vt.T * Sinv) @ U.T


def learn_ols(iv_data, dv_data, alpha: float):
    """learn!(lp::UnivariateLinearProblem, α) (linear-learn.jl:11-23): returns (β, σ, Σ)."""
    AtA, Atb = ols_sums(iv_data, dv_data)  # :16-17
    Q = pinv_julia(AtA, alpha)  # :19
    beta = Q @ Atb  # :20
    sigma = np.std(Atb - AtA @ beta, ddof=1)  # :21  (Statistics.std is the corrected sample deviation)
    return beta, sigma, sigma ** 2 * Q  # :22


def ooc_accumulate(configs, ws=(30.0, 1.0), eweight_normalized="squared"):
    """AtWA .+= A'*W*A; AtWb .+= A'*W*b per configuration (linear-learn.jl:328-357).
    configs: iterable of (global_descr K, force_descrs 3natoms x K, ref_energy, ref_forces 3natoms)."""
    AtWA = None
    for gd, fdm, e, f in configs:
        gd = np.asarray(gd, dtype=np.float64)
        fdm = np.asarray(fdm, dtype=np.float64)
        A = np.vstack([gd[None, :], fdm])
        b = np.concatenate([[e], np.asarray(f, dtype=np.float64)])
        natoms = fdm.shape[0] // 3
        if eweight_normalized is None:
            we = 1.0
        elif eweight_normalized == "standard":
            we = 1 / natoms
        elif eweight_normalized == "squared":
            we = 1 / natoms ** 2
        else:
            raise ValueError("eweight_normalized can only be nothing, :standard, or :squared")
        w = np.concatenate([[we * ws[0]], ws[1] * np.ones(len(f))])
        if AtWA is None:
            AtWA = np.zeros((A.shape[1], A.shape[1]))
            AtWb = np.zeros(A.shape[1])
        AtWA += A.T @ (w[:, None] * A)
        AtWb += A.T @ (w * b)
    return AtWA, AtWb


# ----------------------------------------------------------------------------------------------------------------
# DimensionReduction (src/DimensionReduction)
# ----------------------------------------------------------------------------------------------------------------
def second_moment(rows, use_c: bool | None = None) -> np.ndarray:
    """Q = mean(di*di' for di in d) (dimension_reduction.jl:12): sequential sum of outer products, divided by n."""
    X = np.ascontiguousarray(np.asarray(rows, dtype=np.float64))
    n, K = X.shape
    if use_c is None:
        use_c = n * K * K > 2_000_000
    if use_c:
        S = np.zeros((K, K))
        rc = clib().orc_outer_sum(_p(X), n, K, K, None, None, None, _p(S), None)
        assert rc == 0
    else:
        S = np.zeros((K, K))
        for r in range(n):
            S = S + np.outer(X[r], X[r])
    return S / n


def compute_eigen(rows):
    """eigen(Symmetric(Q)) (dimension_reduction.jl:11-14): ascending eigenvalues, eigenvectors in columns."""
    Q = second_moment(rows)
    lam, phi = np.linalg.eigh(symmetric(Q))
    return lam, phi


def select_eigendirections(rows, tol):
    """dimension_reduction.jl:15-28."""
    lam, phi = compute_eigen(rows)
    lam, phi = lam[::-1], phi[:, ::-1]
    if isinstance(tol, (int, np.integer)) and not isinstance(tol, bool):
        W = phi[:, :tol]
    else:
        resid = 1.0 - np.cumsum(lam) / np.sum(lam)
        W = phi[:, resid > tol]
    return lam, W


@dataclass
class PCAState:
    tol: float | int = 0.01
    lam: np.ndarray | None = None
    W: np.ndarray | None = None
    m: np.ndarray | None = None  # [] in the reference


def pcastate_fit(local_descriptors, pca: PCAState, force_descriptor_rows=None):
    """fit!(ds, pca::PCAState) (pca_state.jl:20-40). As written in the reference the force-descriptor branch always throws
    (undefined `fd`, pca_state.jl:29) and is swallowed, so only the global-sum descriptors are used; pass
    force_descriptor_rows to opt into the intended stacked-rows mode (BASELINE config 4)."""
    d = [global_sum(ld) for ld in local_descriptors]  # :23
    if force_descriptor_rows is not None:
        d = d + [np.asarray(r, dtype=np.float64) for r in force_descriptor_rows]
    if pca.m is None:
        s = d[0].copy()
        for v in d[1:]:
            s = s + v
        pca.m = s / len(d)  # :35
    dm = [di - pca.m for di in d]  # :37
    pca.lam, pca.W = select_eigendirections(dm, pca.tol)
    return pca


def centered_cov(X, mean=None):
    """(mean, (1/n) sum (x-m)(x-m)') with m = sum(x)/n when not supplied; vectorised (used at larger sizes)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    if mean is None:
        mean = X.sum(axis=0) / n
    Xc = X - mean
    return mean, (Xc.T @ Xc) / n


# ----------------------------------------------------------------------------------------------------------------
# KernelSteinDiscrepancy (src/Kernels/divergences.jl:27-49) with the RBF gradient kernels (kernels.jl:81-126, distances.jl:76-110)
# ----------------------------------------------------------------------------------------------------------------
def _cinv_matrix(e: Euclidean) -> np.ndarray:
    if e.cinv is None:
        return np.eye(e.dim)
    return np.diag(e.cinv) if np.ndim(e.cinv) == 1 else np.asarray(e.cinv, dtype=np.float64)


def compute_divergence_ksd(x, score, k: RBF) -> float:
    """compute_divergence(x, ::KernelSteinDiscrepancy) restated literally: the i <= j pair loop, every gradient function as written
    (compute_gradx/grady/gradxy_distance: 2*Cinv*(A-B), -2*Cinv*(A-B), -2*Cinv; compute_gradx/grady_kernel: -0.5*k*grad_d/l^2;
    compute_gradxy_kernel: k .* (-0.5*gradxy_d/l^2 .+ 0.25*gradx_d'*grady_d/l^4) — the `.+` broadcasts the scalar over the matrix)."""
    xs = [np.asarray(v, dtype=np.float64) for v in x]
    N = len(xs)
    sq = [np.asarray(score(v), dtype=np.float64) for v in xs]
    C = _cinv_matrix(k.d)
    ksd = 0.0
    for i in range(N):
        for j in range(i, N):
            m = 1 if i == j else 2
            kk = compute_kernel(xs[i], xs[j], k)
            gxd = 2 * C @ (xs[i] - xs[j])
            gyd = -2 * C @ (xs[i] - xs[j])
            gxyd = -2 * C
            grady_k = -0.5 * kk * gyd / k.ell ** 2
            gradx_k = -0.5 * kk * gxd / k.ell ** 2
            gradxy_k = kk * (-0.5 * gxyd / k.ell ** 2 + 0.25 * (gxd @ gyd) / k.ell ** 4)
            sks = sq[i] @ (kk * sq[j])
            sk = sq[i] @ grady_k
            ks = gradx_k @ sq[j]
            trk = np.trace(gradxy_k)
            ksd += m * (sks + sk + ks + trk) / (N * (N - 1.0))
    return float(ksd)


def select_eigenvectors_kdpp(lam, k, uniforms):
    """First phase of a k-DPP draw — which eigenvectors (Kulesza & Taskar 2012, Alg. 7 + Alg. 8; what Determinantal.sample(L, k), called
    at src/SubsetSelection/dpp.jl:51, does before its projection-DPP phase; external dependency, restated from the published algorithm).
    Elementary symmetric polynomials e_l(lam_1..lam_n) in exact rational-free form: scaled floating point per column would do, here
    they are kept as Python floats in the log domain with math.log / log1p (independent of the product's numpy / CUDA forms).
    uniforms[i] decides eigenvalue n = N - i (visited from the last one down). Returns the selected 0-based indices, descending."""
    import math

    lam = [float(x) for x in lam]
    N = len(lam)
    NEG = float("-inf")
    loglam = [math.log(x) if x > 0.0 else NEG for x in lam]

    def lae(a, b):
        mx = max(a, b)
        return mx if mx == NEG else mx + math.log1p(math.exp(-abs(a - b)))

    E = [[NEG] * (N + 1) for _ in range(k + 1)]
    E[0] = [0.0] * (N + 1)
    for n in range(1, N + 1):
        for l in range(1, min(n, k) + 1):
            E[l][n] = lae(E[l][n - 1], loglam[n - 1] + E[l - 1][n - 1])
    J, l = [], k
    for n in range(N, 0, -1):
        if l == 0:
            break
        if l == n:
            J.extend(range(n - 1, -1, -1))
            break
        logp = loglam[n - 1] + E[l - 1][n - 1] - E[l][n]
        u = uniforms[N - n]
        if (math.log(u) if u > 0.0 else NEG) < logp:
            J.append(n - 1)
            l -= 1
    return J


def sample_projection_dpp(V, uniforms):
    """Projection-DPP sampling over the orthonormal columns of V (n x m) with caller-supplied uniforms, one per step — the O(n m^2)
    Gram-Schmidt form of Kulesza & Taskar (2012) Alg. 1 that Determinantal.jl's sample_pdpp implements (Determinantal.jl is an external
    dependency of the reference, version unpinned; restated from the published algorithm). Inverse-CDF item choice: the first j with
    cumsum(p)[j] > u * sum(p). Returns the items in draw order."""
    V = np.asarray(V, dtype=np.float64)
    n, m = V.shape
    p = np.sum(V * V, axis=1)
    F = np.zeros((m, 0))
    items = []
    for i in range(m):
        pc = np.maximum(p, 0.0)
        cs = np.cumsum(pc)
        target = uniforms[i] * cs[-1]
        cand = np.nonzero((cs > target) & (pc > 0.0))[0]
        item = int(cand[0]) if len(cand) else int(np.nonzero(pc > 0.0)[0][-1])
        items.append(item)
        f = V[item, :].copy()
        for _ in range(2):
            f = f - F @ (F.T @ f)
        f = f / np.sqrt(f @ f)
        F = np.hstack([F, f[:, None]])
        t = V @ f
        p = p - t * t
        p[p < 0.0] = 0.0
        p[item] = 0.0
    return items


def greedy_map(L, k):
    """Greedy MAP of an L-ensemble (Determinantal.greedy_subset, called at src/SubsetSelection/dpp.jl:60; external dependency, restated
    from the published fast greedy algorithm of Chen, Zhang & Zhou 2018): at each step the item with the largest conditional variance."""
    L = np.asarray(L, dtype=np.float64)
    n = L.shape[0]
    d = np.diag(L).copy()
    C = np.zeros((k, n))
    chosen = []
    for t in range(k):
        j = int(np.argmax(d))
        if d[j] <= 0:
            break
        chosen.append(j)
        e = (L[j, :] - C[:t, j] @ C[:t, :]) / np.sqrt(d[j])
        C[t, :] = e
        d = d - e * e
        d[chosen] = -np.inf
    return chosen


# ----------------------------------------------------------------------------------------------------------------
# DBSCAN's distance matrices (src/SubsetSelection/dbscan.jl:100-170, kabsch.jl)
# ----------------------------------------------------------------------------------------------------------------
def periodic_rmsd(p1, p2, box_lengths) -> float:
    """periodic_rmsd (dbscan.jl:110-124), operation by operation: d = p1[i,:] - p2[i,:]; d .- round.(d ./ L) .* L (round half to even,
    like Julia's round); sum += norm(d)^2 with norm = sqrt of the sequential sum of squares; sqrt(sum / n_atoms)."""
    p1, p2, L = np.asarray(p1, dtype=np.float64), np.asarray(p2, dtype=np.float64), np.asarray(box_lengths, dtype=np.float64)
    total = 0.0
    for i in range(p1.shape[0]):
        d = p1[i, :] - p2[i, :]
        d = d - np.rint(d / L) * L
        ss = d[0] * d[0]
        ss = ss + d[1] * d[1]
        ss = ss + d[2] * d[2]
        nr = np.sqrt(ss)
        total = total + nr * nr
    return float(np.sqrt(total / p1.shape[0]))


def kabsch_rmsd(P, Q) -> float:
    """kabsch_rmsd(P, Q) = rmsd(P, kabsch(P, Q)) (kabsch.jl:91-124): centre both, A = Qc' * Pc, (u, d, v) = svd(A),
    f = sign(det(v) * det(u)), rotated = Qc * u * diag(1, 1, f) * v' + centroid_P, sqrt(sum((P - rotated).^2) / natoms)."""
    P, Q = np.asarray(P, dtype=np.float64), np.asarray(Q, dtype=np.float64)
    cp, cq = P.sum(axis=0) / P.shape[0], Q.sum(axis=0) / Q.shape[0]
    Pc, Qc = P - cp, Q - cq
    A = Qc.T @ Pc
    u, _, vt = np.linalg.svd(A)
    v = vt.T
    f = int(np.sign(np.linalg.det(v) * np.linalg.det(u)))
    rotated = Qc @ u @ np.diag([1.0, 1.0, float(f)]) @ v.T + cp
    return float(np.sqrt(np.sum((P - rotated) ** 2) / P.shape[0]))


def distance_matrix(positions, box_lengths=None) -> np.ndarray:
    """distance_matrix_periodic (dbscan.jl:132-147) if box_lengths is given, else distance_matrix_kabsch (:156-170)."""
    n = len(positions)
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            D[i, j] = periodic_rmsd(positions[i], positions[j], box_lengths) if box_lengths is not None else kabsch_rmsd(positions[i], positions[j])
            D[j, i] = D[i, j]
    return D

