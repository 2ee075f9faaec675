"""Host mirror of reference src/DimensionReduction (dimension_reduction.jl, pca.jl, pca_state.jl, as.jl) for the hot path:
the covariance / second-moment build of compute_eigen and the mean/centring of fit!(ds, ::PCAState) run on the GPU
(pl_cov); eigen, reordering and direction selection stay on the host as in the reference.

Julia `fit!` / `transform!` are spelled `fit_` / `transform_`.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .data import Configuration, DataSet, ForceDescriptors, LocalDescriptors, get_force_descriptors, get_local_descriptors, get_values


class DimensionReducer:
    pass


def covariance(rows, mean=None, center: bool = False):
    """GPU build of Q = (1/n) sum_r (x_r - m)(x_r - m)' (dimension_reduction.jl:12 with pca_state.jl:34-37).
    center=False: m = 0 (PCA / ActiveSubspace). center=True: m = `mean` if given else sum(x)/n. Returns (Q, m)."""
    X = _lib.as_f64(rows)
    if X.ndim != 2:
        raise ValueError("rows must be n vectors of equal length")
    n, K = X.shape
    C = np.empty((K, K))
    m = None
    given = 0
    if center:
        if mean is not None and len(mean) > 0:
            m = _lib.as_f64(mean).copy()
            given = 1
        else:
            m = np.zeros(K)
    _lib.check(_lib.lib().pl_cov(_lib.dptr(X), n, K, K, _lib.dptr(m), given, 1 if center else 0, _lib.dptr(C)), "pl_cov")
    return C, m


def compute_eigen(d, _precentered_mean=None):
    """compute_eigen(d) (dimension_reduction.jl:11-14): Q = Symmetric(mean(di*di')), eigen(Symmetric(Q)) -> (λ ascending, ϕ)."""
    Q, _ = covariance(d)
    return eigen_symmetric(Q)


def eigen_symmetric(Q):
    """eigen(Symmetric(Q)) (dimension_reduction.jl:13). LAPACK on the host by default, exactly as the reference (signs and the basis of
    degenerate clusters are solver-specific, so the default must not change solver); PLB200_EIGEN=gpu opts into pl_eigh (own solver up to n = 2048, cuSOLVER above)."""
    import os

    if os.environ.get("PLB200_EIGEN", "lapack") == "gpu":
        from .dpp import eigh

        return eigh(Q)
    lam, phi = np.linalg.eigh(np.asarray(Q, dtype=np.float64), UPLO="U")  # Q is exactly symmetric by construction
    return lam, phi


def _select(lam, phi, tol):
    lam, phi = lam[::-1], phi[:, ::-1]  # reorder (dimension_reduction.jl:17, 24)
    Σ = 1.0 - np.cumsum(lam) / np.sum(lam)
    if isinstance(tol, (int, np.integer)) and not isinstance(tol, (bool, np.bool_)):
        W = phi[:, :tol]  # :26
    else:
        W = phi[:, Σ > tol]  # :19
    return lam, W


def select_eigendirections(d, tol):
    """select_eigendirections(d, tol::Float64 | tol::Int) (dimension_reduction.jl:15-28)."""
    lam, phi = compute_eigen(d)
    return _select(lam, phi, tol)


class PCA(DimensionReducer):
    """PCA(; tol = 0.01) (pca.jl:10-16)."""

    def __init__(self, tol=0.01):
        self.tol = tol


class ActiveSubspace(DimensionReducer):
    """ActiveSubspace(Q, ∇Q; tol = 0.01) (as.jl:10-18)."""

    def __init__(self, Q, gradQ, tol=0.01):
        self.Q, self.gradQ, self.tol = Q, gradQ, tol


class PCAState(DimensionReducer):
    """mutable struct PCAState (pca_state.jl:9-18): tol, λ, W, m; m == [] until the first fit!."""

    def __init__(self, tol=0.01, λ=None, W=None, m=None):
        self.tol = tol
        self.λ = [] if λ is None else λ
        self.W = [] if W is None else W
        self.m = [] if m is None else m


def _global_sums(ds: DataSet):
    try:
        from .kernels import global_features

        return global_features([get_values(get_local_descriptors(c)) for c in ds], mean=False)  # sum.(get_values.(...)) on the GPU
    except KeyError:
        raise RuntimeError("No local descriptors found in DataSet") from None


def fit(ds: DataSet, dr: DimensionReducer):
    """fit(ds, ::PCA) (pca.jl:22-30) and fit(ds, ::ActiveSubspace) (as.jl:24-27): UN-centred second moment."""
    if isinstance(dr, PCA):
        d = _global_sums(ds)
    elif isinstance(dr, ActiveSubspace):
        d = [np.asarray(dr.gradQ(c), dtype=np.float64) for c in ds]
    else:
        raise TypeError(type(dr))
    return select_eigendirections(d, dr.tol)


def fit_(ds: DataSet, pca: PCAState, use_force_descriptors: bool = False):
    """fit!(ds, pca::PCAState) (pca_state.jl:20-40).

    As written, the reference's force-descriptor branch always throws (undefined `fd`, pca_state.jl:29) and is swallowed,
    so only the per-configuration global-sum descriptors are used; that is the default here. use_force_descriptors=True
    opts into the evidently intended behaviour (force-descriptor rows stacked under the descriptors — BASELINE config 4)."""
    d = _global_sums(ds)
    if use_force_descriptors:
        ff = [get_values(get_force_descriptors(c)).reshape(-1, d[0].shape[0]) for c in ds]
        rows = np.concatenate([np.stack(d, axis=0)] + ff, axis=0)
    else:
        rows = np.stack(d, axis=0)
    mean_given = len(pca.m) > 0  # `if pca.m == []` (pca_state.jl:34)
    Q, m = covariance(rows, mean=pca.m if mean_given else None, center=True)
    if not mean_given:
        pca.m = m
    lam, phi = eigen_symmetric(Q)  # LAPACK like the reference; PLB200_EIGEN=gpu opts into pl_eigh
    pca.λ, pca.W = _select(lam, phi, pca.tol)
    return None


def project(X, W, shift=None):
    """(X - shift) · W on the GPU: X rows x K, W K x p. Runs on the LINEAR epilogue of the Gram engine (X · (Wᵀ)ᵀ); the shift is
    applied as X·W - shift·W, exact to ~|shift|/σ ulps. The 128-wide output tile is wasteful for small p — the dedicated
    skinny-N tile is future work — but the rows never leave the device path."""
    from .kernels import KernelMatrix, _Linear

    X = _lib.as_f64(X)
    Wt = np.ascontiguousarray(np.asarray(W, dtype=np.float64).T)
    Y = KernelMatrix(X, Wt, _Linear())
    if shift is not None:
        Y = Y - np.asarray(shift, dtype=np.float64) @ np.asarray(W, dtype=np.float64)
    return Y


def transform_(ds: DataSet, dr: PCAState):
    """transform!(ds, dr) (pca_state.jl:42-62): W'(l - m/natoms_first) per atom, W'(fc - m) per force-descriptor row
    (SURVEY.md 8f row f3). The projections run on the GPU (`project`); the per-configuration re-wrapping stays on the host."""
    W = np.asarray(dr.W)
    m = np.asarray(dr.m)
    new = []
    try:
        ldc = [get_values(get_local_descriptors(c)) for c in ds]
        ml = m / ldc[0].shape[0]  # pca_state.jl:45: length(ldc[1]) = natoms of the FIRST configuration
        counts = [l.shape[0] for l in ldc]
        Y = project(np.concatenate(ldc, axis=0), W, ml)
        off = 0
        for c, n in zip(ds, counts):
            new.append(c + Configuration(LocalDescriptors(Y[off:off + n])))
            off += n
    except KeyError:
        new = list(ds)
    try:
        fds = [get_values(get_force_descriptors(c)) for c in new]
        Yf = project(np.concatenate([fd.reshape(-1, fd.shape[2]) for fd in fds], axis=0), W, m)
        out, off = [], 0
        for c, fd in zip(new, fds):
            n = fd.shape[0] * 3
            out.append(c + Configuration(ForceDescriptors(Yf[off:off + n].reshape(fd.shape[0], 3, -1))))
            off += n
        new = out
    except KeyError:
        pass
    ds.Configurations[:] = new
    return None
