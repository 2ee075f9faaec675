"""Host mirror of reference src/SubsetSelection/dpp.jl: kDPP over an L-ensemble whose kernel matrix AND eigendecomposition are
built on the GPU.

Hot path: `K = KernelMatrix(...)` inside the constructors (dpp.jl:25, :40). Widened: `EllEnsemble(K)`'s dense
symmetric eigendecomposition (Determinantal.jl [external], dpp.jl:26, :41) runs on the device right behind the Gram kernel
(pl_ell_ensemble; the library's own eigensolver up to n = 2048, cuSOLVER above); U stays device-resident, inclusion probabilities are computed there and a k-DPP draw
downloads only the eigenvectors it selected.

What stays on the host is Determinantal.jl's sampling logic, restated so that a Python user of this mirror gets a working
selector (Kulesza & Taskar, "Determinantal point processes for machine learning", 2012: Alg. 1 spectral sampling, Alg. 7/8
k-DPP eigenvector selection via elementary symmetric polynomials; greedy MAP by Schur-complement updates). In a Julia deployment
those calls stay in Determinantal.jl (version unpinned — Project.toml has no compat bound, Manifest is git-ignored) and see the
GPU-built Symmetric K or the (λ, U) pair.
PARITY UNPINNED for these host routines: they use numpy's RNG, not Julia's, so indices are not comparable draw-for-draw.
"""
from __future__ import annotations

import os

import numpy as np

from .data import DataSet, LocalDescriptors
from . import _lib
from .kernels import Feature, Kernel, Symmetric, compute_features


class SubsetSelector:
    pass


def eigh(A, fill: str = "full", vectors: bool = True):
    """eigen(Symmetric(A)) on the GPU (pl_eigh): (λ ascending, U with U[:, k] = eigenvector k) like numpy.linalg.eigh."""
    A = np.asarray(A) if isinstance(A, Symmetric) else _lib.as_f64(A)
    A = np.ascontiguousarray(A)
    n = A.shape[0]
    if A.ndim != 2 or A.shape[1] != n:
        raise ValueError("eigh needs a square matrix")
    W = np.empty(n)
    V = np.empty((n, n)) if vectors else None
    fl = {"full": _lib.PL_FILL_FULL, "U": _lib.PL_FILL_RM_UPPER, "L": _lib.PL_FILL_RM_LOWER}[fill]
    _lib.check(_lib.lib().pl_eigh(_lib.dptr(A), n, n, fl, _lib.dptr(W), _lib.dptr(V), n), "pl_eigh")
    return (W, V.T) if vectors else W


_resident_owner = None  # the EllEnsemble whose eigenvectors currently live in the library (one at a time)


class EllEnsemble:
    """L-ensemble: L = U diag(λ) U', scaled by α (Determinantal.jl EllEnsemble + rescale!).

    EllEnsemble(K): eigendecomposition of a host kernel matrix on the GPU (pl_eigh).
    EllEnsemble.from_features(X, kernel): SURVEY.md 8f row f1 — kernel matrix AND eigendecomposition on the GPU (pl_ell_ensemble);
    U stays device-resident: inclusion probabilities are computed there and a k-DPP draw downloads only the eigenvectors it selected."""

    def __init__(self, K=None):
        self.α = 1.0
        self._U = None
        self._resident = False
        if K is None:
            return
        self.L = np.asarray(K) if isinstance(K, Symmetric) else np.asarray(K, dtype=np.float64)
        if self.L.shape[0] != self.L.shape[1]:
            raise ValueError("EllEnsemble needs a square kernel matrix")
        lam, U = eigh(self.L)
        self.λ = np.maximum(lam, 0.0)  # clip round-off negatives of a PSD kernel
        self._U = U

    @classmethod
    def from_eigen(cls, lam, U, L=None):
        """wrap an eigendecomposition computed elsewhere (what a Julia caller does when it hands Determinantal a precomputed (λ, U))"""
        self = cls()
        self.λ = np.maximum(np.asarray(lam, dtype=np.float64), 0.0)
        self._U = np.asarray(U, dtype=np.float64)
        if L is not None:
            self.L = np.asarray(L, dtype=np.float64)
        return self

    @classmethod
    def from_features(cls, X, k: Kernel, keep_L: bool = True):
        global _resident_owner
        from .kernels import _features_matrix, _for_flattened, _kargs

        F = X
        X = _features_matrix(F)
        n, d = X.shape
        kid, alpha, kind, cinv, ell_, beta = _kargs(_for_flattened(k, F, d), d)
        self = cls()
        if _resident_owner is not None and _resident_owner is not self:
            _resident_owner._evict()
        lam = np.empty(n)
        Kst = np.zeros((n, n)) if keep_L else None  # zeros(n, n) (kernels.jl:216)
        _lib.check(_lib.lib().pl_ell_ensemble(kid, _lib.dptr(X), n, d, d, alpha, kind, _lib.dptr(cinv), ell_, beta, _lib.dptr(lam), _lib.dptr(Kst), n),
                   "pl_ell_ensemble")
        self._Lsym = Symmetric(Kst.T, "U") if keep_L else None  # row-major lower == Julia's column-major uplo = :U
        self.λ = np.maximum(lam, 0.0)
        self._n = n
        self._resident = True
        _resident_owner = self
        return self

    # the kernel matrix (greedy_subset needs it); materialised from the Symmetric storage on first use
    @property
    def L(self):
        if "_Lfull" not in self.__dict__:
            if getattr(self, "_Lsym", None) is None:
                raise _lib.PLB200Error("this L-ensemble was built with keep_L=False")
            self._Lfull = np.asarray(self._Lsym)
        return self._Lfull

    @L.setter
    def L(self, v):
        self._Lfull = v

    @property
    def nitems(self):
        return len(self.λ)

    @property
    def U(self):
        """all eigenvectors as columns (downloaded from the device on first use for a resident ensemble)"""
        if self._U is None:
            self._U = self.eigvecs(np.arange(self.nitems))
        return self._U

    def eigvecs(self, J):
        """U[:, J] — for a resident ensemble only these len(J) eigenvectors cross PCIe (pl_ell_eigvecs)."""
        import ctypes

        J = np.ascontiguousarray(J, dtype=np.int64)
        if self._U is not None:
            return self._U[:, J].copy()
        if not self._resident:
            raise _lib.PLB200Error("the device-resident eigenvectors of this L-ensemble were replaced by a newer ensemble; rebuild it")
        V = np.empty((len(J), self._n))
        _lib.check(_lib.lib().pl_ell_eigvecs(J.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(J), self._n, _lib.dptr(V), self._n), "pl_ell_eigvecs")
        return np.ascontiguousarray(V.T)

    def _evict(self):
        """another ensemble takes the device slot: keep small ones usable by downloading U first"""
        global _resident_owner
        if self._resident and self._U is None and self.nitems <= 4096:
            self._U = self.eigvecs(np.arange(self.nitems))
        self._resident = False
        _resident_owner = None


def workspace_trim(keep_resident: bool = True):
    """pl_workspace_trim: hand the library's device scratch back (it only grows otherwise). keep_resident=False also drops the resident
    L-ensemble (small ones are downloaded first, like any eviction) and an open normal-equation accumulator."""
    if not keep_resident and _resident_owner is not None:
        _resident_owner._evict()
    _lib.check(_lib.lib().pl_workspace_trim(1 if keep_resident else 0), "pl_workspace_trim")


def rescale_(ell: EllEnsemble, k: int):
    """rescale!(L, k): pick α so that the expected cardinality sum(αλ/(1+αλ)) equals k (bisection in log α)."""
    lam = ell.λ
    rank = int(np.sum(lam > lam.max() * 1e-14))
    if not (0 < k <= rank):
        raise ValueError(f"cannot rescale an L-ensemble of numerical rank {rank} to expected size {k}")
    if k == rank:
        ell.α = 1e12 / max(lam.max(), 1e-300)
        return ell
    f = lambda la: np.sum(np.exp(la) * lam / (1.0 + np.exp(la) * lam)) - k
    lo, hi = -60.0, 60.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if f(mid) > 0:
            hi = mid
        else:
            lo = mid
    ell.α = float(np.exp(0.5 * (lo + hi)))
    return ell


def _log_esp(lam, k):
    """logE[l, n] = log e_l(λ_1..λ_n), elementary symmetric polynomials (Kulesza & Taskar Alg. 7) in the log domain:
    with hundreds of items e_k spans hundreds of orders of magnitude and under/overflows in plain floating point."""
    N = len(lam)
    with np.errstate(divide="ignore"):
        loglam = np.log(lam)
    E = np.full((k + 1, N + 1), -np.inf)
    E[0, :] = 0.0
    for n in range(1, N + 1):
        E[1:, n] = np.logaddexp(E[1:, n - 1], loglam[n - 1] + E[:-1, n - 1])
    return E, loglam


def _esp_select_host(lam, k, u1):
    """Kulesza & Taskar Alg. 8 on the host (ensembles that are not device-resident)."""
    N = len(lam)
    E, loglam = _log_esp(lam, k)
    J = []
    l = k
    with np.errstate(divide="ignore"):
        logu = np.log(u1)
    for n in range(N, 0, -1):
        if l == 0:
            break
        if l == n:  # every remaining eigenvector is needed
            J.extend(range(n - 1, -1, -1))
            l = 0
            break
        logp = loglam[n - 1] + E[l - 1, n - 1] - E[l, n]
        if logu[N - n] < logp:
            J.append(n - 1)
            l -= 1
    return J


def _esp_select_device(lam, k, u1):
    """the same selection by pl_esp_select: the (N+1) x (k+1) table is built and walked on the device"""
    import ctypes

    lam = np.ascontiguousarray(lam, dtype=np.float64)
    u1 = np.ascontiguousarray(u1, dtype=np.float64)
    idx = np.empty(max(k, 1), dtype=np.int64)
    cnt = ctypes.c_int64(0)
    _lib.check(_lib.lib().pl_esp_select(_lib.dptr(lam), len(lam), int(k), _lib.dptr(u1), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                        ctypes.byref(cnt)), "pl_esp_select")
    return [int(i) for i in idx[: cnt.value]]


def sample(ell: EllEnsemble, k: int, rng=None):
    """Determinantal.sample(L, k): exact k-DPP sample (Kulesza & Taskar Alg. 8 + Alg. 1). Returns sorted 0-based indices.
    For a device-resident ensemble both phases run on the GPU: phase 1 (which eigenvectors: elementary symmetric polynomials, O(N k),
    pl_esp_select) and phase 2 (projection-DPP sampling over them, pl_ell_sample); the uniforms of both come from `rng`. Ensembles that
    are not resident (built from host eigenpairs) use the numpy forms of the same two phases."""
    rng = np.random.default_rng() if rng is None else rng
    lam = np.maximum(ell.α * ell.λ, 0.0)
    N = len(lam)
    u1 = rng.random(N)  # one uniform per eigenvalue, visited from the last one down (u1[N - n] decides eigenvalue n)
    on_device = ell._resident and _resident_owner is ell
    if on_device and not os.environ.get("PLB200_ESP_HOST") and (N + 1) * (k + 2) * 8 <= 24 * 2**30:
        J = _esp_select_device(lam, k, u1)
    else:
        J = _esp_select_host(lam, k, u1)
    if on_device and len(J) > 0:
        # second phase on the device (pl_ell_sample): U stays where it is, the uniforms come from this generator
        import ctypes

        Jarr = np.ascontiguousarray(J, dtype=np.int64)
        u = np.ascontiguousarray(rng.random(len(J)))
        out = np.empty(len(J), dtype=np.int64)
        i64p = ctypes.POINTER(ctypes.c_int64)
        _lib.check(_lib.lib().pl_ell_sample(Jarr.ctypes.data_as(i64p), len(J), N, _lib.dptr(u), out.ctypes.data_as(i64p)), "pl_ell_sample")
        return sorted(int(i) for i in out)
    V = ell.eigvecs(J)
    items = []
    for _ in range(len(J)):
        p = np.sum(V * V, axis=1)
        p = np.maximum(p, 0.0)
        i = int(rng.choice(N, p=p / p.sum()))
        items.append(i)
        j = int(np.argmax(np.abs(V[i, :])))  # eliminate e_i from span(V)
        Vj = V[:, j].copy()
        V = np.delete(V, j, axis=1)
        if V.shape[1] == 0:
            break
        V = V - np.outer(Vj, V[i, :] / Vj[i])
        V, _ = np.linalg.qr(V)
    return sorted(items)


def greedy_subset(ell: EllEnsemble, k: int):
    """greedy_subset(L, k): greedy MAP (largest conditional variance first, Cholesky-style updates)."""
    if ell._resident and _resident_owner is ell and ell.nitems <= 32768:
        import ctypes

        out = np.empty(k, dtype=np.int64)
        _lib.check(_lib.lib().pl_ell_greedy(int(k), ell.nitems, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))), "pl_ell_greedy")
        return [int(i) for i in out if i >= 0]
    L = ell.L
    N = L.shape[0]
    d = np.diag(L).astype(np.float64).copy()
    C = np.zeros((k, N))
    chosen = []
    for t in range(k):
        i = int(np.argmax(d))
        chosen.append(i)
        if d[i] <= 0:
            break
        e = (L[i, :] - C[:t, i] @ C[:t, :]) / np.sqrt(d[i])
        C[t, :] = e
        d = d - e * e
        d[chosen] = -np.inf
    return chosen


def inclusion_prob(ell: EllEnsemble):
    """Determinantal.inclusion_prob(L): diagonal of the marginal kernel U diag(αλ/(1+αλ)) U'."""
    if ell._resident and _resident_owner is ell:
        p = np.empty(ell.nitems)
        _lib.check(_lib.lib().pl_ell_inclusion_prob(float(ell.α), ell.nitems, _lib.dptr(p)), "pl_ell_inclusion_prob")
        return p
    g = ell.α * ell.λ
    g = g / (1.0 + g)
    return np.sum(ell.U * ell.U * g[None, :], axis=1)


class kDPP(SubsetSelector):
    """kDPP(ds, f, k; batch_size = length(ds) ÷ 2, dt) and kDPP(features, k; batch_size) (dpp.jl:9-44).
    `K` is the (rescaled) EllEnsemble as in the reference; `dpp.K.L` is the GPU-built kernel matrix."""

    def __init__(self, *args, batch_size=None, dt=LocalDescriptors):
        if len(args) == 3 and isinstance(args[0], DataSet) and isinstance(args[1], Feature) and isinstance(args[2], Kernel):
            ds, f, k = args
            features = compute_features(ds, f, dt=dt)  # kernels.jl:251
        elif len(args) == 2 and isinstance(args[1], Kernel):
            features, k = args
        else:
            raise TypeError("kDPP(ds, f, k; batch_size, dt) | kDPP(features, k; batch_size)")
        n = len(features)
        self.batch_size = n // 2 if batch_size is None else int(batch_size)
        # K = KernelMatrix(...) (dpp.jl:25, :40) and EllEnsemble(K) (:26, :41) fused on the GPU; K is kept as ell.L
        ell = EllEnsemble.from_features(features, k)
        rescale_(ell, self.batch_size)
        self.K = ell


def get_random_subset(dpp, batch_size: int | None = None, rng=None):
    """get_random_subset(dpp; batch_size) (dpp.jl:50-53); also dispatches on RandomSelector (random.jl:21-26) and DBSCANSelector
    (dbscan.jl:49-55) like the generic function of the reference."""
    if isinstance(dpp, RandomSelector):
        return _random_subset_random(dpp, batch_size, rng)
    if isinstance(dpp, DBSCANSelector):
        return _random_subset_dbscan(dpp, batch_size, rng)
    return sample(dpp.K, dpp.batch_size if batch_size is None else batch_size, rng)


def get_dpp_mode(dpp: kDPP, batch_size: int | None = None):
    """get_dpp_mode(dpp; batch_size) (dpp.jl:59-62)."""
    return greedy_subset(dpp.K, dpp.batch_size if batch_size is None else batch_size)


def get_inclusion_prob(dpp: kDPP):
    """get_inclusion_prob(dpp) (dpp.jl:68-70)."""
    return np.asarray(inclusion_prob(dpp.K)).reshape(-1)


# ------------------------------------------------------------------ DBSCAN's distance matrices (dbscan.jl) --------
def _positions(positions) -> np.ndarray:
    P = np.ascontiguousarray(np.asarray(positions, dtype=np.float64))
    if P.ndim != 3 or P.shape[2] != 3:
        raise ValueError("positions must be n configurations x natoms x 3 (every configuration with the same number of atoms)")
    return P


def distance_matrix_periodic(positions, box_lengths) -> np.ndarray:
    """distance_matrix_periodic(ds) (dbscan.jl:132-147) on the GPU (pl_rmsd_periodic): periodic_rmsd of every pair of configurations in
    one orthorhombic box (the reference errors if the boxes differ). positions: n x natoms x 3."""
    P = _positions(positions)
    L = _lib.as_f64(box_lengths)
    if L.shape != (3,):
        raise ValueError("box_lengths must have 3 entries")
    n = P.shape[0]
    D = np.zeros((n, n))
    _lib.check(_lib.lib().pl_rmsd_periodic(_lib.dptr(P.reshape(-1)), n, P.shape[1], _lib.dptr(L), _lib.dptr(D), n), "pl_rmsd_periodic")
    return D


def distance_matrix_kabsch(positions) -> np.ndarray:
    """distance_matrix_kabsch(ds) (dbscan.jl:156-170) on the GPU (pl_rmsd_kabsch): RMSD after the optimal superposition."""
    P = _positions(positions)
    n = P.shape[0]
    D = np.zeros((n, n))
    _lib.check(_lib.lib().pl_rmsd_kabsch(_lib.dptr(P.reshape(-1)), n, P.shape[1], _lib.dptr(D), n), "pl_rmsd_kabsch")
    return D


class RandomSelector(SubsetSelector):
    """RandomSelector(num_configs; batch_size) (random.jl:1-15): a permutation sampler over 1:num_configs."""

    def __init__(self, num_configs: int, batch_size=None):
        self.num_configs = int(num_configs)
        self.batch_size = self.num_configs if batch_size is None else int(batch_size)


def _random_subset_random(r: RandomSelector, batch_size=None, rng=None):
    """get_random_subset(r::RandomSelector, batch_size) (random.jl:21-26): randperm(num_configs)[1:batch_size] (0-based here)."""
    rng = np.random.default_rng() if rng is None else rng
    bs = r.batch_size if batch_size is None else int(batch_size)
    if not (bs <= r.num_configs):
        raise AssertionError("batch_size <= num_configs")
    return [int(i) for i in rng.permutation(r.num_configs)[:bs]]


class DBSCANSelector(SubsetSelector):
    """DBSCANSelector(ds, eps, minpts, sample_size) (dbscan.jl:14-38). The O(n^2) distance matrix between configurations
    (get_clusters, :83-98: periodic RMSD if the system is periodic, Kabsch RMSD otherwise) is built on the GPU; the clustering itself is
    Clustering.jl's `dbscan` in the reference ([external]) and scikit-learn's DBSCAN on the precomputed matrix here.
    positions: n x natoms x 3; box_lengths: the 3 box edges of a periodic system, or None."""

    def __init__(self, positions, eps, minpts, sample_size, box_lengths=None):
        from sklearn.cluster import DBSCAN

        D = distance_matrix_periodic(positions, box_lengths) if box_lengths is not None else distance_matrix_kabsch(positions)
        labels = DBSCAN(eps=float(eps), min_samples=int(minpts), metric="precomputed").fit(D).labels_
        self.clusters = [np.nonzero(labels == c)[0].tolist() for c in range(int(labels.max()) + 1)]  # noise (label -1) belongs to no cluster
        self.eps, self.minpts, self.sample_size = eps, minpts, int(sample_size)
        self.distances = D


def _random_subset_dbscan(s: DBSCANSelector, batch_size=None, rng=None):
    """get_random_subset(s::DBSCANSelector, batch_size) (dbscan.jl:49-55): batch_size ÷ length(clusters) samples WITH replacement from
    every cluster (`c[rand(1:length(c), batch_size)]`, :66-71)."""
    rng = np.random.default_rng() if rng is None else rng
    bs = s.sample_size if batch_size is None else int(batch_size)
    if not s.clusters:
        return []
    per = bs // len(s.clusters)
    out = []
    for c in s.clusters:
        out.extend(int(c[i]) for i in rng.integers(0, len(c), size=per))
    return out

