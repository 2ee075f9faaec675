"""Host mirror of reference src/Learning (linear-learning-problem.jl, linear-learn.jl) for the hot path:
LinearProblem construction (which defines the packed row layout) and every normal-equation assembly site; the
assembly runs on the GPU (pl_normal_eq), the small K x K solves stay on the host exactly as in the reference.

Julia `learn!` is spelled `learn_` here (trailing underscore = mutates its first argument).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .data import DataSet, get_energy, get_force_descriptors, get_forces, get_local_descriptors, get_values


class _LinearProblemMeta(type):
    def __call__(cls, *args, **kw):
        if cls is LinearProblem:
            return _linear_problem(*args, **kw)  # LinearProblem(ds) constructs the right concrete subtype
        return super().__call__(*args, **kw)


class LinearProblem(metaclass=_LinearProblemMeta):
    """abstract type LinearProblem (linear-learning-problem.jl:8); LinearProblem(ds) is the constructor function (:71-145)."""


class UnivariateLinearProblem(LinearProblem):
    """linear-learning-problem.jl:22-29."""

    def __init__(self, iv_data, dv_data, beta, beta0, sigma, Sigma):
        self.iv_data, self.dv_data = iv_data, dv_data
        self.β, self.β0, self.σ, self.Σ = beta, beta0, sigma, Sigma

    def __repr__(self):
        return f"UnivariateLinearProblem{{T, {self.β}, {self.σ}}}"


class CovariateLinearProblem(LinearProblem):
    """linear-learning-problem.jl:48-58."""

    def __init__(self, e, f, B, dB, beta, beta0, sigma_e, sigma_f, Sigma):
        self.e, self.f, self.B, self.dB = e, f, B, dB
        self.β, self.β0, self.σe, self.σf, self.Σ = beta, beta0, sigma_e, sigma_f, Sigma

    def __repr__(self):
        return f"CovariateLinearProblem{{T, {self.β}, {self.σe}, {self.σf}}}"


def _linear_problem(ds: DataSet):
    """LinearProblem(ds) (linear-learning-problem.jl:71-145): detect (energy, descriptors) / (forces, force descriptors)."""
    try:
        descriptors = [np.add.reduce(get_values(get_local_descriptors(c)), axis=0) for c in ds]  # :75 sum.(...)
        energies = [get_values(get_energy(c)) for c in ds]
        d_flag = True
    except KeyError:
        d_flag, descriptors, energies = False, None, None
    try:
        # :81 reduce(vcat, get_values(fd)) -> 3natoms vectors (atom-major, xyz inner); :127 reduce(hcat, fi) -> K x 3natoms.
        # Stored transposed (3natoms x K, row-major) — byte-identical to Julia's column-major K x 3natoms.
        force_descriptors = [get_values(get_force_descriptors(c)).reshape(-1, get_values(get_force_descriptors(c)).shape[2]) for c in ds]
        forces = [get_values(get_forces(c)).reshape(-1) for c in ds]  # :126 reduce(vcat, fi)
        fd_flag = True
    except KeyError:
        fd_flag, force_descriptors, forces = False, None, None
    if d_flag and not fd_flag:
        dim = len(descriptors[0])
        return UnivariateLinearProblem(descriptors, energies, np.zeros(dim), np.zeros(1), np.array([1.0]), np.zeros((dim, dim)))
    if fd_flag and not d_flag:
        dim = force_descriptors[0].shape[1]
        return UnivariateLinearProblem([m.T for m in force_descriptors], forces, np.zeros(dim), np.zeros(1), np.array([1.0]), np.zeros((dim, dim)))
    if d_flag and fd_flag:
        dim_d, dim_fd = len(descriptors[0]), force_descriptors[0].shape[1]
        if dim_d != dim_fd:
            raise ValueError("Descriptors and Force Descriptors have different dimension!")
        dim = dim_d
        return CovariateLinearProblem(energies, forces, descriptors, [m.T for m in force_descriptors], np.zeros(dim), np.zeros(1),
                                      np.array([1.0]), np.array([1.0]), np.zeros((dim, dim)))
    raise ValueError("Either no (Energy, Descriptors) or (Forces, Force Descriptors) in DataSet")


# ----------------------------------------------------------------------------------------------------------------
# packing: nested reference containers -> one row-major R x K buffer (the device layout), zero transposition
# ----------------------------------------------------------------------------------------------------------------
def _rows_of(iv):
    """rows contributed by one iv_data / B / dB entry: a K-vector is one row; a K x m matrix is m rows (its columns)."""
    iv = np.asarray(iv, dtype=np.float64)
    return iv[None, :] if iv.ndim == 1 else iv.T


def pack_rows(blocks) -> np.ndarray:
    return np.ascontiguousarray(np.concatenate([_rows_of(b) for b in blocks], axis=0))


def pack_rhs(vals) -> np.ndarray:
    return np.ascontiguousarray(np.concatenate([np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in vals]))


def normal_equations(A, b=None, w=None, n_energy_rows=0, w_e=1.0, w_f=1.0, intercept=False, lam=0.0):
    """GPU assembly of (A'WA + λI, A'Wb) through pl_normal_eq. A: R x K rows; see include/plb200.h for the weights."""
    A = _lib.as_f64(A)
    R, K = A.shape
    n = K + (1 if intercept else 0)
    b = None if b is None else _lib.as_f64(b)
    w = None if w is None else _lib.as_f64(w)
    M = np.empty((n, n))
    v = np.empty(n) if b is not None else None
    _lib.check(
        _lib.lib().pl_normal_eq(_lib.dptr(A), R, K, K, _lib.dptr(w), int(n_energy_rows), float(w_e), float(w_f), _lib.dptr(b),
                                1 if intercept else 0, float(lam), _lib.dptr(M), _lib.dptr(v)),
        "pl_normal_eq",
    )
    return M, v


def normal_equations_blocks(blocks, b=None, w=None, n_energy_rows=0, w_e=1.0, w_f=1.0, intercept=False, lam=0.0, chunk_rows=0, layout="julia"):
    """Streamed GPU assembly over a LIST of row blocks (pl_normal_eq_blocks): no host-side concatenation. Orientation follows the
    reference's containers and is decided by SHAPE, exactly as pack_rows does: a 1-D entry is a K-vector (one row: lp.B[i], iv_data[i]);
    a 2-D entry is K x m (lp.dB[i], linear-learning-problem.jl:127) and contributes its m columns as rows — a Fortran-ordered / transposed
    view (Julia's column-major block) is used as it lies in memory, a C-ordered K x m array is copied once. layout="rows" declares
    2-D entries to be rows x K instead (already packed row blocks)."""
    import ctypes

    keep, ptrs, rows = [], [], []
    K = None
    for blk in blocks:
        a = np.asarray(blk, dtype=np.float64)
        if a.ndim == 1:
            a = a[None, :]
        elif layout != "rows":
            a = a.T  # K x m -> m rows of K; free when the block is column-major like Julia's
        a = np.ascontiguousarray(a)
        if K is None:
            K = a.shape[1]
        if a.shape[1] != K:
            raise ValueError("DimensionMismatch: blocks of different descriptor length")
        keep.append(a)
        ptrs.append(a.ctypes.data)
        rows.append(a.shape[0])
    if K is None:
        raise ValueError("no blocks")
    nb = len(keep)
    PtrArr = ctypes.POINTER(ctypes.c_double) * nb
    parr = PtrArr(*[ctypes.cast(p, ctypes.POINTER(ctypes.c_double)) for p in ptrs])
    rarr = (ctypes.c_int64 * nb)(*rows)
    n = K + (1 if intercept else 0)
    b = None if b is None else _lib.as_f64(b)
    w = None if w is None else _lib.as_f64(w)
    M = np.empty((n, n))
    v = np.empty(n) if b is not None else None
    _lib.check(
        _lib.lib().pl_normal_eq_blocks(parr, rarr, nb, K, _lib.dptr(w), int(n_energy_rows), float(w_e), float(w_f), _lib.dptr(b),
                                       1 if intercept else 0, float(lam), int(chunk_rows), _lib.dptr(M), _lib.dptr(v)),
        "pl_normal_eq_blocks",
    )
    return M, v


class NormalEquationAccumulator:
    """Out-of-core accumulation of (A'WA, A'Wb) on the GPU (pl_neq_begin / pl_neq_add / pl_neq_finish): what ooc_learn! does with
    `AtWA .+= A'*W*A; AtWb .+= A'*W*b` per configuration (linear-learn.jl:355-356), without the rows ever existing as one matrix.
    Rows are buffered on the host up to `batch_rows` and shipped as one batch (a per-configuration launch would be latency-bound)."""

    def __init__(self, K: int, batch_rows: int = 1 << 16):
        self.K = int(K)
        self.batch_rows = int(batch_rows)
        self._rows, self._rhs, self._w, self._n = [], [], [], 0
        _lib.check(_lib.lib().pl_neq_begin(self.K), "pl_neq_begin")

    def add(self, A, b, w):
        """rows A (m x K), right-hand sides b (m), per-row weights w (m)"""
        A = np.asarray(A, dtype=np.float64).reshape(-1, self.K)
        self._rows.append(A)
        self._rhs.append(np.asarray(b, dtype=np.float64).reshape(-1))
        self._w.append(np.asarray(w, dtype=np.float64).reshape(-1))
        self._n += A.shape[0]
        if self._n >= self.batch_rows:
            self.flush()

    def flush(self):
        if self._n == 0:
            return
        A = np.ascontiguousarray(np.concatenate(self._rows, axis=0))
        b = np.ascontiguousarray(np.concatenate(self._rhs))
        w = np.ascontiguousarray(np.concatenate(self._w))
        _lib.check(_lib.lib().pl_neq_add(_lib.dptr(A), A.shape[0], self.K, _lib.dptr(w), 0, 1.0, 1.0, _lib.dptr(b)), "pl_neq_add")
        self._rows, self._rhs, self._w, self._n = [], [], [], 0

    def result(self, intercept: bool = False, lam: float = 0.0):
        self.flush()
        n = self.K + (1 if intercept else 0)
        M, v = np.empty((n, n)), np.empty(n)
        _lib.check(_lib.lib().pl_neq_finish(1 if intercept else 0, float(lam), _lib.dptr(M), _lib.dptr(v)), "pl_neq_finish")
        return M, v


def _solve(M, v):
    """βs = M \\ v, falling back to pinv(M)*v on a singular system (linear-learn.jl:199-205, 252-258)."""
    try:
        return np.linalg.solve(M, v)
    except np.linalg.LinAlgError as e:
        print(e)
        print("Linear system will be solved using pinv.")
        return np.linalg.pinv(M) @ v


def assemble_matrices(lp: CovariateLinearProblem, ws):
    """assemble_matrices(lp, ws) (linear-learn.jl:286-299): (A, W, b) with W returned as its diagonal."""
    A = np.concatenate([pack_rows(lp.B), pack_rows(lp.dB)], axis=0)
    b = np.concatenate([pack_rhs(lp.e), pack_rhs(lp.f)])
    W = np.concatenate([ws[0] * np.ones(len(lp.e)), ws[1] * np.ones(A.shape[0] - len(lp.e))])
    return A, W, b


def learn_(lp: LinearProblem, *args, **kw):
    """learn!(lp, ...) dispatch (linear-learn.jl):
    learn_(lp)                    default WLS, ws = ones, int = false          (:278-284)
    learn_(lp, ws::list, int)     weighted least squares [; λ for covariate]   (:178-215, :226-268)
    learn_(lp, α::Real)           OLS / MLE assembly                            (:11-23, :37-69)
    """
    if len(args) == 0:
        n = 1 if isinstance(lp, UnivariateLinearProblem) else 2
        return learn_(lp, [1.0] * n, False)
    if len(args) == 2 and isinstance(args[1], (bool, np.bool_)):
        return _learn_wls(lp, list(args[0]), bool(args[1]), **kw)
    if len(args) == 1 and np.isscalar(args[0]):
        return _learn_ols(lp, float(args[0]))
    raise TypeError("learn_(lp) | learn_(lp, ws, int; λ) | learn_(lp, α); the minibatch/Adam variants stay in PotentialLearning.jl")


def _learn_wls(lp, ws, intercept: bool, λ: float = 0.0, lam: float | None = None):
    if lam is not None:
        λ = lam
    if isinstance(lp, UnivariateLinearProblem):
        if λ != 0.0:
            raise TypeError("learn!(::UnivariateLinearProblem, ws, int) has no λ keyword (linear-learn.jl:178-182)")
        A = pack_rows(lp.iv_data)
        b = pack_rhs(lp.dv_data)
        # Q = Diagonal(ws[1] * ones(length(e_train))) (:196); the intercept column is all ones (:189-190)
        M, v = normal_equations(A, b, None, A.shape[0], ws[0], ws[0], intercept, 0.0)
    else:
        # A = [B; dB] (linear-learn.jl:238-244) is never materialised on the host: lp.B / lp.dB[i] are handed to the library as they lie
        b = np.concatenate([pack_rhs(lp.e), pack_rhs(lp.f)])
        M, v = normal_equations_blocks(list(lp.B) + list(lp.dB), b, None, len(lp.e), ws[0], ws[1], intercept, λ)
    βs = _solve(M, v)
    if intercept:
        lp.β0[:] = βs[0]
        lp.β[:] = βs[1:]
    else:
        lp.β[:] = βs
    return lp


def ols_sums(iv_data, dv_data):
    """AtA = sum(v*v' for v in iv), Atb = sum(v*b for (v,b) in zip(iv,dv)) (linear-learn.jl:16-17, 45-49, 94-95, 139-143)."""
    A = pack_rows(iv_data)
    b = pack_rhs(dv_data)
    return normal_equations(A, b)


def batch_sums(lp: LinearProblem, inds):
    """The per-step assembly of the minibatch learn!(lp, ss::SubsetSelector, α; …) variants (linear-learn.jl:92-95 univariate,
    :137-143 covariate): sums over the configurations `inds = get_random_subset(ss)` only. The selected blocks are streamed to the GPU
    as they lie (pl_normal_eq_blocks); the Flux / Adam / MvNormal optimisation around these sums stays in PotentialLearning.jl.
    Returns (AtA, Atb) resp. (AtAe, Atbe, AtAf, Atbf). The final full-data sums (:107, :157-158) are ols_sums(...)."""
    inds = [int(i) for i in inds]
    if isinstance(lp, UnivariateLinearProblem):
        return normal_equations_blocks([lp.iv_data[i] for i in inds], pack_rhs([lp.dv_data[i] for i in inds]))
    AtAe, Atbe = normal_equations_blocks([lp.B[i] for i in inds], pack_rhs([lp.e[i] for i in inds]))
    AtAf, Atbf = normal_equations_blocks([lp.dB[i] for i in inds], pack_rhs([lp.f[i] for i in inds]))
    return AtAe, Atbe, AtAf, Atbf


def _learn_ols(lp, α: float):
    if isinstance(lp, UnivariateLinearProblem):
        AtA, Atb = ols_sums(lp.iv_data, lp.dv_data)  # :16-17 on the GPU
        Q = _pinv_rtol(AtA, α)  # pinv(AtA, α) == pinv(AtA; rtol = α) (:19)
        lp.β[:] = Q @ Atb
        lp.σ[:] = np.std(Atb - AtA @ lp.β, ddof=1)
        lp.Σ[:, :] = lp.σ[0] ** 2 * Q
        return lp
    # covariate MLE (:37-69): only the four sums are on the hot path; the Adam/MvNormal optimisation stays in Julia
    AtAe, Atbe = ols_sums(lp.B, lp.e)
    AtAf, Atbf = ols_sums(lp.dB, lp.f)
    return AtAe, Atbe, AtAf, Atbf


def _pinv_rtol(M, rtol):
    """Julia's positional `pinv(M, tol)` (linear-learn.jl:19, :108, :61-66) is `pinv(M; rtol = tol)` [LinearAlgebra, external]: singular values
    s with s <= rtol * maximum(s) are dropped (strict `>` keeps the rest); atol = 0. rtol = 0 keeps every non-zero singular value."""
    U, s, Vt = np.linalg.svd(M)
    tol = float(rtol) * (s.max() if s.size else 0.0)
    keep = s > tol
    sinv = np.where(keep, 1.0 / np.where(keep, s, 1.0), 0.0)
    return (Vt.T * sinv) @ U.T


def ooc_learn_(lb_beta: np.ndarray, configs, ws=(30.0, 1.0), symmetrize=True, λ=0.01, reg_style="default", AtWA=None, AtWb=None,
               eweight_normalized="squared", batch_rows: int = 1 << 16):
    """ooc_learn!(lb, ds_train; ...) (linear-learn.jl:301-391) with the per-configuration accumulation
    AtWA .+= A'*W*A; AtWb .+= A'*W*b (:355-356) accumulated OUT OF CORE on the GPU (NormalEquationAccumulator: batches of
    `batch_rows` rows, per-row weights — the energy weight ws[1]*{1, 1/natoms, 1/natoms^2} differs per configuration, :341-352);
    `configs` may be a generator: nothing but the running sums is kept.
    configs: iterable of (global_descr K, force_descrs 3natoms x K, ref_energy, ref_forces 3natoms); descriptor evaluation
    itself (InteratomicPotentials) is upstream of the hot path. Returns (AtWA, AtWb) like the reference."""
    if AtWA is None or AtWb is None:
        acc = None
        for gd, fdm, e, f in configs:  # one configuration at a time, as the reference: nothing is kept but the running sums
            fdm = np.asarray(fdm, dtype=np.float64)
            natoms = fdm.shape[0] // 3
            if eweight_normalized is None:
                we_norm = 1.0
            elif eweight_normalized == "standard":
                we_norm = 1 / natoms
            elif eweight_normalized == "squared":
                we_norm = 1 / natoms ** 2
            else:
                raise ValueError("eweight_normalized can only be nothing, :standard, or :squared")
            gd = np.asarray(gd, dtype=np.float64)
            if acc is None:
                acc = NormalEquationAccumulator(gd.shape[0], batch_rows)
            acc.add(np.concatenate([gd[None, :], fdm], axis=0), np.concatenate([[np.float64(e)], np.asarray(f, dtype=np.float64)]),
                    np.concatenate([[we_norm * ws[0]], ws[1] * np.ones(fdm.shape[0])]))
        if acc is None:
            raise ValueError("no configurations")
        AtWA, AtWb = acc.result()
    else:
        AtWA, AtWb = np.array(AtWA, copy=True), np.array(AtWb, copy=True)
    if symmetrize:
        AtWA = np.triu(AtWA) + np.triu(AtWA, 1).T  # Symmetric(AtWA)
    if λ is not None:
        if reg_style == "default":
            AtWA = AtWA + λ * np.eye(AtWA.shape[0])
        if reg_style in ("scale_thresh", "scale"):
            for i in range(AtWA.shape[0]):
                reg_elem = AtWA[i, i] * (1 + λ)
                if reg_style == "scale_thresh":
                    reg_elem = max(reg_elem, λ)
                AtWA[i, i] = reg_elem
    β = np.linalg.solve(AtWA, AtWb)
    print(f"condition number of AtWA: {np.linalg.cond(AtWA)}")
    lb_beta[:] = β
    return AtWA, AtWb


# ----------------------------------------------------------------------------------------------------------------
# after the fit: learn!(lb, ds, args...) wrapper and predictions
# ----------------------------------------------------------------------------------------------------------------
class LinearBasisPotential:
    """Stand-in for InteratomicPotentials.LinearBasisPotential ([external]): only the coefficient storage the hot path
    touches (β, β0). The descriptor basis itself (ACE/SNAP evaluation) stays in InteratomicPotentials.jl."""

    def __init__(self, dim: int = 0):
        self.β = np.zeros(dim)
        self.β0 = np.zeros(1)


def learn_potential_(lb: LinearBasisPotential, ds: DataSet, *args, **kw):
    """learn!(lb::LinearBasisPotential, ds::DataSet, args...) (linear-learning-problem.jl:158-170)."""
    lp = LinearProblem(ds)
    learn_(lp, *args, **kw)
    lb.β = np.resize(lb.β, len(lp.β))  # resize!(lb.β, length(lp.β))
    lb.β[:] = lp.β
    lb.β0[:] = lp.β0
    return lp


def predict(A, beta, beta0=0.0, n_energy_rows=0):
    """GPU evaluation y = A·β + β0·[r < n_energy_rows] over packed rows (pl_predict)."""
    A = _lib.as_f64(A)
    R, K = A.shape
    beta = _lib.as_f64(beta)
    if beta.shape != (K,):
        raise ValueError("DimensionMismatch: β does not match the descriptor length")
    y = np.empty(R)
    _lib.check(_lib.lib().pl_predict(_lib.dptr(A), R, K, K, _lib.dptr(beta), float(beta0), int(n_energy_rows), _lib.dptr(y)), "pl_predict")
    return y


def get_all_energies(ds: DataSet, lb: LinearBasisPotential | None = None):
    """get_all_energies(ds) / get_all_energies(ds, lb) (src/Data/utils.jl:10-15, 42-49): reference energies, or the
    model's B·β + β0 with B the per-configuration global-sum descriptors."""
    if lb is None:
        return [get_values(get_energy(c)) for c in ds]
    B = np.stack([np.add.reduce(get_values(get_local_descriptors(c)), axis=0) for c in ds])
    return predict(B, lb.β, lb.β0[0], B.shape[0])


def get_all_forces(ds: DataSet, lb: LinearBasisPotential | None = None):
    """get_all_forces(ds) / get_all_forces(ds, lb) (src/Data/utils.jl:24-30, 58-65): flattened forces, or dB·β."""
    if lb is None:
        return np.concatenate([get_values(get_forces(c)).reshape(-1) for c in ds])
    dB = np.concatenate([get_values(get_force_descriptors(c)).reshape(-1, len(lb.β)) for c in ds], axis=0)
    return predict(dB, lb.β, 0.0, 0)


# ----------------------------------------------------------------------------------------------------------------
# Metrics (src/Metrics/metrics.jl:11-53) of the predictions, reduced on the GPU (pl_metrics)
# ----------------------------------------------------------------------------------------------------------------
def calc_all_metrics(x_pred, x):
    """[mae, rmse, rsq, mean_cos] in one device pass pair (pl_metrics)."""
    xp, xr = _lib.as_f64(x_pred).reshape(-1), _lib.as_f64(x).reshape(-1)
    if xp.shape != xr.shape or xp.size == 0:
        raise ValueError("DimensionMismatch: x_pred and x must have the same non-zero length")
    out = np.empty(4)
    _lib.check(_lib.lib().pl_metrics(_lib.dptr(xp), _lib.dptr(xr), xp.size, _lib.dptr(out)), "pl_metrics")
    return out


def mae(x_pred, x):
    """mae(x_pred, x) (metrics.jl:11-13)."""
    return float(calc_all_metrics(x_pred, x)[0])


def rmse(x_pred, x):
    """rmse(x_pred, x) (metrics.jl:23-25)."""
    return float(calc_all_metrics(x_pred, x)[1])


def rsq(x_pred, x):
    """rsq(x_pred, x) (metrics.jl:35-37)."""
    return float(calc_all_metrics(x_pred, x)[2])


def mean_cos(x_pred, x):
    """mean_cos(x_pred, x) (metrics.jl:47-53): mean cosine between predicted and true force 3-vectors, NaNs filtered."""
    return float(calc_all_metrics(x_pred, x)[3])


def get_metrics(x_pred, x, metrics=(mae, rmse, rsq), label="x"):
    """get_metrics(x_pred, x; metrics, label) (metrics.jl:67-75): {"<label>_<metric>": value} in the given order."""
    vals = calc_all_metrics(x_pred, x)
    idx = {"mae": 0, "rmse": 1, "rsq": 2, "mean_cos": 3}
    return {f"{label}_{m.__name__}": float(vals[idx[m.__name__]]) for m in metrics}

