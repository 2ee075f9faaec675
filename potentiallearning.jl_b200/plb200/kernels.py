"""Host mirror of reference src/Kernels (features.jl, distances.jl, kernels.jl) for the part on the hot path:
GlobalMean / GlobalSum features, Euclidean distance, DotProduct and RBF kernels, and the four KernelMatrix methods.
Same names, argument meaning and error behaviour as the Julia API; the arithmetic runs in libplb200.so on the GPU.

Widened beyond the north-star kernels on the same Gram engine: InverseMultiquadric(Euclidean) and
CorrelationMatrix features (a Symmetric d x d feature is flattened: dot(A,B) and the identity-Cinv Euclidean distance of two
Symmetric matrices are the dot product / squared distance of their d*d entries).
KernelSteinDiscrepancy with an RBF(Euclidean) kernel runs as four Gram matrices + one fused pair reduction (pl_ksd_rbf).
Out of scope (stay Julia): Forstner, stand-alone kernel gradients.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .data import Configuration, DataSet, LocalDescriptors, get_local_descriptors, get_values


# ------------------------------------------------------------------ features (features.jl) ------------------------
class Feature:
    pass


class GlobalSum(Feature):
    """GlobalSum (features.jl:15-21)."""


class GlobalMean(Feature):
    """GlobalMean (features.jl:28-34)."""


class CorrelationMatrix(Feature):
    """CorrelationMatrix (features.jl:37-50): Symmetric(mean(Bi*Bi' for Bi in B)), a d x d global descriptor."""


def compute_feature(x, f: Feature, dt=LocalDescriptors) -> np.ndarray:
    """compute_feature(B::LocalDescriptors, gm) / compute_feature(c::Configuration, gm; dt) (features.jl:19-21, 32-34, 59-65).
    `dt` is accepted and ignored for GlobalMean/GlobalSum exactly as in the reference (features.jl:59-65)."""
    if isinstance(x, Configuration):
        x = get_local_descriptors(x)
    b = get_values(x) if isinstance(x, LocalDescriptors) else np.asarray(x, dtype=np.float64)
    if isinstance(f, CorrelationMatrix):
        acc = np.zeros((b.shape[1], b.shape[1]))
        for bi in b:  # mean(Bi * Bi' ...): sequential sum of outer products / natoms (features.jl:49)
            acc = acc + np.outer(bi, bi)
        return Symmetric(acc / b.shape[0])
    s = np.add.reduce(b, axis=0)  # sequential row accumulation == Julia's sum over a Vector{Vector}
    if isinstance(f, GlobalMean):
        return s / b.shape[0]
    if isinstance(f, GlobalSum):
        return s
    raise TypeError(f"feature {type(f).__name__} is outside the B200 hot path (stays in PotentialLearning.jl)")


def compute_features(ds: DataSet, f: Feature, dt=LocalDescriptors):
    """compute_features(ds, f; dt) = compute_feature.(ds, (f,); dt) (features.jl:76-78).
    GlobalMean / GlobalSum over a whole DataSet run on the GPU (pl_global_feature: one segmented reduction over the stacked local
    descriptors, Julia's summation order); returns the N x d feature matrix (row i = feature of configuration i)."""
    if isinstance(f, (GlobalMean, GlobalSum)):
        return global_features([get_values(get_local_descriptors(c)) for c in ds], mean=isinstance(f, GlobalMean))
    if isinstance(f, CorrelationMatrix) and dt is LocalDescriptors:
        return correlation_features([get_values(get_local_descriptors(c)) for c in ds])
    return [compute_feature(c, f, dt=dt) for c in ds]


def correlation_features(local_blocks):
    """CorrelationMatrix features of N configurations on the GPU (pl_correlation_feature): a list of Symmetric d x d matrices,
    bit-identical to mean(Bi*Bi' for Bi in B) evaluated in the reference's order (features.jl:48-50)."""
    import ctypes

    L, offsets = stack_locals(local_blocks)
    N = len(offsets) - 1
    if N == 0:
        return []
    if np.any(np.diff(offsets) <= 0):
        raise ValueError("ArgumentError: reducing over an empty collection is not allowed (configuration without local descriptors)")
    d = L.shape[1]
    C = np.empty((N, d, d))
    _lib.check(_lib.lib().pl_correlation_feature(_lib.dptr(L), d, offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), N, d, _lib.dptr(C)),
               "pl_correlation_feature")
    return [Symmetric(C[i]) for i in range(N)]


def stack_locals(local_blocks):
    """per-configuration natoms_i x d arrays -> (stacked rows, int64 offsets[N+1])"""
    blocks = [_lib.as_f64(b) for b in local_blocks]
    if any(b.ndim != 2 for b in blocks):
        raise ValueError("local descriptors must be natoms x d per configuration")
    offsets = np.zeros(len(blocks) + 1, dtype=np.int64)
    np.cumsum([b.shape[0] for b in blocks], out=offsets[1:])
    L = np.ascontiguousarray(np.concatenate(blocks, axis=0)) if blocks else np.zeros((0, 1))
    return L, offsets


def global_features(local_blocks, mean: bool = True, stacked=None) -> np.ndarray:
    """GlobalMean (mean=True) / GlobalSum (mean=False) features of N configurations on the GPU (pl_global_feature).
    local_blocks: list of natoms_i x d arrays, or stacked=(L, offsets) if the caller already holds the packed form."""
    import ctypes

    L, offsets = stack_locals(local_blocks) if stacked is None else (_lib.as_f64(stacked[0]), np.ascontiguousarray(stacked[1], dtype=np.int64))
    N = len(offsets) - 1
    if N == 0:
        return np.zeros((0, L.shape[1]))
    if np.any(np.diff(offsets) <= 0):
        raise ValueError("ArgumentError: reducing over an empty collection is not allowed (configuration without local descriptors)")
    d = L.shape[1]
    X = np.empty((N, d))
    _lib.check(_lib.lib().pl_global_feature(_lib.dptr(L), d, offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), N, d, 1 if mean else 0,
                                            _lib.dptr(X), d), "pl_global_feature")
    return X


def global_features_dev(L, offsets, mean: bool = True, out=None):
    """Device-resident form on CUDA tensors: L (total_atoms x d float64), offsets (N+1 int64). Enqueues on the current stream."""
    import torch

    assert L.is_cuda and L.dtype == torch.float64 and L.dim() == 2 and L.stride(1) == 1
    assert offsets.is_cuda and offsets.dtype == torch.int64 and offsets.is_contiguous()
    N, d = offsets.numel() - 1, L.shape[1]
    if out is None:
        out = torch.empty((N, d + (d & 1)), device=L.device, dtype=torch.float64)[:, :d]
    _lib.check(_lib.lib().pl_global_feature_dev(L.data_ptr(), L.stride(0), offsets.data_ptr(), N, d, 1 if mean else 0, out.data_ptr(), out.stride(0),
                                                L.shape[0], torch.cuda.current_stream(L.device).cuda_stream), "pl_global_feature_dev")
    return out


# ------------------------------------------------------------------ distances (distances.jl) ----------------------
class Distance:
    pass


class Diagonal:
    """LinearAlgebra.Diagonal stand-in: Euclidean(Diagonal(v))."""

    def __init__(self, diag):
        self.diag = _lib.as_f64(diag)
        if self.diag.ndim != 1:
            raise ValueError("Diagonal needs a vector")


class Symmetric:
    """LinearAlgebra.Symmetric stand-in. `data` is the parent storage; only the `uplo` triangle is meaningful
    (the reference returns Symmetric(K) with uplo = :U over zeros(n,n) storage, kernels.jl:216-222)."""

    def __init__(self, data, uplo: str = "U"):
        self.data = np.asarray(data, dtype=np.float64)
        if self.data.ndim != 2 or self.data.shape[0] != self.data.shape[1]:
            raise ValueError("Symmetric needs a square matrix")
        self.uplo = uplo

    @property
    def shape(self):
        return self.data.shape

    def __array__(self, dtype=None, copy=None):
        d = self.data
        full = np.triu(d) + np.triu(d, 1).T if self.uplo == "U" else np.tril(d) + np.tril(d, -1).T
        return full if dtype is None else full.astype(dtype)

    def __getitem__(self, idx):
        i, j = idx
        if (self.uplo == "U") == (i <= j):
            return self.data[i, j]
        return self.data[j, i]


class Euclidean(Distance):
    """Euclidean(dim::Int) -> Cinv = 1.0*I(dim); Euclidean(Cinv::Union{Symmetric,Diagonal}) (distances.jl:39-51)."""

    def __init__(self, arg):
        if isinstance(arg, (int, np.integer)):
            self.dim = int(arg)
            self.kind = _lib.PL_CINV_IDENTITY
            self.Cinv = None
        elif isinstance(arg, Diagonal):
            self.dim = arg.diag.shape[0]
            self.kind = _lib.PL_CINV_DIAGONAL
            self.Cinv = arg.diag
        elif isinstance(arg, Symmetric):
            self.dim = arg.shape[0]
            self.kind = _lib.PL_CINV_FULL
            self.Cinv = np.ascontiguousarray(np.asarray(arg))
        else:
            raise TypeError("Euclidean(dim::Int) or Euclidean(Cinv::Union{Symmetric,Diagonal})")


# ------------------------------------------------------------------ kernels (kernels.jl) --------------------------
class Kernel:
    pass


class DotProduct(Kernel):
    """DotProduct(; α = 2) (kernels.jl:21-26); α is an Int power."""

    def __init__(self, alpha: int = 2):
        if not isinstance(alpha, (int, np.integer)):
            raise TypeError("DotProduct.α is an Int")
        self.alpha = int(alpha)


class RBF(Kernel):
    """RBF(d; α = 1e-8, ℓ = 1.0, β = 1.0) (kernels.jl:51-59). α is stored but never applied (kernels.jl:73-74)."""

    def __init__(self, d: Distance, alpha=1e-8, ell=1.0, beta=1.0):
        if not isinstance(d, Euclidean):
            raise TypeError("only RBF(Euclidean) is on the B200 hot path (Forstner stays in PotentialLearning.jl)")
        self.d, self.alpha, self.ell, self.beta = d, alpha, ell, beta


class InverseMultiquadric(Kernel):
    """InverseMultiquadric(d; c2 = 1.0, ℓ = 1.0) (kernels.jl:139-152): (c2 + d2/ℓ²)^(-1/2); asserts 0 < c2, 0 < ℓ."""

    def __init__(self, d: Distance, c2=1.0, ell=1.0):
        if not isinstance(d, Euclidean):
            raise TypeError("only InverseMultiquadric(Euclidean) is on the B200 Gram engine")
        assert 0 < c2
        assert 0 < ell
        self.d, self.c2, self.ell = d, c2, ell


def get_parameters(k: Kernel):
    """get_parameters (kernels.jl:28, 61, 153)."""
    if isinstance(k, DotProduct):
        return (k.alpha,)
    if isinstance(k, InverseMultiquadric):
        return (k.c2, k.ell)
    return (k.alpha, k.ell, k.beta)


def _kargs(k: Kernel, d: int):
    if isinstance(k, DotProduct):
        return _lib.PL_KERNEL_DOT, k.alpha, 0, None, 1.0, 1.0
    if isinstance(k, RBF):
        e = k.d
        if e.dim != d:
            raise ValueError(f"DimensionMismatch: Euclidean of dimension {e.dim} applied to features of length {d}")
        return _lib.PL_KERNEL_RBF, 0, e.kind, e.Cinv, float(k.ell), float(k.beta)
    if isinstance(k, InverseMultiquadric):
        e = k.d
        if e.dim != d:
            raise ValueError(f"DimensionMismatch: Euclidean of dimension {e.dim} applied to features of length {d}")
        return _lib.PL_KERNEL_IMQ, 0, e.kind, e.Cinv, float(k.ell), float(k.c2)
    raise TypeError(f"kernel {type(k).__name__} is outside the B200 hot path")


def _features_matrix(F) -> np.ndarray:
    """Vector{Vector{T}} -> row-major n x d (== Julia's reduce(hcat, F), column-major d x n).
    Vector{Symmetric} features (CorrelationMatrix) are flattened to their d*d entries: dot(A,B) of two matrices and the
    identity-Cinv Euclidean distance tr((C1-C2)'(C1-C2)) (distances.jl:62-68) are exactly the vector forms of the entries."""
    if len(F) > 0 and isinstance(F[0], Symmetric):
        return np.ascontiguousarray(np.stack([np.asarray(f).reshape(-1) for f in F]))
    X = _lib.as_f64(F)
    if X.ndim != 2:
        raise ValueError("features must be n vectors of equal length")
    return X


def compute_distance(B1, B2, e: Euclidean) -> float:
    """(B1-B2)'*Cinv*(B1-B2) (distances.jl:58-60), evaluated on the GPU as -2*log(RBF) of a 1 x 1 kernel matrix would lose
    accuracy, so it is obtained from the LINEAR Gram of the difference vector: d2 = v' (Cinv v)."""
    v = (_lib.as_f64(B1) - _lib.as_f64(B2))[None, :]
    if e.kind == _lib.PL_CINV_IDENTITY:
        w = v
    elif e.kind == _lib.PL_CINV_DIAGONAL:
        w = v * e.Cinv[None, :]
    else:
        w = KernelMatrix(v, e.Cinv, _Linear())
    return float(KernelMatrix(v, w, _Linear())[0, 0])


class _Linear(Kernel):
    """plain Gram X1*X2' (PL_KERNEL_LINEAR)"""


def compute_kernel(A, B, k: Kernel) -> float:
    """compute_kernel(A, B, k) for two features, vectors or Symmetric matrices (kernels.jl:30-36, 68-75, 156-164)."""
    if isinstance(A, Symmetric):
        return float(KernelMatrix([A], [B], k)[0, 0])
    return float(KernelMatrix([_lib.as_f64(A)], [_lib.as_f64(B)], k)[0, 0])


def KernelMatrix(*args, dt=LocalDescriptors, fill: str = "full", out=None):
    """The four reference methods (kernels.jl:211-269), dispatched on argument types:

    KernelMatrix(F, k)            -> Symmetric (n x n). fill="full": both triangles written; fill="U": only the upper
                                     triangle of `.data` (i <= j), the rest zeros — the reference's exact storage.
    KernelMatrix(F1, F2, k)       -> ndarray m x n
    KernelMatrix(ds, f, k; dt)    -> features then the symmetric form
    KernelMatrix(ds1, ds2, f, k)  -> features then the rectangular form
    """
    if len(args) == 2:
        F, k = args
        X = _features_matrix(F)
        return _kernel_matrix_sym(X, _for_flattened(k, F, X.shape[1]), fill, out)
    if len(args) == 3 and isinstance(args[0], DataSet):
        ds, f, k = args
        return KernelMatrix(compute_features(ds, f, dt=dt), k, fill=fill, out=out)
    if len(args) == 3:
        F1, F2, k = args
        X1, X2 = _features_matrix(F1), _features_matrix(F2)
        return _kernel_matrix_cross(X1, X2, _for_flattened(k, F1, X1.shape[1]))
    if len(args) == 4:
        ds1, ds2, f, k = args
        return KernelMatrix(compute_features(ds1, f, dt=dt), compute_features(ds2, f, dt=dt), k)
    raise TypeError("KernelMatrix(F,k) | (F1,F2,k) | (ds,f,k) | (ds1,ds2,f,k)")


def _for_flattened(k: Kernel, F, d_flat: int) -> Kernel:
    """Kernel acting on flattened Symmetric features: the Euclidean dimension refers to the matrix order."""
    if len(F) > 0 and isinstance(F[0], Symmetric) and isinstance(k, (RBF, InverseMultiquadric)):
        if k.d.kind != _lib.PL_CINV_IDENTITY:
            raise TypeError("Euclidean(Cinv) on Symmetric features needs tr(Csqrt*(C1-C2)'*Cinv*(C1-C2)*Csqrt): stays in PotentialLearning.jl")
        if k.d.dim ** 2 != d_flat:
            raise ValueError("DimensionMismatch: Euclidean dimension does not match the matrix features")
        e = Euclidean(d_flat)
        return RBF(e, k.alpha, k.ell, k.beta) if isinstance(k, RBF) else InverseMultiquadric(e, k.c2, k.ell)
    return k


def _kernel_matrix_sym(X: np.ndarray, k: Kernel, fill: str, out=None) -> Symmetric:
    n, d = X.shape
    lib = _lib.lib()
    if out is None:
        K = np.zeros((n, n))  # zeros(n, n) (kernels.jl:216)
    else:
        # a caller-owned result buffer, typically _lib.pinned_zeros((n, n)) kept across calls: D2H into pinned memory runs at the PCIe
        # rate (36 ms for config 2) instead of 105 ms (pageable) / 510 ms (fresh calloc pages)
        K = out
        if K.shape != (n, n) or K.dtype != np.float64 or not K.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 n x n array")
        K.fill(0.0)
    fl = _lib.PL_FILL_FULL if fill == "full" else _lib.PL_FILL_RM_UPPER
    if isinstance(k, _Linear):
        _lib.check(lib.pl_gram_cross(_lib.PL_KERNEL_LINEAR, _lib.dptr(X), n, d, _lib.dptr(X), n, d, d, 0, 0, None, 1.0, 1.0, _lib.dptr(K), n), "pl_gram_cross")
        return Symmetric(K)
    kid, alpha, kind, cinv, ell, beta = _kargs(k, d)
    if kid == _lib.PL_KERNEL_DOT:
        _lib.check(lib.pl_gram_dot(_lib.dptr(X), n, d, d, alpha, _lib.dptr(K), n, fl), "pl_gram_dot")
    elif kid == _lib.PL_KERNEL_IMQ:
        _lib.check(lib.pl_gram_imq(_lib.dptr(X), n, d, d, kind, _lib.dptr(cinv), beta, ell, _lib.dptr(K), n, fl), "pl_gram_imq")
    else:
        _lib.check(lib.pl_gram_rbf(_lib.dptr(X), n, d, d, kind, _lib.dptr(cinv), ell, beta, _lib.dptr(K), n, fl), "pl_gram_rbf")
    return Symmetric(K, "U")


def _kernel_matrix_cross(X1: np.ndarray, X2: np.ndarray, k: Kernel) -> np.ndarray:
    m, d = X1.shape
    n, d2 = X2.shape
    if d != d2:
        raise ValueError("DimensionMismatch: features of different length")
    lib = _lib.lib()
    K = np.zeros((m, n))  # zeros(m, n) (kernels.jl:236)
    if isinstance(k, _Linear):
        kid, alpha, kind, cinv, ell, beta = _lib.PL_KERNEL_LINEAR, 0, 0, None, 1.0, 1.0
    else:
        kid, alpha, kind, cinv, ell, beta = _kargs(k, d)
    _lib.check(lib.pl_gram_cross(kid, _lib.dptr(X1), m, d, _lib.dptr(X2), n, d, d, alpha, kind, _lib.dptr(cinv), ell, beta, _lib.dptr(K), n), "pl_gram_cross")
    return K


# ------------------------------------------------------------------ device-resident form --------------------------
def kernel_matrix_dev(X, k: Kernel, out=None, X2=None, fill: int = _lib.PL_FILL_FULL, part: int = 0, nparts: int = 1, packed: bool = False):
    """Kernel matrix of a CUDA float64 torch tensor X (n x d, row-major) without leaving the device.
    part/nparts shard the 128x128 tiles round-robin across GPUs; packed=True stores only the owned tiles."""
    import torch

    lib = _lib.lib()
    assert X.is_cuda and X.dtype == torch.float64 and X.dim() == 2 and X.stride(1) == 1
    n, d = X.shape
    sym = X2 is None
    n2 = n if sym else X2.shape[0]
    kid, alpha, kind, cinv, ell, beta = _kargs(k, d) if not isinstance(k, _Linear) else (_lib.PL_KERNEL_LINEAR, 0, 0, None, 1.0, 1.0)
    dcinv = None
    if cinv is not None:
        dcinv = torch.as_tensor(cinv, device=X.device, dtype=torch.float64).contiguous()
    if out is None:
        if packed:
            cnt = int(lib.pl_gram_tile_count(n, 0 if sym else n2, part, nparts))
            out = torch.empty((cnt, 128, 128), device=X.device, dtype=torch.float64)
        else:
            out = torch.zeros((n, n2), device=X.device, dtype=torch.float64)
    ldo = 128 if packed else out.stride(0)
    stream = torch.cuda.current_stream(X.device).cuda_stream
    _lib.check(
        lib.pl_gram_dev(kid, X.data_ptr(), n, X.stride(0), None if sym else X2.data_ptr(), n2, 0 if sym else X2.stride(0), d, alpha, kind,
                        None if dcinv is None else dcinv.data_ptr(), ell, beta, out.data_ptr(), ldo, fill, part, nparts, 1 if packed else 0,
                        stream),
        "pl_gram_dev",
    )
    return out


def gram_tile_coords(n1: int, n2: int, part: int, nparts: int):
    """(ti, tj) in units of 128 of every tile owned by `part` (same order as the packed output); (-1,-1) = empty slot."""
    import ctypes

    lib = _lib.load()
    cnt = int(lib.pl_gram_tile_count(n1, n2, part, nparts))
    ti, tj = ctypes.c_int64(), ctypes.c_int64()
    out = np.empty((cnt, 2), dtype=np.int64)
    for k in range(cnt):
        _lib.check(lib.pl_gram_tile_coords(n1, n2, part, nparts, k, ctypes.byref(ti), ctypes.byref(tj)), "pl_gram_tile_coords")
        out[k] = (ti.value, tj.value)
    return out


# ------------------------------------------------------------------ divergences (divergences.jl) ------------------
class Divergence:
    """abstract type Divergence (divergences.jl:7)."""


class KernelSteinDiscrepancy(Divergence):
    """KernelSteinDiscrepancy(; score, kernel) (divergences.jl:18-25)."""

    def __init__(self, score=None, kernel=None):
        if score is None or kernel is None:
            raise TypeError("KernelSteinDiscrepancy(; score, kernel)")
        self.score, self.kernel = score, kernel


def compute_divergence(x, div: KernelSteinDiscrepancy) -> float:
    """compute_divergence(x, div::KernelSteinDiscrepancy) (divergences.jl:27-49) on the GPU (pl_ksd_rbf): the scores div.score.(x) are
    evaluated here (a user function, as in the reference); the O(N^2 d) pair loop runs as four Gram matrices + one fused reduction.
    Only RBF(Euclidean) kernels (the reference's IMQ cross derivative multiplies two column vectors and throws for d > 1)."""
    k = div.kernel
    if not isinstance(k, RBF):
        raise TypeError("KernelSteinDiscrepancy on the GPU needs an RBF(Euclidean) kernel")
    X = _features_matrix([np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in x])
    S = _lib.as_f64([np.atleast_1d(np.asarray(div.score(v), dtype=np.float64)) for v in X])
    n, d = X.shape
    if S.shape != X.shape:
        raise ValueError("DimensionMismatch: score(x) must have the length of x")
    _, _, kind, cinv, ell, beta = _kargs(k, d)
    out = np.zeros(1)
    _lib.check(_lib.lib().pl_ksd_rbf(_lib.dptr(X), _lib.dptr(S), n, d, d, d, kind, _lib.dptr(cinv), ell, beta, _lib.dptr(out)), "pl_ksd_rbf")
    return float(out[0])

