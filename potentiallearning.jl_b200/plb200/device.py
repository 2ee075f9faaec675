"""Device-resident entry points on CUDA float64 torch tensors (torch is plumbing: device memory and streams only).
Every function enqueues on torch's current stream through the `_dev` C ABI of libplb200.so and does not synchronise."""
from __future__ import annotations

from . import _lib


def _stream(t):
    import torch

    return torch.cuda.current_stream(t.device).cuda_stream


def _check_mat(t, name):
    import torch

    if not (t.is_cuda and t.dtype == torch.float64 and t.dim() == 2 and t.stride(1) == 1):
        raise ValueError(f"{name} must be a CUDA float64 row-major matrix")
    # the library's workspace lives on ONE device: bind it to the tensor's device at first use, refuse tensors of any other device
    _lib.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


class DeviceEngine:
    """The B200 compute engine used by plb200.distributed (and directly by the benchmarks)."""

    name = "b200"

    def syrk_partial(self, A, w=None, n_energy_rows=0, w_e=1.0, w_f=1.0, b=None, mean=None, want_vec=False):
        """Unscaled partial sums of a row shard (pl_syrk_dev): S (K x K) and, if want_vec, vec = [v ; c ; s] (2K+2)."""
        import torch

        _check_mat(A, "A")
        R, K = A.shape
        S = torch.empty((K, K), device=A.device, dtype=torch.float64)
        vec = torch.empty(2 * K + 2, device=A.device, dtype=torch.float64) if want_vec else None
        _lib.check(
            _lib.lib().pl_syrk_dev(A.data_ptr(), R, K, A.stride(0), None if w is None else w.data_ptr(), int(n_energy_rows), float(w_e),
                                   float(w_f), None if b is None else b.data_ptr(), None if mean is None else mean.data_ptr(), S.data_ptr(),
                                   None if vec is None else vec.data_ptr(), _stream(A)),
            "pl_syrk_dev",
        )
        return S, vec

    def normal_eq_finish(self, S, vec, intercept=False, lam=0.0, want_rhs=True):
        import torch

        K = S.shape[0]
        n = K + (1 if intercept else 0)
        M = torch.empty((n, n), device=S.device, dtype=torch.float64)
        v = torch.empty(n, device=S.device, dtype=torch.float64) if (want_rhs and vec is not None) else None
        _lib.check(
            _lib.lib().pl_normal_eq_finish_dev(S.data_ptr(), None if vec is None else vec.data_ptr(), K, 1 if intercept else 0, float(lam),
                                               M.data_ptr(), None if v is None else v.data_ptr(), _stream(S)),
            "pl_normal_eq_finish_dev",
        )
        return M, v

    def normal_eq_sharded(self, A, b=None, w=None, n_energy_rows=0, w_e=1.0, w_f=1.0, intercept=False, lam=0.0):
        """(A'WA + lam I, A'Wb) over ALL ranks' shards (pl_normal_eq_sharded_dev): SYRK -> exchange -> finish inside the library, one stream."""
        import torch

        _check_mat(A, "A")
        R, K = A.shape
        n = K + (1 if intercept else 0)
        M = torch.empty((n, n), device=A.device, dtype=torch.float64)
        v = torch.empty(n, device=A.device, dtype=torch.float64) if b is not None else None
        _lib.check(
            _lib.lib().pl_normal_eq_sharded_dev(A.data_ptr(), R, K, A.stride(0), None if w is None else w.data_ptr(), int(n_energy_rows), float(w_e), float(w_f),
                                                None if b is None else b.data_ptr(), 1 if intercept else 0, float(lam), M.data_ptr(),
                                                None if v is None else v.data_ptr(), _stream(A)),
            "pl_normal_eq_sharded_dev",
        )
        return M, v

    def cov_sharded(self, X, n_total, mean=None, center=True):
        """(Q, m) over ALL ranks' rows (pl_cov_sharded_dev): one pass, the K and K x K exchanges inside the library."""
        import torch

        _check_mat(X, "X")
        n, K = X.shape
        C = torch.empty((K, K), device=X.device, dtype=torch.float64)
        m = None
        if center:
            m = mean.clone() if mean is not None else torch.empty(K, device=X.device, dtype=torch.float64)
        _lib.check(
            _lib.lib().pl_cov_sharded_dev(X.data_ptr(), n, int(n_total), K, X.stride(0), None if m is None else m.data_ptr(), 1 if mean is not None else 0,
                                          1 if center else 0, C.data_ptr(), _stream(X)),
            "pl_cov_sharded_dev",
        )
        return C, m

    def cov_shifted(self, X, z):
        """S_z = sum (x - z)(x - z)' and rho = sum (x - z) of a row shard in ONE pass over the rows (pl_cov_shifted_dev)."""
        import torch

        _check_mat(X, "X")
        R, K = X.shape
        S = torch.empty((K, K), device=X.device, dtype=torch.float64)
        rho = torch.empty(K, device=X.device, dtype=torch.float64)
        _lib.check(_lib.lib().pl_cov_shifted_dev(X.data_ptr(), R, K, X.stride(0), z.data_ptr(), S.data_ptr(), rho.data_ptr(), _stream(X)), "pl_cov_shifted_dev")
        return S, rho

    def cov_recentre(self, S, rho, delta, n_rows):
        """in place: S <- S - rho delta' - delta rho' + n delta delta' (moves a second moment about z to the mean z + delta)"""
        _lib.check(_lib.lib().pl_cov_recentre_dev(S.data_ptr(), S.shape[0], rho.data_ptr(), delta.data_ptr(), float(n_rows), _stream(S)), "pl_cov_recentre_dev")
        return S

    def colsum(self, A):
        import torch

        _check_mat(A, "A")
        R, K = A.shape
        s = torch.empty(K, device=A.device, dtype=torch.float64)
        _lib.check(_lib.lib().pl_colsum_dev(A.data_ptr(), R, K, A.stride(0), s.data_ptr(), _stream(A)), "pl_colsum_dev")
        return s

    def scale_div(self, x, denom):
        import torch

        out = torch.empty_like(x)
        _lib.check(_lib.lib().pl_scale_div_dev(x.data_ptr(), float(denom), x.numel(), out.data_ptr(), _stream(x)), "pl_scale_div_dev")
        return out

    def sym_scale_div(self, S, denom):
        """out = sym(S) / denom from the j >= i triangle: exactly symmetric whatever order the all-reduce summed in."""
        import torch

        K = S.shape[0]
        out = torch.empty_like(S)
        _lib.check(_lib.lib().pl_sym_scale_div_dev(S.data_ptr(), K, float(denom), out.data_ptr(), _stream(S)), "pl_sym_scale_div_dev")
        return out
