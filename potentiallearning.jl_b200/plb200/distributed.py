"""Multi-GPU host logic, one process per GPU. On B200s the whole step — SYRK, the ONE exchange (peer memory over NVLink, NCCL
fallback) and the finish kernel — runs inside libplb200 (pl_normal_eq_sharded_dev / pl_cov_sharded_dev after pl_comm_init_rank; see
csrc/multi.inl): the functions below hand over to it whenever the library owns a communicator. The explicit form that follows
(engine partials + an all-reduce through torch.distributed) is the same sharding logic spelled out; it is what the gloo / CPU tests
exercise with a checker engine, and what runs on GPUs when only a torch process group exists.
The reference has no distributed code at all; sharding follows section 8(e):

  * normal equations / covariance: shard ROWS (whole configurations per rank), each rank produces unscaled partial
    sums with the fused SYRK kernel, ONE all-reduce of K*K (+2K+2) doubles, then the tiny finish kernel;
    the centred covariance needs one earlier all-reduce of the K column sums, but still only one pass over the rows
    (shifted second moment + re-centring, see covariance_sharded).
  * Gram matrices: tiles are dealt round-robin to ranks; X is replicated; NO collective on the result (tile-packed
    shards stay on their GPU).

`engine` is the compute backend: DeviceEngine (libplb200.so on a B200) in production — there is no CPU engine in
the product. The gloo/CPU tests inject a checker engine built on the oracle to exercise this file's logic only.
"""
from __future__ import annotations

from .device import DeviceEngine


def shard_rows(n_rows: int, rank: int, world: int, align: int = 1):
    """Contiguous, balanced row range of `rank`; boundaries are multiples of `align` (e.g. rows per configuration)."""
    units = (n_rows + align - 1) // align
    base, rem = divmod(units, world)
    lo_u = rank * base + min(rank, rem)
    hi_u = lo_u + base + (1 if rank < rem else 0)
    return min(lo_u * align, n_rows), min(hi_u * align, n_rows)


def _library_owns_comm() -> bool:
    from . import _lib

    return _lib._inited and _lib.comm_info()["nranks"] > 1


def _all_reduce(t, group):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def normal_equations_sharded(A_local, b_local=None, w_local=None, n_energy_rows_local=0, w_e=1.0, w_f=1.0, intercept=False, lam=0.0,
                             group=None, engine=None):
    """(A'WA + λI, A'Wb) of the row-wise concatenation of every rank's shard. Each rank's shard is laid out
    [its energy rows ; its force rows]. Returns identical (M, v) on every rank."""
    import torch

    engine = engine or DeviceEngine()
    if isinstance(engine, DeviceEngine) and group is None and _library_owns_comm():
        return engine.normal_eq_sharded(A_local, b_local, w_local, n_energy_rows_local, w_e, w_f, intercept, lam)
    want_vec = intercept or (b_local is not None)
    S, vec = engine.syrk_partial(A_local, w_local, n_energy_rows_local, w_e, w_f, b_local, None, want_vec)
    if vec is not None:
        flat = torch.cat([S.reshape(-1), vec])  # one bucket -> one all-reduce
        _all_reduce(flat, group)
        K = S.shape[0]
        S, vec = flat[: K * K].view(K, K), flat[K * K:]
    else:
        _all_reduce(S, group)
    return engine.normal_eq_finish(S, vec, intercept, lam, want_rhs=b_local is not None)


def covariance_sharded(X_local, n_total: int, mean=None, center=True, group=None, engine=None):
    """(Q, m) with Q = (1/n) sum_r (x_r - m)(x_r - m)' over all ranks' rows (compute_eigen's Q + PCAState centring).

    center and no mean given: ONE pass over the rows per rank. Each rank shifts by z = the mean of a prefix of its own rows, the SYRK
    returns S_z = sum (x - z)(x - z)' and rho = sum (x - z); the column sums n_r z + rho are all-reduced for the global mean m
    (pca.m = sum(d)/length(d), pca_state.jl:35), every rank re-centres its S_z about m (delta = m - z: tiny, nothing cancels), and ONE
    all-reduce adds the shards. A persisted mean (pca.m set, pca_state.jl:34) centres directly."""
    engine = engine or DeviceEngine()
    if isinstance(engine, DeviceEngine) and group is None and _library_owns_comm():
        return engine.cov_sharded(X_local, n_total, mean, center)
    n_local = X_local.shape[0]
    if center and mean is None:
        prefix = min(n_local, 65536)
        if prefix > 0:
            z = engine.scale_div(engine.colsum(X_local[:prefix]), float(prefix))
        else:
            z = X_local.new_zeros(X_local.shape[1])
        S, rho = engine.cov_shifted(X_local, z)
        s = z * float(n_local) + rho  # this shard's column sums (K numbers: glue, not compute)
        _all_reduce(s, group)
        m = engine.scale_div(s, float(n_total))
        engine.cov_recentre(S, rho, m - z, float(n_local))
        _all_reduce(S, group)
        return engine.sym_scale_div(S, float(n_total)), m
    m = mean if center else None
    S, _ = engine.syrk_partial(X_local, None, 0, 1.0, 1.0, None, m, False)
    _all_reduce(S, group)
    return engine.sym_scale_div(S, float(n_total)), m  # exactly symmetric even after a >2-rank ring all-reduce
