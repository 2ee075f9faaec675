"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in plb200/distributed.py: row sharding, the single
all-reduce of the SYRK partials (+ the earlier column-sum all-reduce of the centred covariance), identical results on
every rank. The compute engine injected here is a CHECKER built on the oracle (numpy on CPU tensors) — test
infrastructure only; the product engine is plb200.device.DeviceEngine (libplb200.so on a B200)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class CheckerEngine:
    """Same interface as plb200.device.DeviceEngine; CPU tensors, oracle arithmetic."""

    name = "oracle-checker"

    def syrk_partial(self, A, w=None, n_energy_rows=0, w_e=1.0, w_f=1.0, b=None, mean=None, want_vec=False):
        An = A.numpy()
        R, K = An.shape
        wn = w.numpy() if w is not None else np.concatenate([w_e * np.ones(n_energy_rows), w_f * np.ones(R - n_energy_rows)])
        Y = An - mean.numpy() if mean is not None else An
        S = torch.from_numpy(Y.T @ (wn[:, None] * Y))
        vec = None
        if want_vec:
            bn = b.numpy() if b is not None else np.zeros(R)
            ind = np.concatenate([np.ones(n_energy_rows), np.zeros(R - n_energy_rows)])
            vec = torch.from_numpy(np.concatenate([Y.T @ (wn * bn), Y.T @ (wn * ind), [np.sum(wn * ind), np.sum(wn * ind * bn)]]))
        return S, vec

    def normal_eq_finish(self, S, vec, intercept=False, lam=0.0, want_rhs=True):
        K = S.shape[0]
        Sn = S.numpy()
        if intercept:
            v = vec.numpy()
            M = np.zeros((K + 1, K + 1))
            M[1:, 1:] = Sn
            M[0, 1:] = M[1:, 0] = v[K:2 * K]
            M[0, 0] = v[2 * K]
            rhs = np.concatenate([[v[2 * K + 1]], v[:K]])
        else:
            M = Sn.copy()
            rhs = vec.numpy()[:K] if vec is not None else None
        M = np.triu(M) + np.triu(M, 1).T + lam * np.eye(M.shape[0])  # j >= i triangle only, like normal_eq_finish_kernel
        return torch.from_numpy(M), (torch.from_numpy(rhs) if (want_rhs and rhs is not None) else None)

    def cov_shifted(self, X, z):
        Y = X.numpy() - z.numpy()
        return torch.from_numpy(Y.T @ Y), torch.from_numpy(Y.sum(axis=0))

    def cov_recentre(self, S, rho, delta, n_rows):
        r, dl = rho.numpy(), delta.numpy()
        S -= torch.from_numpy(np.outer(r, dl) + np.outer(dl, r) - n_rows * np.outer(dl, dl))
        return S

    def colsum(self, A):
        return A.sum(dim=0)

    def scale_div(self, x, denom):
        return x / denom

    def sym_scale_div(self, S, denom):
        return (torch.triu(S) + torch.triu(S, 1).T) / denom


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from plb200.distributed import covariance_sharded, normal_equations_sharded, shard_rows

    rng = np.random.default_rng(123)
    ncfg, rows_per_cfg, K = 10, 13, 7  # 1 energy row + 12 force rows per configuration
    Ae = rng.standard_normal((ncfg, K))
    Af = rng.standard_normal((ncfg * 12, K))
    be, bf = rng.standard_normal(ncfg), rng.standard_normal(ncfg * 12)
    # shard whole configurations: each rank holds [its energy rows ; its force rows]
    c0, c1 = shard_rows(ncfg, rank, world)
    A_loc = torch.from_numpy(np.vstack([Ae[c0:c1], Af[12 * c0:12 * c1]]))
    b_loc = torch.from_numpy(np.concatenate([be[c0:c1], bf[12 * c0:12 * c1]]))
    eng = CheckerEngine()
    M, v = normal_equations_sharded(A_loc, b_loc, None, c1 - c0, 100.0, 1.0, True, 0.01, engine=eng)
    X = 10.0 * rng.standard_normal(K) + rng.standard_normal((101, K))
    r0, r1 = shard_rows(101, rank, world)
    C, m = covariance_sharded(torch.from_numpy(X[r0:r1].copy()), 101, engine=eng)
    C2, m2 = covariance_sharded(torch.from_numpy(X[r0:r1].copy()), 101, mean=torch.from_numpy(np.ones(K)), engine=eng)
    C3, _ = covariance_sharded(torch.from_numpy(X[r0:r1].copy()), 101, center=False, engine=eng)
    np.savez(os.path.join(tmpdir, f"rank{rank}.npz"), M=M.numpy(), v=v.numpy(), C=C.numpy(), m=m.numpy(), C2=C2.numpy(), C3=C3.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_normal_equations_and_covariance_world2(tmp_path):
    from oracle import oracle as orc

    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    ncfg, K = 10, 7
    Ae = rng.standard_normal((ncfg, K))
    Af = rng.standard_normal((ncfg * 12, K))
    be, bf = rng.standard_normal(ncfg), rng.standard_normal(ncfg * 12)
    A = np.vstack([Ae, Af])
    Ai = np.hstack([np.concatenate([np.ones(ncfg), np.zeros(ncfg * 12)])[:, None], A])
    q = np.concatenate([100.0 * np.ones(ncfg), np.ones(ncfg * 12)])
    Mo, vo = orc.normal_equations(Ai, np.concatenate([be, bf]), q, 0.01)
    X = 10.0 * rng.standard_normal(K) + rng.standard_normal((101, K))
    mo, Co = orc.centered_cov(X)
    _, Co2 = orc.centered_cov(X, np.ones(K))
    res = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    for r in res:
        assert np.allclose(r["M"], Mo, rtol=1e-12, atol=1e-12) and np.allclose(r["v"], vo, rtol=1e-12, atol=1e-12)
        assert np.allclose(r["m"], mo, rtol=1e-14) and np.allclose(r["C"], Co, rtol=1e-11, atol=1e-13)
        assert np.allclose(r["C2"], Co2, rtol=1e-11, atol=1e-12)
        assert np.allclose(r["C3"], X.T @ X / 101, rtol=1e-12)
    for k in ("M", "v", "C", "m"):
        assert np.array_equal(res[0][k], res[1][k]), "every rank ends with the identical all-reduced result"
