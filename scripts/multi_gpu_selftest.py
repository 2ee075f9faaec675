#!/usr/bin/env python
"""Multi-GPU self-test of libplb200 against the CPU oracle / numpy fp64 (needs >= 2 B200s; run by tests/test_multi_gpu.py and by hand).

  python scripts/multi_gpu_selftest.py single [ndev]       ONE process drives ndev GPUs (pl_init_all): the host entry points shard
                                                           transparently — what a Julia session sees.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_selftest.py ranks
                                                           one process per GPU (pl_init + pl_comm_init_rank): the *_sharded entry points.
Prints one JSON line with the measured errors; exits non-zero on a failed check. Tolerance 1e-11 (contract 1e-10)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "potentiallearning.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

RTOL = 1e-11


def normerr(C, Cref):
    scale = np.max(np.abs(Cref))
    e = float(np.max(np.abs(C - Cref)) / scale)
    big = np.abs(Cref) > 1e-3 * scale
    if big.any():
        e = max(e, float(np.max(np.abs(C - Cref)[big] / np.abs(Cref)[big])))
    return e


def ref_normal_eq(A, w, b, ne, ic, lam):
    """(A'QA + lam I, A'Qb) with the intercept column [1_N; 0_M] first (linear-learn.jl:238-253), numpy fp64"""
    if ic:
        col = np.zeros((A.shape[0], 1))
        col[:ne] = 1.0
        A = np.hstack([col, A])
    WA = A * w[:, None]
    M = A.T @ WA + lam * np.eye(A.shape[1])
    return M, WA.T @ b


def problem(seed, R, K, ne):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((R, K))
    A[:ne] *= np.sqrt(96.0)
    b = rng.standard_normal(R)
    return A, b


def check(name, err, out, tol=RTOL):
    out[name] = err
    if not (err < tol):
        print(json.dumps(out))
        raise SystemExit(f"FAILED {name}: {err} >= {tol}")


def gram_rows_ref(orc, X, k, rows):
    return orc.kernel_matrix_cross(X[rows], X, k)


def run_single(ndev):
    import ctypes

    import plb200
    from oracle import oracle as orc
    from plb200 import _lib

    lib = _lib.init_all(ndev)
    nd = _lib.device_count()
    out = {"mode": "single", "devices": nd, "nccl": _lib.comm_info()}
    assert nd == ndev or ndev == 0

    # ---- normal equations: 90 000 x 500 (360 MB >= the 256 MB sharding threshold), two-segment weights, intercept, ridge ----
    R, K, ne = 90_000, 500, 311
    A, b = problem(1, R, K, ne)
    w2 = np.where(np.arange(R) < ne, 100.0, 1.0)
    Mref, vref = ref_normal_eq(A, w2, b, ne, True, 0.01)
    M, v = np.zeros((K + 1, K + 1)), np.zeros(K + 1)
    t0 = time.perf_counter()
    _lib.check(lib.pl_normal_eq(_lib.dptr(A), R, K, K, None, ne, 100.0, 1.0, _lib.dptr(b), 1, 0.01, _lib.dptr(M.reshape(-1)), _lib.dptr(v)), "pl_normal_eq")
    out["normal_eq_s"] = time.perf_counter() - t0
    check("normal_eq_M", normerr(M, Mref), out)
    check("normal_eq_v", normerr(v, vref), out)
    assert np.array_equal(M, M.T), "exact symmetry after the all-reduce"
    out["coll_ms"] = _lib.last_coll_ms() if nd > 1 else 0.0
    # per-row weights (ooc_learn!'s per-configuration weights), no intercept, no rhs
    wr = np.random.default_rng(2).random(R) + 0.5
    Mref2, _ = ref_normal_eq(A, wr, b, ne, False, 0.0)
    M2 = np.zeros((K, K))
    _lib.check(lib.pl_normal_eq(_lib.dptr(A), R, K, K, _lib.dptr(wr), ne, 1.0, 1.0, None, 0, 0.0, _lib.dptr(M2.reshape(-1)), None), "pl_normal_eq (w)")
    check("normal_eq_rowweights", normerr(M2, Mref2), out)
    # the same rows as a LIST of blocks (lp.B block + ragged lp.dB blocks), sharded across block boundaries
    cuts = [0, ne] + sorted(np.random.default_rng(3).choice(np.arange(ne + 1, R), 37, replace=False).tolist()) + [R]
    blocks = [np.ascontiguousarray(A[cuts[i]:cuts[i + 1]]) for i in range(len(cuts) - 1)]
    ptrs = (ctypes.POINTER(ctypes.c_double) * len(blocks))(*[_lib.dptr(x) for x in blocks])
    rows = (ctypes.c_int64 * len(blocks))(*[x.shape[0] for x in blocks])
    M3, v3 = np.zeros((K + 1, K + 1)), np.zeros(K + 1)
    _lib.check(lib.pl_normal_eq_blocks(ptrs, rows, len(blocks), K, None, ne, 100.0, 1.0, _lib.dptr(b), 1, 0.01, 0, _lib.dptr(M3.reshape(-1)), _lib.dptr(v3)),
               "pl_normal_eq_blocks")
    check("normal_eq_blocks_M", normerr(M3, Mref), out)
    check("normal_eq_blocks_v", normerr(v3, vref), out)
    del A, blocks

    # ---- covariance: 40 000 x 1000, mean/sigma = 10 (kills one-pass formulas), centred, then with the persisted mean ----
    n, K = 40_000, 1000
    rng = np.random.default_rng(4)
    X = rng.standard_normal((n, K)) + 10.0 * rng.standard_normal(K)
    mref = X.sum(0) / n
    Xc = X - mref
    Cref = Xc.T @ Xc / n
    C, m = np.zeros((K, K)), np.zeros(K)
    _lib.check(lib.pl_cov(_lib.dptr(X), n, K, K, _lib.dptr(m), 0, 1, _lib.dptr(C.reshape(-1))), "pl_cov")
    check("cov_mean", float(np.max(np.abs(m - mref) / np.abs(mref))), out)
    check("cov_C", normerr(C, Cref), out)
    assert np.array_equal(C, C.T)
    C2 = np.zeros((K, K))
    _lib.check(lib.pl_cov(_lib.dptr(X), n, K, K, _lib.dptr(m), 1, 1, _lib.dptr(C2.reshape(-1))), "pl_cov (mean given)")
    check("cov_C_mean_given", normerr(C2, Cref), out)
    Cu = np.zeros((K, K))
    _lib.check(lib.pl_cov(_lib.dptr(X), n, K, K, None, 0, 0, _lib.dptr(Cu.reshape(-1))), "pl_cov (uncentred)")
    check("cov_uncentred", normerr(Cu, X.T @ X / n), out)
    del X, Xc

    # ---- kernel matrices: N = 9000 (>= 8192: tile-column ranges per device), Symmetric(:U) storage, against the oracle pair loop ----
    N, d = 9000, 40
    rng = np.random.default_rng(5)
    F = 2.0 * rng.standard_normal(d) + rng.standard_normal((N, d)) / np.sqrt(96.0)
    rows_s = np.sort(rng.choice(N, 48, replace=False))
    for name, call, ko in (
        ("rbf", lambda Kd: lib.pl_gram_rbf(_lib.dptr(F), N, d, d, 0, None, 1.0, 1.0, _lib.dptr(Kd.reshape(-1)), N, _lib.PL_FILL_JULIA_U), orc.RBF(orc.Euclidean(d))),
        ("dot", lambda Kd: lib.pl_gram_dot(_lib.dptr(F), N, d, d, 2, _lib.dptr(Kd.reshape(-1)), N, _lib.PL_FILL_JULIA_U), orc.DotProduct()),
    ):
        Kd = np.zeros((N, N))
        t0 = time.perf_counter()
        _lib.check(call(Kd), "pl_gram_" + name)
        out[f"gram_{name}_s"] = time.perf_counter() - t0
        assert np.array_equal(np.triu(Kd, 1), np.zeros((N, N))), "row-major lower == Julia's :U triangle only"
        ref = gram_rows_ref(orc, F, ko, rows_s)  # rows of the full symmetric matrix
        err = 0.0
        for q, i in enumerate(rows_s):  # stored: Kd[j][i] for j >= i  (column i of Julia's upper triangle)
            err = max(err, float(np.max(np.abs(Kd[i:, i] - ref[q, i:]) / np.abs(ref[q, i:]))))
            err = max(err, float(np.max(np.abs(Kd[i, : i + 1] - ref[q, : i + 1]) / np.abs(ref[q, : i + 1]))))
        check(f"gram_{name}_sampled_rows", err, out)
        if name == "dot":
            Kdot = Kd

    # ---- resident, tile-packed kernel matrix + block fetch (config 5's form) ----
    h = ctypes.c_void_p()
    _lib.check(lib.pl_gram_resident(_lib.PL_KERNEL_DOT, _lib.dptr(F), N, d, d, 2, 0, None, 1.0, 1.0, ctypes.byref(h)), "pl_gram_resident")
    nn, npart, tl, bl = ctypes.c_int64(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64()
    _lib.check(lib.pl_gram_resident_info(h, ctypes.byref(nn), ctypes.byref(npart), ctypes.byref(tl), ctypes.byref(bl)), "info")
    T = (N + 127) // 128
    assert nn.value == N and npart.value == nd and T * (T + 1) // 2 <= tl.value <= (T + 1) // 2 * (T + 1), (nn.value, npart.value, tl.value)  # slots of the folded enumeration
    full = np.tril(Kdot) + np.tril(Kdot, -1).T  # symmetric matrix from the :U storage
    worst = 0.0
    for (i0, i1, j0, j1) in [(0, 300, 0, 300), (100, 400, 5000, 5700), (8000, 9000, 200, 260), (8873, 9000, 8873, 9000), (4000, 4129, 3900, 4300), (17, 18, 0, 9000)]:
        blk = np.full((i1 - i0, j1 - j0), np.nan)
        _lib.check(lib.pl_gram_fetch_block(h, i0, i1, j0, j1, _lib.dptr(blk.reshape(-1)), j1 - j0), "pl_gram_fetch_block")
        worst = max(worst, float(np.max(np.abs(blk - full[i0:i1, j0:j1]))))
    check("gram_resident_fetch_maxabs_vs_dense", worst, out, tol=1e-300)  # the same kernel, the same bits
    _lib.check(lib.pl_gram_free(h), "pl_gram_free")
    _lib.finalize()
    out["ok"] = True
    print(json.dumps(out))


def run_ranks():
    import ctypes

    import torch
    import torch.distributed as dist

    import plb200
    from plb200 import _lib

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # control plane only: the data-path collectives are the library's own
    lib = _lib.init(local)
    _lib.comm_init_from_torch()
    out = {"mode": "ranks", "world": world, "nccl": _lib.comm_info()}
    assert out["nccl"]["nranks"] == world

    # ---- normal equations: every rank holds [its energy rows ; its force rows] ----
    R, K, ne = 60_000, 257, 203
    A, b = problem(11, R, K, ne)
    w2 = np.where(np.arange(R) < ne, 30.0, 1.0)
    Mref, vref = ref_normal_eq(A, w2, b, ne, True, 0.5)
    e0, e1 = rank * ne // world, (rank + 1) * ne // world
    f0, f1 = ne + rank * (R - ne) // world, ne + (rank + 1) * (R - ne) // world
    Al = np.ascontiguousarray(np.vstack([A[e0:e1], A[f0:f1]]))
    bl = np.concatenate([b[e0:e1], b[f0:f1]])
    M, v = np.zeros((K + 1, K + 1)), np.zeros(K + 1)
    _lib.check(lib.pl_normal_eq_sharded(_lib.dptr(Al), Al.shape[0], K, K, None, e1 - e0, 30.0, 1.0, _lib.dptr(bl), 1, 0.5, _lib.dptr(M.reshape(-1)), _lib.dptr(v)),
               "pl_normal_eq_sharded")
    check("sharded_normal_eq_M", normerr(M, Mref), out)
    check("sharded_normal_eq_v", normerr(v, vref), out)
    assert np.array_equal(M, M.T)
    # device-resident form on torch tensors
    dA, db = torch.as_tensor(Al, device="cuda"), torch.as_tensor(bl, device="cuda")
    dM, dv = torch.empty((K + 1, K + 1), device="cuda", dtype=torch.float64), torch.empty(K + 1, device="cuda", dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.pl_normal_eq_sharded_dev(dA.data_ptr(), Al.shape[0], K, K, None, e1 - e0, 30.0, 1.0, db.data_ptr(), 1, 0.5, dM.data_ptr(), dv.data_ptr(), st),
               "pl_normal_eq_sharded_dev")
    torch.cuda.synchronize()
    check("sharded_dev_normal_eq_M", normerr(dM.cpu().numpy(), Mref), out)
    out["coll_ms"] = _lib.last_coll_ms()
    # every rank holds the same bits
    g = [torch.empty_like(dM.cpu()) for _ in range(world)]
    dist.all_gather(g, dM.cpu())
    assert all(torch.equal(g[0], x) for x in g), "all ranks hold identical results"

    # ---- covariance ----
    n, K = 30_000, 300
    rng = np.random.default_rng(12)
    X = rng.standard_normal((n, K)) + 10.0 * rng.standard_normal(K)
    mref = X.sum(0) / n
    Cref = (X - mref).T @ (X - mref) / n
    lo, hi = rank * n // world, (rank + 1) * n // world
    Xl = np.ascontiguousarray(X[lo:hi])
    C, m = np.zeros((K, K)), np.zeros(K)
    _lib.check(lib.pl_cov_sharded(_lib.dptr(Xl), hi - lo, n, K, K, _lib.dptr(m), 0, 1, _lib.dptr(C.reshape(-1))), "pl_cov_sharded")
    check("sharded_cov_mean", float(np.max(np.abs(m - mref) / np.abs(mref))), out)
    check("sharded_cov_C", normerr(C, Cref), out)
    assert np.array_equal(C, C.T)
    dX = torch.as_tensor(Xl, device="cuda")
    dC, dm = torch.empty((K, K), device="cuda", dtype=torch.float64), torch.empty(K, device="cuda", dtype=torch.float64)
    _lib.check(lib.pl_cov_sharded_dev(dX.data_ptr(), hi - lo, n, K, K, dm.data_ptr(), 0, 1, dC.data_ptr(), st), "pl_cov_sharded_dev")
    torch.cuda.synchronize()
    check("sharded_dev_cov_C", normerr(dC.cpu().numpy(), Cref), out)

    # ---- resident kernel matrix: this rank's tiles, block fetch is collective ----
    N, d = 3000, 26
    rng = np.random.default_rng(13)
    F = 2.0 * rng.standard_normal(d) + rng.standard_normal((N, d))
    G = F @ F.T
    nrm = np.sqrt(np.diag(G))
    Kref = (G / np.outer(nrm, nrm)) ** 2
    h = ctypes.c_void_p()
    _lib.check(lib.pl_gram_resident(_lib.PL_KERNEL_DOT, _lib.dptr(F), N, d, d, 2, 0, None, 1.0, 1.0, ctypes.byref(h)), "pl_gram_resident")
    worst = 0.0
    for (i0, i1, j0, j1) in [(0, 3000, 0, 3000), (2900, 3000, 10, 500), (128, 256, 128, 256)]:
        blk = np.full((i1 - i0, j1 - j0), np.nan)
        _lib.check(lib.pl_gram_fetch_block(h, i0, i1, j0, j1, _lib.dptr(blk.reshape(-1)), j1 - j0), "pl_gram_fetch_block")
        worst = max(worst, float(np.max(np.abs(blk - Kref[i0:i1, j0:j1]) / Kref[i0:i1, j0:j1])))
    check("resident_fetch_vs_numpy", worst, out)
    _lib.check(lib.pl_gram_free(h), "pl_gram_free")
    _lib.comm_destroy()
    dist.barrier()
    if rank == 0:
        out["ok"] = True
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "single"
    if mode == "single":
        run_single(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    else:
        run_ranks()
