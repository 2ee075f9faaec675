"""GPU probe of the widened rows: feature-reduction bandwidth, streamed normal equations end to end
from pinned host memory, eigendecomposition timings. CUDA-event timings, results as one JSON object."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import numpy as np
import torch

import plb200
from plb200 import _lib

plb200.init(0)
out = {}


def timeit(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


which = sys.argv[1:] or ["feat", "neq", "eig", "proj", "dpp"]

if "feat" in which:
    # config-2 shape: 20 000 configurations x 96 atoms x 300 descriptors = 4.6 GB of local descriptors
    for N, nat, d in ((20000, 96, 300), (1000, 32, 26), (200000, 8, 500)):
        g = torch.Generator(device="cuda").manual_seed(1)
        L = torch.randn((N * nat, d), device="cuda", dtype=torch.float64, generator=g)
        off = (torch.arange(N + 1, device="cuda", dtype=torch.int64) * nat).contiguous()
        X = torch.empty((N, d), device="cuda", dtype=torch.float64)
        best, med = timeit(lambda: plb200.global_features_dev(L, off, True, out=X))
        ref = L.view(N, nat, d).sum(1) / nat
        out[f"feat_{N}x{nat}x{d}"] = {"ms": best, "GBs": (L.numel() + X.numel()) * 8 / best / 1e6,
                                      "max_abs_vs_torch": float((X - ref).abs().max())}
        del L, X, ref

if "neq" in which:
    # config 3 end to end from pinned host memory, 1/4 size by default (PLB_NEQ_FRAC=1 for all 11.56 GB)
    frac = float(os.environ.get("PLB_NEQ_FRAC", "0.25"))
    ncfg, nat, K = int(10000 * frac), 96, 500
    R = ncfg * (1 + 3 * nat)
    A = torch.empty((R, K), dtype=torch.float64).pin_memory()
    gen = torch.Generator().manual_seed(3)
    step = 1 << 18
    for r0 in range(0, R, step):
        A[r0:r0 + step].normal_(generator=gen)
    b = torch.randn(R, dtype=torch.float64, generator=gen).pin_memory()
    n = K + 1
    M = torch.empty((n, n), dtype=torch.float64).pin_memory()
    v = torch.empty(n, dtype=torch.float64).pin_memory()
    lib = _lib.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    P = lambda t: ctypes.cast(t.data_ptr(), dp)

    def run():
        _lib.check(lib.pl_normal_eq(P(A), R, K, K, None, ncfg, 100.0, 1.0, P(b), 1, 0.01, P(M), P(v)), "pl_normal_eq")

    ts = []
    for i in range(4):
        t0 = time.perf_counter()
        run()
        ts.append((time.perf_counter() - t0) * 1e3)
    out["neq_stream_pinned"] = {"R": R, "K": K, "GB": R * K * 8 / 1e9, "ms_best": min(ts[1:]), "ms_all": ts,
                                "rows_per_s": R / (min(ts[1:]) / 1e3), "h2d_GBs": R * K * 8 / (min(ts[1:]) / 1e3) / 1e9,
                                "timing": plb200.last_timing()}
    # block-list form: the same rows as ncfg+1 blocks, as lp.B / lp.dB[i] lie in memory
    An = A.numpy()
    blocks = [An[:ncfg]] + [An[ncfg + i * 3 * nat: ncfg + (i + 1) * 3 * nat] for i in range(ncfg)]
    t0 = time.perf_counter()
    M2, v2 = plb200.normal_equations_blocks(blocks, b.numpy(), None, ncfg, 100.0, 1.0, True, 0.01, layout="rows")
    out["neq_blocks_pinned"] = {"nblocks": len(blocks), "ms": (time.perf_counter() - t0) * 1e3,
                                "max_rel_vs_contiguous": float(np.max(np.abs(M2 - M.numpy())) / np.max(np.abs(M.numpy())))}
    # the same rows in PAGEABLE memory (a plain numpy / Julia array): staged by the library (hoststage.h) vs the driver's bounce buffer
    Apg = np.array(An, copy=True)
    bpg = np.array(b.numpy(), copy=True)
    Mpg, vpg = np.empty((n, n)), np.empty(n)
    PP = lambda a: a.ctypes.data_as(dp)

    def run_pg():
        _lib.check(lib.pl_normal_eq(PP(Apg), R, K, K, None, ncfg, 100.0, 1.0, PP(bpg), 1, 0.01, PP(Mpg), PP(vpg)), "pl_normal_eq")

    for label, env in (("staged", None), ("unstaged", "1")):
        if env:
            os.environ["PLB200_NO_STAGING"] = env
        ts = []
        for i in range(3):
            t0 = time.perf_counter()
            run_pg()
            ts.append((time.perf_counter() - t0) * 1e3)
        os.environ.pop("PLB200_NO_STAGING", None)
        out[f"neq_stream_pageable_{label}"] = {"ms_best": min(ts), "h2d_GBs": R * K * 8 / (min(ts) / 1e3) / 1e9,
                                               "bitwise_equal_to_pinned": bool(np.array_equal(Mpg, M.numpy()))}
    del A, Apg

if "eig" in which:

    for n in (1000, 4000) + ((20000,) if os.environ.get("PLB_EIG_BIG") else ()):
        rng = np.random.default_rng(n)
        X = rng.standard_normal((n, 300)) / np.sqrt(96) + rng.standard_normal(300)
        k = plb200.RBF(plb200.Euclidean(300))
        t0 = time.perf_counter()
        ell = plb200.EllEnsemble.from_features(X, k, keep_L=False)
        t_first = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        ell = plb200.EllEnsemble.from_features(X, k, keep_L=False)
        t_dev = (time.perf_counter() - t0) * 1e3
        rec = {"ell_ensemble_ms_first": t_first, "ell_ensemble_ms": t_dev, "timing": plb200.last_timing()}
        t0 = time.perf_counter()
        plb200.rescale_(ell, n // 10)
        p = plb200.inclusion_prob(ell)
        rec["inclusion_prob_ms"] = (time.perf_counter() - t0) * 1e3
        rec["inclusion_sum_err"] = float(abs(p.sum() - n // 10))
        if n <= 4000:
            K = np.asarray(plb200.KernelMatrix(X, k))
            t0 = time.perf_counter()
            lo = np.linalg.eigvalsh(K)
            rec["numpy_eigvalsh_ms"] = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            lo, Uo = np.linalg.eigh(K)
            rec["numpy_eigh_ms"] = (time.perf_counter() - t0) * 1e3
            rec["lambda_max_rel_err"] = float(np.max(np.abs(ell.λ - np.maximum(lo, 0))) / lo.max())
        out[f"eig_{n}"] = rec

if "proj" in which:
    # transform!-shaped projections: 2.5 M rows x K = 1000 onto p directions, device-resident (the LINEAR Gram epilogue)
    from plb200.kernels import _Linear, kernel_matrix_dev

    g = torch.Generator(device="cuda").manual_seed(5)
    Rp, Kp = 2_500_000, 1000
    Xd = torch.randn((Rp, Kp), device="cuda", dtype=torch.float64, generator=g)
    for pdim in (8, 32, 64, 96, 128):
        Wt = torch.randn((pdim, Kp), device="cuda", dtype=torch.float64, generator=g)
        Yd = torch.empty((Rp, pdim + (pdim & 1)), device="cuda", dtype=torch.float64)[:, :pdim]
        best, med = timeit(lambda: kernel_matrix_dev(Xd, _Linear(), out=Yd, X2=Wt), 1, 3)
        ref = Xd[:4096] @ Wt.T
        out[f"project_{Rp}x{Kp}_p{pdim}"] = {"ms": best, "read_GBs": Rp * Kp * 8 / best / 1e6, "tflops": 2.0 * Rp * Kp * pdim / best / 1e9,
                                             "max_abs_err_4096_rows": float((Yd[:4096] - ref).abs().max())}
    del Xd

if "dpp" in which:
    # kDPP end to end at config-2 size: features -> RBF kernel matrix -> eigendecomposition (device) -> rescale -> k-DPP draw
    n = 20000
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, 300)) / np.sqrt(96) + rng.standard_normal(300)
    t0 = time.perf_counter()
    ell = plb200.EllEnsemble.from_features(X, plb200.RBF(plb200.Euclidean(300)), keep_L=False)
    rec = {"ell_ensemble_ms": (time.perf_counter() - t0) * 1e3}
    for k in [int(x) for x in os.environ.get("PLB_DPP_K", "500,2000").split(",")]:
        t0 = time.perf_counter()
        plb200.rescale_(ell, k)
        t1 = time.perf_counter()
        items = plb200.sample(ell, k, rng=np.random.default_rng(1))
        t2 = time.perf_counter()
        rec[f"k{k}"] = {"rescale_ms": (t1 - t0) * 1e3, "sample_ms": (t2 - t1) * 1e3, "device_ms": plb200.last_timing()["kernel_ms"],
                        "distinct": len(set(items)) == k}
        t0 = time.perf_counter()
        mode = plb200.greedy_subset(ell, k)
        rec[f"k{k}"]["greedy_map_ms"] = (time.perf_counter() - t0) * 1e3
        rec[f"k{k}"]["greedy_items"] = len(set(mode))
    out["kdpp_n20000"] = rec

out["launches"] = plb200.launch_count()
print(json.dumps(out))
