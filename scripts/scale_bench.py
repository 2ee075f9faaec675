"""Per-configuration throughput of the hot path at N GPUs (one process per GPU; run under torchrun for N > 1):
  C2  RBF Gram      N=20 000  d=300   tiles round-robin, no collective
  C3  normal eqs    10 000 configs x 96 atoms (R = 2 890 000), K=500, weighted, intercept; rows sharded + ONE all-reduce
  C4  PCAState cov  n=10^7 x K=1000 (80 GB) rows sharded, column-sum all-reduce + ONE all-reduce   [--c4-rows to scale down]
  C5  DotProduct Gram N=200 000 d=500, tile-packed shards (20 GB/GPU at 8 GPUs)                     [needs >= 160 GB/N per GPU]
Prints one JSON object (rank 0). All times are CUDA-event times on the launching stream, max over ranks."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import numpy as np
import torch
import torch.distributed as dist

import plb200
from plb200 import _lib
from plb200.device import DeviceEngine
from plb200.distributed import covariance_sharded, normal_equations_sharded, shard_rows

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="c2,c3,c4,c5")
ap.add_argument("--c4-rows", type=int, default=10_000_000)
ap.add_argument("--c5-n", type=int, default=200_000)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ["NCCL_DEBUG"] = "WARN"  # (NCCL still prints its version banner on stdout: the result is also written to gpurun_out/scale_n<N>.json)
    dist.init_process_group("nccl", device_id=dev)
plb200.init(local_rank)
eng = DeviceEngine()
out = {"n_gpus": world}


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=args.reps, warm=2):
    for _ in range(warm):
        fn()
    barrier()
    ts = []
    for _ in range(reps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return min(ts), float(np.median(ts))


def randn_rows(n, k, gen, scale=1.0, shift=None, chunk=1 << 20):
    X = torch.empty((n, k), device=dev, dtype=torch.float64)
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        X[r0:r1].normal_(generator=gen)
        if scale != 1.0:
            X[r0:r1] *= scale
        if shift is not None:
            X[r0:r1] += shift
    return X


cfgs = args.configs.split(",")
gen = torch.Generator(device=dev).manual_seed(1000 + rank)
gen_shared = torch.Generator(device=dev).manual_seed(999)

if "c2" in cfgs:
    N, d = 20000, 300
    mu = 2.0 * torch.randn(d, device=dev, dtype=torch.float64, generator=gen_shared)
    X = mu + torch.randn((N, d), device=dev, dtype=torch.float64, generator=gen_shared) / np.sqrt(96.0)
    cnt = int(_lib.load().pl_gram_tile_count(N, 0, rank, world))
    tiles = torch.empty((cnt, 128, 128), device=dev, dtype=torch.float64)
    k = plb200.RBF(plb200.Euclidean(d))
    best, med = timed(lambda: plb200.kernel_matrix_dev(X, k, out=tiles, part=rank, nparts=world, packed=True))
    out["c2_rbf_gram"] = {"ms": best, "ms_median": med, "tflops": N * (N + 1) * d / best / 1e9, "shape": [N, d]}
    del X, tiles

if "c3" in cfgs:
    ncfg, natoms, K = 10000, 96, 500
    c0, c1 = shard_rows(ncfg, rank, world)
    ne = c1 - c0
    nf = ne * 3 * natoms
    A = randn_rows(ne + nf, K, gen)
    A[:ne] *= np.sqrt(96.0)
    b = torch.randn(ne + nf, device=dev, dtype=torch.float64, generator=gen)
    R = ncfg * (1 + 3 * natoms)
    best, med = timed(lambda: normal_equations_sharded(A, b, None, ne, 100.0, 1.0, True, 0.01, engine=eng))
    out["c3_normal_eq"] = {"ms": best, "ms_median": med, "rows_per_s": R / best * 1e3, "tflops": R * K * (K + 1) / best / 1e9, "shape": [R, K],
                           "rows_this_rank": ne + nf}
    # kernel only (no all-reduce / finish)
    best_k, _ = timed(lambda: eng.syrk_partial(A, None, ne, 100.0, 1.0, b, None, True))
    out["c3_normal_eq"]["syrk_only_ms"] = best_k
    M, v = normal_equations_sharded(A, b, None, ne, 100.0, 1.0, True, 0.01, engine=eng)
    out["c3_normal_eq"]["symmetric"] = bool(torch.equal(M, M.T))
    if world > 1:  # every rank must hold the identical result
        M0 = M.clone()
        dist.broadcast(M0, 0)
        out["c3_normal_eq"]["identical_on_ranks"] = bool(torch.equal(M, M0))
    del A, b

if "c4" in cfgs:
    n, K = args.c4_rows, 1000
    r0, r1 = shard_rows(n, rank, world)
    shift = 10.0 * torch.randn(K, device=dev, dtype=torch.float64, generator=gen_shared)
    X = randn_rows(r1 - r0, K, gen, shift=shift)
    best, med = timed(lambda: covariance_sharded(X, n, engine=eng), reps=3, warm=1)
    out["c4_pca_cov"] = {"ms": best, "ms_median": med, "tflops": n * K * (K + 1) / best / 1e9, "rows_per_s": n / best * 1e3, "shape": [n, K],
                         "gb_this_rank": (r1 - r0) * K * 8 / 1e9}
    C, _ = covariance_sharded(X, n, engine=eng)
    out["c4_pca_cov"]["symmetric"] = bool(torch.equal(C, C.T))
    del C
    best_c, _ = timed(lambda: eng.colsum(X), reps=3, warm=1)
    out["c4_pca_cov"]["colsum_ms"] = best_c
    out["c4_pca_cov"]["colsum_GBs"] = (r1 - r0) * K * 8 / best_c / 1e6
    del X

if "c5" in cfgs:
    torch.cuda.empty_cache()
    N, d = args.c5_n, 500
    mu = 2.0 * torch.randn(d, device=dev, dtype=torch.float64, generator=gen_shared)
    X = mu + torch.randn((N, d), device=dev, dtype=torch.float64, generator=gen_shared) / np.sqrt(96.0)
    cnt = int(_lib.load().pl_gram_tile_count(N, 0, rank, world))
    need_gb = cnt * 128 * 128 * 8 / 1e9
    free_gb = torch.cuda.mem_get_info(dev)[0] / 1e9
    if need_gb > free_gb - 4:
        out["c5_dot_gram"] = {"skipped": f"shard needs {need_gb:.0f} GB, {free_gb:.0f} GB free on this GPU"}
    else:
        tiles = torch.empty((cnt, 128, 128), device=dev, dtype=torch.float64)
        k = plb200.DotProduct()
        best, med = timed(lambda: plb200.kernel_matrix_dev(X, k, out=tiles, part=rank, nparts=world, packed=True), reps=3, warm=1)
        out["c5_dot_gram"] = {"ms": best, "ms_median": med, "tflops": N * (N + 1) * d / best / 1e9, "shape": [N, d], "shard_gb": need_gb}
        # parity on sampled tiles against a float64 torch reference of the same formula (the oracle pair loop is checked in tests/)
        coords = plb200.gram_tile_coords(N, 0, rank, world)
        rn = 1.0 / torch.sqrt((X * X).sum(1))
        worst = 0.0
        for kk in np.linspace(0, cnt - 1, 8).astype(int):
            ti, tj = coords[kk]
            if ti < 0:
                continue
            i0, j0 = 128 * ti, 128 * tj
            Xi, Xj = X[i0:i0 + 128], X[j0:j0 + 128]
            ref = ((Xi @ Xj.T) * rn[i0:i0 + 128, None] * rn[None, j0:j0 + 128]) ** 2
            got = tiles[kk][: ref.shape[0], : ref.shape[1]]
            if ti == tj:
                ref.fill_diagonal_(1.0)
            worst = max(worst, float(((got - ref).abs() / ref.abs()).max()))
        out["c5_dot_gram"]["max_rel_err_sampled_tiles_vs_torch_fp64"] = worst
        del tiles
    del X

# ---- CPU legs on the box's host cores (rank 0, single-GPU run only): the reference's own formulation on a bounded sample ----
if world == 1 and rank == 0 and "cpu" in cfgs:
    import time

    from oracle import oracle as orc

    orc.build_c()
    ncores = os.cpu_count()
    rng = np.random.default_rng(7)
    # C3: A'*(Q*A) exactly as linear-learn.jl:253 forms it (Q*A materialised, one BLAS gemm on all cores), 1/20 of the rows
    Rs, K = 144_500, 500
    A = rng.standard_normal((Rs, K))
    q = np.concatenate([100.0 * np.ones(500), np.ones(Rs - 500)])
    t0 = time.perf_counter()
    M = A.T @ (q[:, None] * A)
    dt = time.perf_counter() - t0
    out["c3_cpu_reference_shaped"] = {"rows_per_s": Rs / dt, "tflops(RK(K+1))": Rs * K * (K + 1) / dt / 1e12, "cores": ncores, "kind": "port",
                                      "sample": f"{Rs} of 2 890 000 rows: Q*A materialised then OpenBLAS dgemm A'*(QA) (numpy), {dt:.2f} s"}
    # C4: the reference's per-row outer-product sum (dimension_reduction.jl:12), single thread, 1 500 rows x K = 1000
    n4, K4 = 1500, 1000
    X = 10.0 * rng.standard_normal(K4) + rng.standard_normal((n4, K4))
    t0 = time.perf_counter()
    orc.second_moment(X, use_c=True)
    dt = time.perf_counter() - t0
    out["c4_cpu_reference_shaped"] = {"rows_per_s": n4 / dt, "tflops(nK(K+1))": n4 * K4 * (K4 + 1) / dt / 1e12, "cores": 1, "kind": "port",
                                      "sample": f"{n4} of 10^7 rows: sequential sum of K x K outer products (oracle/pl_oracle.c orc_outer_sum), {dt:.2f} s; "
                                                "the reference form allocates one K x K temporary per row (80 TB at full size)"}
    t0 = time.perf_counter()
    Xc = X - X.mean(0)
    C = Xc.T @ Xc / n4
    dt = time.perf_counter() - t0
    out["c4_cpu_strong_blas"] = {"rows_per_s": n4 / dt, "cores": ncores, "sample": f"numpy two-pass centred X'X/n on {n4} rows, {dt:.3f} s"}

out["launches"] = plb200.launch_count()
if rank == 0:
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"scale_n{world}.json"), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
