"""Functional check of the sharded paths on real GPUs (run under torchrun, one process per GPU): the N-GPU results of the normal
equations, the centred covariance (one-pass shifted form) and the tile-sharded Gram matrix must equal the single-GPU results computed
on rank 0 from the same data, to rounding; S must be exactly symmetric and bitwise identical on every rank. Prints one JSON line."""
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

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
plb200.init(local_rank)
eng = DeviceEngine()
res = {"n_gpus": world}


def same_on_all_ranks(t):
    if world == 1:
        return True
    ref = t.clone()
    dist.broadcast(ref, 0)
    ok = torch.tensor([1 if torch.equal(ref, t) else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return bool(ok.item())


# ---- normal equations: 4000 configurations x 12 atoms, K = 143 (odd shapes on purpose), every rank generates the same full data ----
g = torch.Generator(device=dev).manual_seed(7)
ncfg, nat, K = 4000, 12, 143
B = torch.randn((ncfg, K), device=dev, dtype=torch.float64, generator=g) * 3
dB = torch.randn((ncfg * 3 * nat, K), device=dev, dtype=torch.float64, generator=g)
e = torch.randn(ncfg, device=dev, dtype=torch.float64, generator=g)
f = torch.randn(ncfg * 3 * nat, device=dev, dtype=torch.float64, generator=g)
c0, c1 = shard_rows(ncfg, rank, world)  # whole configurations per rank
A_loc = torch.cat([B[c0:c1], dB[c0 * 3 * nat:c1 * 3 * nat]]).contiguous()
b_loc = torch.cat([e[c0:c1], f[c0 * 3 * nat:c1 * 3 * nat]]).contiguous()
M, v = normal_equations_sharded(A_loc, b_loc, None, c1 - c0, 100.0, 1.0, True, 0.01, engine=eng)
A_all = torch.cat([B, dB]).contiguous()
b_all = torch.cat([e, f]).contiguous()
S1, vec1 = eng.syrk_partial(A_all, None, ncfg, 100.0, 1.0, b_all, None, True)
M1, v1 = eng.normal_eq_finish(S1, vec1, True, 0.01)
res["neq_rel_err_vs_1gpu"] = float((M - M1).abs().max() / M1.abs().max())
res["neq_rhs_rel_err_vs_1gpu"] = float((v - v1).abs().max() / v1.abs().max())
res["neq_symmetric"] = bool(torch.equal(M, M.T))
res["neq_identical_on_ranks"] = same_on_all_ranks(M) and same_on_all_ranks(v)

# ---- centred covariance, mean/sigma = 10 ----
n, Kc = 300_000, 129
X = torch.randn((n, Kc), device=dev, dtype=torch.float64, generator=g) + 10.0 * torch.randn(Kc, device=dev, dtype=torch.float64, generator=g)
r0, r1 = shard_rows(n, rank, world)
C, m = covariance_sharded(X[r0:r1].contiguous(), n, engine=eng)
m_ref = X.sum(0) / n
Y = X - m_ref
C_ref = (Y.T @ Y) / n
res["cov_rel_err_vs_torch"] = float((C - C_ref).abs().max() / C_ref.abs().max())
res["cov_mean_rel_err"] = float(((m - m_ref).abs() / m_ref.abs()).max())
res["cov_symmetric"] = bool(torch.equal(C, C.T))
res["cov_identical_on_ranks"] = same_on_all_ranks(C)

# ---- tile-sharded Gram: every tile is owned by exactly one rank and equals the dense single-GPU result ----
ng, d = 1500, 40
Xg = (torch.randn((ng, d), device=dev, dtype=torch.float64, generator=g) * 0.3 + torch.randn(d, device=dev, dtype=torch.float64, generator=g)).contiguous()
kern = plb200.RBF(plb200.Euclidean(d), ell=1.3)
tiles = plb200.kernel_matrix_dev(Xg, kern, part=rank, nparts=world, packed=True)
dense = plb200.kernel_matrix_dev(Xg, kern, fill=_lib.PL_FILL_FULL)
coords = plb200.gram_tile_coords(ng, 0, rank, world)
worst, owned = 0.0, 0
for k, (ti, tj) in enumerate(coords):
    if ti < 0:
        continue
    owned += 1
    i1, j1 = min((ti + 1) * 128, ng), min((tj + 1) * 128, ng)
    blk = dense[ti * 128:i1, tj * 128:j1]
    got = tiles[k][: i1 - ti * 128, : j1 - tj * 128]
    if ti == tj:  # diagonal tiles hold the j >= i half (the other half is its mirror)
        worst = max(worst, float((torch.triu(got) - torch.triu(blk)).abs().max()))
    else:
        worst = max(worst, float((got - blk).abs().max()))
cnt = torch.tensor([owned], device=dev)
w = torch.tensor([worst], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(cnt)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
T = (ng + 127) // 128
res["gram_tiles_owned_total"] = int(cnt.item())
res["gram_tiles_expected"] = T * (T + 1) // 2
res["gram_max_abs_diff_vs_dense"] = float(w.item())

ok = (res["neq_rel_err_vs_1gpu"] < 1e-12 and res["neq_rhs_rel_err_vs_1gpu"] < 1e-12 and res["neq_symmetric"] and res["neq_identical_on_ranks"]
      and res["cov_rel_err_vs_torch"] < 1e-11 and res["cov_mean_rel_err"] < 1e-12 and res["cov_symmetric"] and res["cov_identical_on_ranks"]
      and res["gram_tiles_owned_total"] == res["gram_tiles_expected"] and res["gram_max_abs_diff_vs_dense"] == 0.0)
res["ok"] = bool(ok)
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"multi_gpu_check_n{world}.json"), "w"), indent=1)
    sys.stderr.write(json.dumps(res) + "\n")
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
