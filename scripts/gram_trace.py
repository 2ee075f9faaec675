"""Per-CTA timeline of the Gram kernel (developer build with -DPLB_GRAM_TRACE, see the comment in gram.cuh): where the fixed ~40 us per
launch go. Build: nvcc ... -DPLB_GRAM_TRACE -o scripts/native/libplb200_trace.so potentiallearning.jl_b200/csrc/plb200_api.cu"""
import os
import sys

import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "scripts", "native", "libplb200_trace.so")
SRC = os.path.join(ROOT, "potentiallearning.jl_b200", "csrc", "plb200_api.cu")
if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-pthread", "-shared",
                           "-ldl", "-lpthread", "-DPLB_GRAM_TRACE", "-o", SO, SRC])
os.environ["PLB200_LIB"] = SO
os.environ["PLB200_NO_GRAPH"] = "1"
TRACE = os.path.join(ROOT, "gpurun_out", "gram_trace.txt")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import numpy as np
import torch

import plb200
from plb200 import _lib

plb200.init(0)
lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
N, D = 20000, 300
rng = np.random.default_rng(1002)
X = torch.as_tensor(2.0 * rng.standard_normal(D) + rng.standard_normal((N, D)) / np.sqrt(96.0), device="cuda")
for nparts in (8, 1):
    nt = int(lib.pl_gram_tile_count(N, 0, 0, nparts))
    out = torch.empty((nt, 128, 128), device="cuda", dtype=torch.float64)
    os.environ["PLB200_GRAM_NO_TAIL"] = "1"
    os.environ.pop("PLB200_GRAM_TRACE", None)
    for _ in range(3):
        _lib.check(lib.pl_gram_dev(_lib.PL_KERNEL_RBF, X.data_ptr(), N, D, None, 0, 0, D, 0, 0, None, 1.0, 1.0, out.data_ptr(), 128, _lib.PL_FILL_FULL, 0, nparts, 1, st), "g")
    torch.cuda.synchronize()
    if os.path.exists(TRACE):
        os.remove(TRACE)
    os.environ["PLB200_GRAM_TRACE"] = TRACE
    _lib.check(lib.pl_gram_dev(_lib.PL_KERNEL_RBF, X.data_ptr(), N, D, None, 0, 0, D, 0, 0, None, 1.0, 1.0, out.data_ptr(), 128, _lib.PL_FILL_FULL, 0, nparts, 1, st), "g")
    torch.cuda.synchronize()
    rows = [list(map(int, l.split())) for l in open(TRACE) if not l.startswith("launch")]
    a = np.array(rows, dtype=np.int64)
    t0 = a[:, 1].min()
    ent, wait, m1, e1, end, sm, nd = (a[:, k] for k in range(1, 8))
    us = lambda v: np.round((v - t0) / 1e3, 2)
    print(f"nparts {nparts}: {len(a)} CTAs, tiles per CTA {nd.min()}..{nd.max()}")
    print("  entry            us after the first entry: min / median / max", us(ent).min(), np.median(us(ent)), us(ent).max())
    print("  dependency wait over                    :", us(wait).min(), np.median(us(wait)), us(wait).max())
    print("  first tile: main loop done              :", us(m1).min(), np.median(us(m1)), us(m1).max())
    print("  first tile: epilogue done               :", us(e1).min(), np.median(us(e1)), us(e1).max())
    print("  last tile done                          :", us(end).min(), np.median(us(end)), us(end).max())
    per_tile = (end - e1) / np.maximum(nd - 1, 1) / 1e3
    print("  us per tile after the first             :", round(float(per_tile.min()), 2), round(float(np.median(per_tile)), 2), round(float(per_tile.max()), 2))
    print("  first tile takes (wait over -> epilogue done):", round(float(np.median(e1 - wait)) / 1e3, 2))
    order = np.argsort(per_tile)
    print("  slowest CTAs (block, sm, tiles, us/tile):", [(int(a[i, 0]), int(sm[i]), int(nd[i]), round(float(per_tile[i]), 2)) for i in order[-12:]])
    print("  fastest CTAs (block, sm, tiles, us/tile):", [(int(a[i, 0]), int(sm[i]), int(nd[i]), round(float(per_tile[i]), 2)) for i in order[:8]])
    # by SM id parity / range: B200 has two dies; SM ids 0..73 / 74..147 is a guess at the split
    for name, sel in (("sm < 74", sm < 74), ("sm >= 74", sm >= 74), ("block < 74", a[:, 0] < 74), ("block >= 74", a[:, 0] >= 74)):
        print("   ", name, "median us/tile", round(float(np.median(per_tile[sel])), 2), "max", round(float(per_tile[sel].max()), 2))
    np.save(os.path.join(ROOT, "gpurun_out", f"gram_trace_parts{nparts}.npy"), a)
