"""RBF epilogue probe: config 2 (RBF, N = 20000, d = 300, Symmetric(:U) storage and both triangles) and other shapes with the register
epilogue (PLB200_GRAM_TMAST=0) and the staged TMA-store epilogue (=1), DotProduct beside it. Each setting runs in a fresh process."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "potentiallearning.jl_b200"))
os.environ["PLB200_NO_GRAPH"] = "1"
import numpy as np, torch, plb200
from plb200 import _lib
plb200.init(0); lib = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(kid, n, d, fill, reps=8):
    rng = np.random.default_rng(1002)
    X = torch.as_tensor(2.0 * rng.standard_normal(d) + rng.standard_normal((n, d)) / np.sqrt(96.0), device="cuda")
    K = torch.zeros((n, n), device="cuda", dtype=torch.float64)
    ts = []
    for r in range(reps + 2):
        flush.fill_(r & 1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.pl_gram_dev(kid, X.data_ptr(), n, d, None, 0, 0, d, 2, 0, None, 1.0, 1.0, K.data_ptr(), n, fill, 0, 1, 0, st), "gram")
        e1.record(); torch.cuda.synchronize()
        if r >= 2: ts.append(e0.elapsed_time(e1))
    return {"ms_median": float(np.median(ts)), "ms_min": float(np.min(ts)), "checksum": float(K.sum().item())}
def run_packed(kid, n, d, nparts, reps=8):
    rng = np.random.default_rng(1002)
    X = torch.as_tensor(2.0 * rng.standard_normal(d) + rng.standard_normal((n, d)) / np.sqrt(96.0), device="cuda")
    nt = int(lib.pl_gram_tile_count(n, 0, 0, nparts))
    K = torch.zeros((nt, 128, 128), device="cuda", dtype=torch.float64)
    ts = []
    for r in range(reps + 2):
        flush.fill_(r & 1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.pl_gram_dev(kid, X.data_ptr(), n, d, None, 0, 0, d, 2, 0, None, 1.0, 1.0, K.data_ptr(), 128, _lib.PL_FILL_FULL, 0, nparts, 1, st), "gram")
        e1.record(); torch.cuda.synchronize()
        if r >= 2: ts.append(e0.elapsed_time(e1))
    return {"ms_median": float(np.median(ts)), "ms_min": float(np.min(ts)), "tiles": nt, "checksum": float(K.sum().item())}
out["rbf_c2_packed_1part"] = run_packed(_lib.PL_KERNEL_RBF, 20000, 300, 1)
out["rbf_c2_packed_8part"] = run_packed(_lib.PL_KERNEL_RBF, 20000, 300, 8)
out["dot_c2_packed_1part"] = run_packed(_lib.PL_KERNEL_DOT, 20000, 300, 1)
out["rbf_c2_U"] = run(_lib.PL_KERNEL_RBF, 20000, 300, _lib.PL_FILL_JULIA_U)
out["rbf_c2_full"] = run(_lib.PL_KERNEL_RBF, 20000, 300, _lib.PL_FILL_FULL)
out["rbf_n8000_d26_U"] = run(_lib.PL_KERNEL_RBF, 8000, 26, _lib.PL_FILL_JULIA_U)
out["rbf_n1000_d26_U"] = run(_lib.PL_KERNEL_RBF, 1000, 26, _lib.PL_FILL_JULIA_U, reps=20)
out["dot_c2_U"] = run(_lib.PL_KERNEL_DOT, 20000, 300, _lib.PL_FILL_JULIA_U)
print(json.dumps(out))
''' % (ROOT, ROOT)
res = {}
for lead in sys.argv[1:] or ["0", "1"]:
    env = dict(os.environ, PLB200_GRAM_TMAST=lead)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    try:
        res[lead] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        res[lead] = {"error": r.stderr[-800:]}
print(json.dumps(res, indent=1))
