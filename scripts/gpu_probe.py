"""GPU probe: fp64 denominators (cuBLAS DGEMM via torch.matmul) and first timings of the three kernel families."""
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
from plb200.device import DeviceEngine

plb200.init(0)
out = {}


def timeit(fn, warm=3, reps=10):
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


# cuBLAS DGEMM denominator
n = 8192
a = torch.randn((n, n), device="cuda", dtype=torch.float64)
b = torch.randn((n, n), device="cuda", dtype=torch.float64)
c = torch.empty((n, n), device="cuda", dtype=torch.float64)
best, med = timeit(lambda: torch.matmul(a, b, out=c), 3, 10)
out["dgemm_8192_tflops_burst"] = 2 * n ** 3 / best / 1e9
out["dgemm_8192_tflops_median"] = 2 * n ** 3 / med / 1e9
# sustained: back to back for ~3 s
torch.cuda.synchronize()
t0 = time.time()
cnt = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(5):
        torch.matmul(a, b, out=c)
    cnt += 5
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
out["dgemm_8192_tflops_sustained"] = cnt * 2 * n ** 3 / e0.elapsed_time(e1) / 1e9
# syrk-shaped cuBLAS: A^T A with R=289000, K=500
del a, b, c

# config 2: RBF gram N=20000 d=300
N, d = 20000, 300
g = torch.Generator(device="cuda").manual_seed(1002)
mu = 2.0 * torch.randn(d, device="cuda", dtype=torch.float64, generator=g)
X = mu + torch.randn((N, d), device="cuda", dtype=torch.float64, generator=g) / np.sqrt(96.0)
K = torch.empty((N, N), device="cuda", dtype=torch.float64)
for name, k in (("rbf", plb200.RBF(plb200.Euclidean(d))), ("dot", plb200.DotProduct())):
    for fill, fname in ((3, "full"), (2, "juliaU")):
        best, med = timeit(lambda: plb200.kernel_matrix_dev(X, k, out=K, fill=fill), 3, 10)
        out[f"c2_{name}_{fname}_ms_best"] = best
        out[f"c2_{name}_{fname}_ms_med"] = med
        out[f"c2_{name}_{fname}_tflops_best"] = N * (N + 1) * d / best / 1e9
best, med = timeit(lambda: torch.matmul(X, X.T, out=K), 2, 5)
out["c2_cublas_full_gemm_ms"] = best
out["c2_cublas_full_gemm_tflops(2N^2d)"] = 2 * N * N * d / best / 1e9
del K

# config 3 at 1/4 scale: R = 722 500, K = 500
eng = DeviceEngine()
R, Kf = 722500, 500
A = torch.randn((R, Kf), device="cuda", dtype=torch.float64, generator=g)
bb = torch.randn(R, device="cuda", dtype=torch.float64, generator=g)
best, med = timeit(lambda: eng.syrk_partial(A, None, 2500, 100.0, 1.0, bb, None, True), 2, 5)
out["c3q_syrk_ms_best"] = best
out["c3q_syrk_tflops(RK(K+1))"] = R * Kf * (Kf + 1) / best / 1e9
out["c3q_rows_per_s"] = R / best * 1e3
S = torch.empty((Kf, Kf), device="cuda", dtype=torch.float64)
best, med = timeit(lambda: torch.matmul(A.T, A, out=S), 2, 5)
out["c3q_cublas_gemm_ms"] = best
out["c3q_cublas_gemm_tflops(2RK^2)"] = 2 * R * Kf * Kf / best / 1e9
# config 4 slice: n = 500 000, K = 1000 centred
del A, bb
n4, K4 = 500000, 1000
X4 = 10.0 * torch.randn(K4, device="cuda", dtype=torch.float64, generator=g) + torch.randn((n4, K4), device="cuda", dtype=torch.float64, generator=g)
m = eng.colsum(X4) / n4
best, med = timeit(lambda: eng.syrk_partial(X4, None, 0, 1.0, 1.0, None, m, False), 2, 5)
out["c4s_cov_ms_best"] = best
out["c4s_cov_tflops(nK(K+1))"] = n4 * K4 * (K4 + 1) / best / 1e9
best, med = timeit(lambda: eng.colsum(X4), 2, 5)
out["c4s_colsum_ms_best"] = best
out["c4s_colsum_GBs"] = n4 * K4 * 8 / best / 1e6
out["launches"] = plb200.launch_count()
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
