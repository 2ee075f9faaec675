"""D2H rate of the Symmetric(:U) staircase (row-major lower triangle of a 20000 x 20000 fp64 matrix, 1.61 GB) into pinned host memory,
by shape of the copies: one contiguous 1-D copy of the same byte count, the library's 128-row 2-D blocks, taller blocks, and the
rows of each block split into column panels. cuda-python runtime calls, CUDA-event timing."""
import json
import sys

import numpy as np
import torch
from cuda import cudart

N = 20000
dev = torch.device("cuda", 0)
Kd = torch.zeros((N, N), dtype=torch.float64, device=dev)
Kp = torch.zeros((N, N), dtype=torch.float64).pin_memory()
st = torch.cuda.Stream()
D2H = cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record()
            fn()
            e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def staircase(block_rows, panel_cols=None):
    nbytes = 0
    for r0 in range(0, N, block_rows):
        r1 = min(r0 + block_rows, N)
        w = r1
        c0 = 0
        while c0 < w:
            c1 = w if panel_cols is None else min(c0 + panel_cols, w)
            (err,) = cudart.cudaMemcpy2DAsync(Kp.data_ptr() + (r0 * N + c0) * 8, N * 8, Kd.data_ptr() + (r0 * N + c0) * 8, N * 8, (c1 - c0) * 8, r1 - r0, D2H, st.cuda_stream)
            assert err == cudart.cudaError_t.cudaSuccess, err
            nbytes += (c1 - c0) * 8 * (r1 - r0)
            c0 = c1
    return nbytes


out = {}
tri_bytes = sum(min(r0 + 128, N) * (min(r0 + 128, N) - r0) * 8 for r0 in range(0, N, 128))


def one_d():
    (err,) = cudart.cudaMemcpyAsync(Kp.data_ptr(), Kd.data_ptr(), tri_bytes, D2H, st.cuda_stream)
    assert err == cudart.cudaError_t.cudaSuccess, err


ms = timed(one_d)
out["contiguous_1d"] = {"ms": ms, "GBps": tri_bytes / ms / 1e6}
for br in (128, 256, 512, 1024, 2048):
    nb = {}
    ms = timed(lambda: nb.__setitem__("b", staircase(br)))
    out[f"staircase_{br}_rows"] = {"ms": ms, "GBps": nb["b"] / ms / 1e6, "bytes": nb["b"], "useful_GBps": tri_bytes / ms / 1e6}
for pc in (2048, 8192):
    nb = {}
    ms = timed(lambda: nb.__setitem__("b", staircase(128, pc)))
    out[f"staircase_128_rows_panels_{pc}"] = {"ms": ms, "GBps": nb["b"] / ms / 1e6}
# the same 128-row blocks dealt round-robin to several streams (several copy engines at once)
for ns in (2, 3, 4):
    sts = [torch.cuda.Stream() for _ in range(ns)]

    def multi():
        for s_ in sts:
            s_.wait_stream(st)
        k = 0
        for r0 in range(0, N, 128):
            r1 = min(r0 + 128, N)
            (err,) = cudart.cudaMemcpy2DAsync(Kp.data_ptr() + r0 * N * 8, N * 8, Kd.data_ptr() + r0 * N * 8, N * 8, r1 * 8, r1 - r0, D2H, sts[k % ns].cuda_stream)
            assert err == cudart.cudaError_t.cudaSuccess, err
            k += 1
        for s_ in sts:
            st.wait_stream(s_)

    ms = timed(multi)
    out[f"staircase_128_rows_{ns}_streams"] = {"ms": ms, "GBps": tri_bytes / ms / 1e6}
# full rectangle rows (whole rows, 2-D with full width = 1-D)
ms = timed(lambda: cudart.cudaMemcpyAsync(Kp.data_ptr(), Kd.data_ptr(), N * N * 8, D2H, st.cuda_stream))
out["full_matrix_1d"] = {"ms": ms, "GBps": N * N * 8 / ms / 1e6}
print(json.dumps(out, indent=1))
