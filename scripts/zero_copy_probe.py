"""SM-driven D2H of the Symmetric(:U) staircase into pinned (mapped) host memory: GB/s by number of CTAs. scripts/native/zero_copy_probe.cu."""
import ctypes
import json
import os

import numpy as np
import torch

import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "native", "libzcprobe.so")
if not os.path.exists(SO):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC", "-o", SO,
                           os.path.join(HERE, "native", "zero_copy_probe.cu")])
L = ctypes.CDLL(SO)
L.zc_stair_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
N = 20000
dev = torch.device("cuda", 0)
Kd = torch.randn((N, N), dtype=torch.float64, device=dev)
Kp = torch.zeros((N, N), dtype=torch.float64).pin_memory()
tri_bytes = N * (N + 1) // 2 * 8
st = torch.cuda.current_stream().cuda_stream
out = {}
for nb, th in ((1, 256), (2, 256), (4, 256), (8, 256), (8, 1024), (16, 512), (32, 256), (148, 256)):
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = L.zc_stair_copy(Kd.data_ptr(), Kp.data_ptr(), N, N, 0, N, nb, th, st)
        assert rc == 0, rc
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    out[f"{nb}x{th}"] = {"ms": min(ts), "GBps": tri_bytes / min(ts) / 1e6}
    print(nb, th, out[f"{nb}x{th}"], flush=True)
ok = bool(torch.equal(torch.tril(Kp), torch.tril(Kd.cpu())))
out["correct"] = ok
print(json.dumps(out))
