import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/potentiallearning.jl_b200")
import numpy as np, torch, plb200, ctypes
from plb200 import _lib
plb200.init(0); lib = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
N, D = 20000, 300
rng = np.random.default_rng(1002)
X = torch.as_tensor(2.0 * rng.standard_normal(D) + rng.standard_normal((N, D)) / np.sqrt(96.0), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
lib.pl_set_profiling(1)
for nparts, N in ((8, 20000), (4, 14100), (2, 10000), (1, 7000), (1, 3550), (8, 40000)):
    rng = np.random.default_rng(1002)
    X = torch.as_tensor(2.0 * rng.standard_normal(D) + rng.standard_normal((N, D)) / np.sqrt(96.0), device="cuda")
    nt = int(lib.pl_gram_tile_count(N, 0, 0, nparts))
    out = torch.empty((nt, 128, 128), device="cuda", dtype=torch.float64)
    for tail in ("0",):
        if tail == "0": os.environ["PLB200_GRAM_NO_TAIL"] = "1"
        else: os.environ.pop("PLB200_GRAM_NO_TAIL", None)
        ts, ks = [], []
        for r in range(12):
            flush.fill_(r & 1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.pl_gram_dev(_lib.PL_KERNEL_RBF, X.data_ptr(), N, D, None, 0, 0, D, 0, 0, None, 1.0, 1.0, out.data_ptr(), 128, _lib.PL_FILL_FULL, 0, nparts, 1, st), "g")
            e1.record(); torch.cuda.synchronize()
            ms = ctypes.c_double(0.0)
            lib.pl_last_main_kernel_ms(ctypes.byref(ms))
            if r >= 3: ts.append(e0.elapsed_time(e1)); ks.append(ms.value)
        iters = -(-nt // 148)
        print("nparts", nparts, "N", N, "slots", nt, "waves", round(nt / 148, 2), "step", round(float(np.median(ts)), 4), "kernel", round(float(np.median(ks)), 4), "us per CTA iteration", round(float(np.median(ks)) * 1e3 / iters, 2), flush=True)
