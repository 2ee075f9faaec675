import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/potentiallearning.jl_b200")
import numpy as np, torch, plb200
from plb200 import _lib
plb200.init(0); lib = _lib.lib()
N, D = 20000, 300
rng = np.random.default_rng(1002)
Xh = 2.0 * rng.standard_normal(D) + rng.standard_normal((N, D)) / np.sqrt(96.0)
Xp = torch.as_tensor(Xh).pin_memory()
Kp = torch.zeros((N, N), dtype=torch.float64).pin_memory()
for i in range(6):
    if i == 4: os.environ["PLB200_STREAM_TRACE"] = "1"
    t = time.perf_counter()
    _lib.check(lib.pl_gram_rbf(_lib.dptr(Xp.numpy()), N, D, D, 0, None, 1.0, 1.0, _lib.dptr(Kp.numpy().reshape(-1)), N, _lib.PL_FILL_JULIA_U), "x")
    print("call", i, (time.perf_counter() - t) * 1e3, _lib.last_timing(), flush=True)
