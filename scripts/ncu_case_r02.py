"""ncu driver, round 2: (1) the kDPP kernel build of config 1 (DotProduct, N = 1000, d = 26, Symmetric(:U) storage) on the 64 x 64 tile,
(2) one GPU's packed share of config 2 strong-scaled over 8 GPUs (part 0 of 8) — the 128 x 128 main launch followed by the 64 x 64 tail
launch. PLB200_NO_GRAPH=1 so that every launch is a plain launch."""
import os
import sys

os.environ["PLB200_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import numpy as np
import torch

import plb200
from plb200 import _lib

plb200.init(0)
lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(1)
n, d = 1000, 26
X = torch.as_tensor(rng.standard_normal((n, d)) * 0.3 + 2.0 * rng.standard_normal(d), device="cuda")
K = torch.zeros((n, n), device="cuda", dtype=torch.float64)
for _ in range(3):
    _lib.check(lib.pl_gram_dev(_lib.PL_KERNEL_DOT, X.data_ptr(), n, d, None, 0, 0, d, 2, 0, None, 1.0, 1.0, K.data_ptr(), n, _lib.PL_FILL_JULIA_U, 0, 1, 0, st), "c1")
torch.cuda.synchronize()
N, D = 20000, 300
rng = np.random.default_rng(1002)
X2 = torch.as_tensor(2.0 * rng.standard_normal(D) + rng.standard_normal((N, D)) / np.sqrt(96.0), device="cuda")
nt = int(lib.pl_gram_tile_count(N, 0, 0, 8))
out = torch.empty((nt, 128, 128), device="cuda", dtype=torch.float64)
for _ in range(2):
    _lib.check(lib.pl_gram_dev(_lib.PL_KERNEL_RBF, X2.data_ptr(), N, D, None, 0, 0, D, 0, 0, None, 1.0, 1.0, out.data_ptr(), 128, _lib.PL_FILL_FULL, 0, 8, 1, st), "c2 shard")
torch.cuda.synchronize()
print("ok")
