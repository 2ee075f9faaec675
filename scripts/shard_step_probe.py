#!/usr/bin/env python
"""One GPU's share of config 2 strong-scaled over 8 GPUs (part 0 of 8: 1551 packed tiles = 10.48 waves), timed on ONE GPU exactly as bench.py
times a step: the effect of the tail-wave sub-tiles, the cached CUDA graph and PDL on the step. Writes gpurun_out/r02c_shard_step_probe.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "potentiallearning.jl_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import plb200
from plb200 import _lib

plb200.init(0)
lib = _lib.lib()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(1002)
N, D = 20000, 300
X = torch.as_tensor(2.0 * rng.standard_normal(D) + rng.standard_normal((N, D)) / np.sqrt(96.0), device=dev)
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)
res = []
for nparts in (8, 4, 2):
    nt = int(lib.pl_gram_tile_count(N, 0, 0, nparts))
    out = torch.empty((nt, 128, 128), device=dev, dtype=torch.float64)
    for tail in (False, True):
        for graph in (False, True):
            os.environ.pop("PLB200_GRAM_NO_TAIL", None)
            os.environ.pop("PLB200_NO_GRAPH", None)
            if not tail:
                os.environ["PLB200_GRAM_NO_TAIL"] = "1"
            if not graph:
                os.environ["PLB200_NO_GRAPH"] = "1"

            def step():
                _lib.check(lib.pl_gram_dev(_lib.PL_KERNEL_RBF, X.data_ptr(), N, D, None, 0, 0, D, 0, 0, None, 1.0, 1.0, out.data_ptr(), 128, _lib.PL_FILL_FULL, 0, nparts, 1,
                                           torch.cuda.current_stream().cuda_stream), "pl_gram_dev")

            for _ in range(5):
                step()
            torch.cuda.synchronize()
            ts = []
            for _ in range(20):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            row = {"nparts": nparts, "tiles": nt, "waves": round(nt / 148, 2), "tail_subtiles": tail, "cuda_graph": graph, "ms_mean": round(float(np.mean(ts)), 4),
                   "ms_min": round(min(ts), 4), "strong_scaling_eff_vs_3.652ms": round(3.652 / (nparts * float(np.mean(ts))), 4)}
            res.append(row)
            print(row, flush=True)
os.environ.pop("PLB200_GRAM_NO_TAIL", None)
os.environ.pop("PLB200_NO_GRAPH", None)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r02c_shard_step_probe.json"), "w"), indent=1)
