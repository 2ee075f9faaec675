#!/usr/bin/env python
"""Shallow contractions (d <= 64) at large N: 128 x 128 tiles (TMA-store epilogue for DotProduct) against 64 x 64 tiles, Symmetric(:U) storage and
both triangles. GB/s of result written against the measured copy peak. Decides the tile rule of launch_gram; writes gpurun_out/r02c_tile_rule_probe.json."""
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
hbm = 6545.0
res = []
g = torch.Generator(device=dev).manual_seed(3)
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)
for n in (2000, 4000, 6000, 8000, 10000, 14000):
    K = torch.zeros((n, n), device=dev, dtype=torch.float64)
    for d in (26, 300):
        X = torch.randn((n, d), device=dev, dtype=torch.float64, generator=g) * 0.3 + 2.0
        for kname, kid in (("dot", _lib.PL_KERNEL_DOT), ("rbf", _lib.PL_KERNEL_RBF)):
            for fill, fname in ((_lib.PL_FILL_JULIA_U, "U"),):
                row = {"n": n, "d": d, "kernel": kname, "fill": fname}
                for small in ("0", "1"):
                    os.environ["PLB200_GRAM_SMALL"] = small
                    ts = []
                    for it in range(5):
                        flush.fill_(1.0)
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        _lib.check(lib.pl_gram_dev(kid, X.data_ptr(), n, d, None, 0, 0, d, 2, 0, None, 1.0, 1.0, K.data_ptr(), n, fill, 0, 1, 0,
                                                   torch.cuda.current_stream().cuda_stream), "pl_gram_dev")
                        e1.record()
                        e1.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    ms = min(ts[1:])
                    wbytes = (n * (n + 1) / 2 if fname == "U" else n * n) * 8
                    row["ms_64" if small == "1" else "ms_128"] = round(ms, 4)
                    row["TBps_64" if small == "1" else "TBps_128"] = round(wbytes / (ms * 1e-3) / 1e12, 3)
                row["frac_hbm_best"] = round(max(row["TBps_64"], row["TBps_128"]) * 1e3 / hbm, 3)
                res.append(row)
                print(row, flush=True)
    del K
os.environ.pop("PLB200_GRAM_SMALL", None)
json.dump({"hbm_copy_peak_GBps": hbm, "rows": res}, open(os.path.join(ROOT, "gpurun_out", "r02c_tile_rule_probe.json"), "w"), indent=1)
