#!/usr/bin/env python
"""Small / skinny kernel matrices (config 1: kDPP kernel build, N = 1000 x d = 26): GPU time per matrix with the launch overhead of the host
taken out (CUDA-graph replay of 20 calls) next to back-to-back calls through the C ABI, for the 128 x 128 and 64 x 64 tile shapes and with /
without programmatic dependent launch of the Gram kernel behind its prelude. Writes one JSON document (profiles/r02_small_gram_probe.json)."""
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
out = {"hbm_copy_peak_GBps": hbm, "cases": []}


def run(kid, X, K, fill):
    n, d = X.shape
    _lib.check(lib.pl_gram_dev(kid, X.data_ptr(), n, X.stride(0), None, 0, 0, d, 2, 0, None, 1.0, 1.0, K.data_ptr(), n, fill, 0, 1, 0,
                               torch.cuda.current_stream().cuda_stream), "pl_gram_dev")


def measure(fn):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        fn()
    e1.record()
    e1.synchronize()
    b2b = e0.elapsed_time(e1) / 200 * 1e3
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                fn()
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    e1.synchronize()
    return b2b, e0.elapsed_time(e1) / 400 * 1e3


rng = np.random.default_rng(1)
for n, d in ((1000, 26), (2000, 26), (300, 26), (1000, 300), (4000, 26)):
    X = torch.as_tensor(rng.standard_normal((n, d)) * 0.3 + 2.0 * rng.standard_normal(d), device=dev)
    K = torch.zeros((n, n), device=dev, dtype=torch.float64)
    ref = {}
    for kname, kid in (("dot", _lib.PL_KERNEL_DOT), ("rbf", _lib.PL_KERNEL_RBF)):
        for small in ("0", "1"):
            for nopdl in ("1", "0"):
                os.environ["PLB200_GRAM_SMALL"] = small
                if nopdl == "1":
                    os.environ["PLB200_NO_PDL"] = "1"
                else:
                    os.environ.pop("PLB200_NO_PDL", None)
                K.zero_()
                b2b, graph = measure(lambda: run(kid, X, K, _lib.PL_FILL_JULIA_U))
                torch.cuda.synchronize()
                key = kname
                if key not in ref:
                    ref[key] = K.clone()
                same = bool(torch.equal(ref[key], K))
                bytes_ = n * (n + 1) / 2 * 8 + n * d * 8
                out["cases"].append({"n": n, "d": d, "kernel": kname, "tile": "64x64" if small == "1" else "128x128", "pdl": nopdl == "0",
                                     "us_back_to_back": round(b2b, 2), "us_graph_replay": round(graph, 2),
                                     "GBps_graph": round(bytes_ / (graph * 1e-6) / 1e9, 1), "frac_hbm": round(bytes_ / (graph * 1e-6) / 1e9 / hbm, 4),
                                     "bitwise_equal_to_first_variant": same})
os.environ.pop("PLB200_GRAM_SMALL", None)
os.environ.pop("PLB200_NO_PDL", None)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_small_gram_probe.json"), "w"), indent=1)
for c in out["cases"]:
    print(c)
