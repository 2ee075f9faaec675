#!/usr/bin/env python
"""Builds scripts/native/gridbar_probe.cu on demand (nvcc, sm_100a) and runs it: cycles per publish + grid barrier + consume for seven
barrier designs and dependent-chain latencies of the scalar FP64 pipe (profiles/r02e_gridbar_probe.json). Usage: gridbar_probe.py [out.json]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "native", "gridbar_probe")
SRC = os.path.join(HERE, "native", "gridbar_probe.cu")
if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(SRC):
    subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", EXE, SRC])
out = subprocess.run([EXE], capture_output=True, text=True, check=True).stdout
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(out)
print(out)
