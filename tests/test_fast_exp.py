"""Accuracy of the branch-free exp used by the fused RBF epilogue (csrc/gram.cuh exp_fast / exp_clamped), restated in C
with the same operation order (tests/fast_exp_model.c) and compared with long double expl. No GPU needed."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _lib():
    src = os.path.join(HERE, "fast_exp_model.c")
    out = os.path.join(HERE, "..", "oracle", "_build", "libfastexp_model.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-mfma", "-fPIC", "-shared", "-o", out, src, "-lm"])
    L = ctypes.CDLL(out)
    L.exp_clamped_model.restype = ctypes.c_double
    L.exp_clamped_model.argtypes = [ctypes.c_double]
    L.exp_model_max_rel_err.restype = ctypes.c_double
    L.exp_model_max_rel_err.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int64]
    return L


def test_exp_fast_is_within_two_ulp_over_the_whole_range():
    L = _lib()
    assert L.exp_model_max_rel_err(-707.99, 0.0, 4_000_000) < 2.2e-16  # ~1.2 ulp measured
    assert L.exp_model_max_rel_err(-30.0, 0.0, 4_000_000) < 2.2e-16   # the range RBF kernels live in
    assert L.exp_model_max_rel_err(0.0, 709.0, 1_000_000) < 2.2e-16
    assert L.exp_clamped_model(0.0) == 1.0


def test_exp_clamped_edges():
    L = _lib()
    assert 0.0 < L.exp_clamped_model(-707.999) < 3.4e-308
    assert L.exp_clamped_model(-708.0) == 0.0  # at and below -708 the reference result is < 3.3e-308: flushed to 0
    assert L.exp_clamped_model(-708.0000001) == 0.0
    assert L.exp_clamped_model(-1e300) == 0.0 and L.exp_clamped_model(-np.inf) == 0.0
    xs = np.linspace(-700, 0, 2001)
    ys = np.array([L.exp_clamped_model(float(x)) for x in xs])
    assert np.all(np.diff(ys) > 0), "monotone"
