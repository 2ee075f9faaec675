"""Accuracy of the table-driven 2^(y/256) that IS the fused RBF epilogue (csrc/gram.cuh exp2_tab), restated in C with the same
operation order (tests/fast_exp_model.c) and compared with long double exp2l. No GPU needed."""
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
    L.exp2_tab_model.restype = ctypes.c_double
    L.exp2_tab_model.argtypes = [ctypes.c_double]
    L.exp2_scale.restype = ctypes.c_double
    L.exp2_model_max_rel_err.restype = ctypes.c_double
    L.exp2_model_max_rel_err.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int64]
    return L


def test_exp2_tab_is_within_one_ulp_over_the_whole_range():
    L = _lib()
    s = L.exp2_scale()
    assert L.exp2_model_max_rel_err(-707.99 * s, 0.0, 4_000_000) < 2.3e-16  # 1 ulp: table entry (0.5) + final rounding (0.5)
    assert L.exp2_model_max_rel_err(-30.0 * s, 0.0, 4_000_000) < 2.3e-16  # the range RBF kernels live in
    assert L.exp2_model_max_rel_err(0.0, 709.0 * s, 1_000_000) < 2.3e-16  # beta > 1
    assert L.exp2_tab_model(0.0) == 1.0
    assert L.exp2_tab_model(256.0) == 2.0 and L.exp2_tab_model(-256.0 * 10) == 2.0**-10


def test_exp2_tab_edges():
    L = _lib()
    s = L.exp2_scale()
    assert 0.0 < L.exp2_tab_model(-707.99 * s) < 3.4e-308
    # below x = -708.4 (results under 2.3e-308: denormal or zero in the reference) N is clamped: some number <= 2^-1021
    for x in (-708.5, -709.0, -745.0, -1e4, -5e6):
        assert 0.0 < L.exp2_tab_model(x * s) <= 2.0**-1021
    xs = np.linspace(-700, 0, 2001)
    ys = np.array([L.exp2_tab_model(float(x) * s) for x in xs])
    assert np.all(np.diff(ys) > 0), "monotone"
    # fine-grained monotonicity across table-entry boundaries
    ys = np.array([L.exp2_tab_model(-300.0 + 0.001 * i) for i in range(4000)])
    assert np.all(np.diff(ys) >= 0)
