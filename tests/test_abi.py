"""The C-ABI shared library loads and exports every symbol include/plb200.h declares; without a GPU it fails loudly
(no CPU fallback, no silent success). No compute is attempted here."""
import ctypes
import os
import re

import numpy as np
import pytest

import plb200
from plb200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "plb200.h")


def header_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pl_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_boundary():
    names = header_functions()
    for required in ("pl_init", "pl_finalize", "pl_last_error", "pl_gram_dot", "pl_gram_rbf", "pl_gram_cross", "pl_normal_eq", "pl_cov",
                     "pl_gram_dev", "pl_syrk_dev", "pl_colsum_dev"):
        assert required in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libplb200.so must be built in-tree (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} is declared in include/plb200.h but not exported by libplb200.so"


def test_python_prototypes_match_header():
    assert sorted(_lib.PROTOTYPES) == header_functions()
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.PROTOTYPES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", txt, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), f"{name}: header has {len(params)} parameters, binding has {len(args)}"


def test_constants_match_header():
    txt = open(HEADER).read()
    for cname in ("PL_KERNEL_LINEAR", "PL_KERNEL_DOT", "PL_KERNEL_RBF", "PL_KERNEL_IMQ", "PL_CINV_IDENTITY", "PL_CINV_DIAGONAL", "PL_CINV_FULL",
                  "PL_FILL_RM_UPPER", "PL_FILL_RM_LOWER", "PL_FILL_FULL"):
        m = re.search(r"#define\s+" + cname + r"\s+(\d+)", txt)
        assert m and int(m.group(1)) == getattr(_lib, cname)


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    lib = _lib.load()
    assert lib.pl_init(-1) != 0
    assert b"CPU fallback" in lib.pl_last_error() or b"device" in lib.pl_last_error()
    X = np.ones((4, 2))
    K = np.zeros((4, 4))
    rc = lib.pl_gram_dot(_lib.dptr(X), 4, 2, 2, 2, _lib.dptr(K), 4, _lib.PL_FILL_FULL)
    assert rc != 0 and np.all(K == 0.0), "no result may be produced without the GPU"
    with pytest.raises(plb200.PLB200Error):
        plb200.KernelMatrix(X, plb200.DotProduct())
    with pytest.raises(plb200.PLB200Error):
        plb200.normal_equations(X, np.ones(4))
    with pytest.raises(plb200.PLB200Error):
        plb200.covariance(X)


def test_tile_enumeration_helpers_run_on_host():
    lib = _lib.load()
    for n, nparts in ((1, 1), (700, 3), (897, 3), (20000, 8)):
        T = (n + 127) // 128
        seen = set()
        for part in range(nparts):
            for ti, tj in plb200.gram_tile_coords(n, 0, part, nparts):
                if ti >= 0:
                    assert 0 <= ti <= tj < T and (ti, tj) not in seen
                    seen.add((ti, tj))
        assert len(seen) == T * (T + 1) // 2
        counts = [int(lib.pl_gram_tile_count(n, 0, p, nparts)) for p in range(nparts)]
        assert max(counts) - min(counts) <= 1, "tiles are dealt evenly"


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under potentiallearning.jl_b200/ (the product) may import, load or execute it, and the
    library has no CPU code path to fall back to (every compute entry point starts with need_ready())."""
    pkg = os.path.join(ROOT, "potentiallearning.jl_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl")):
                txt = open(os.path.join(dirpath, f), encoding="utf-8").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|libploracle|pl_oracle", txt, flags=re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
    api = open(os.path.join(pkg, "csrc", "plb200_api.cu")).read()
    body = api[api.index('extern "C" {'):]
    # every exported compute entry point (everything but lifetime / bookkeeping helpers) checks for an initialised sm_100 device
    for m in re.finditer(r"\n(?:int|void\*)\s+(pl_\w+)\s*\([^)]*\)\s*\{(.*?)\n\}\n", body, flags=re.S):
        name, code = m.group(1), m.group(2)
        if name in ("pl_init", "pl_last_timing", "pl_set_profiling", "pl_host_free", "pl_gram_tile_coords"):
            continue
        assert ("need_ready()" in code) or re.search(r"return \w+_impl\(|return gram_host\(|return rmsd_host\(|return normal_eq_stream\(", code), name
