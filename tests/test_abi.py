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
    multi = open(os.path.join(pkg, "csrc", "multi.inl")).read()
    body = api[api.index('extern "C" {'):] + multi[multi.index('extern "C" {'):]
    # every exported compute entry point (everything but lifetime / bookkeeping helpers) checks for an initialised sm_100 device
    for m in re.finditer(r"\n(?:int|void\*)\s+(pl_\w+)\s*\([^)]*\)\s*\{(.*?)\n\}\n", body, flags=re.S):
        name, code = m.group(1), m.group(2)
        if name in ("pl_init", "pl_init_all", "pl_device_count", "pl_last_timing", "pl_last_host_ms", "pl_set_profiling", "pl_host_free", "pl_gram_tile_coords", "pl_syrk_schedule",
                    "pl_comm_unique_id", "pl_comm_info", "pl_comm_destroy", "pl_comm_peer_enabled", "pl_gram_resident_info", "pl_gram_resident_shard", "pl_gram_free"):
            continue  # lifetime / bookkeeping: no compute
        assert ("need_ready()" in code) or re.search(r"return \w+_impl\(|return gram_host\(|return rmsd_host\(|return normal_eq_stream\(", code), name


def _syrk_schedule(R, K, sms, mode, waves=0, diag=0.0):
    import ctypes

    lib = _lib.load()
    npc, ncta = ctypes.c_int64(0), ctypes.c_int64(0)
    assert lib.pl_syrk_schedule(R, K, sms, mode, waves, diag, 0, None, ctypes.byref(npc), ctypes.byref(ncta)) == 0
    tab = np.zeros((max(npc.value, 1), 5), dtype=np.int32)
    assert lib.pl_syrk_schedule(R, K, sms, mode, waves, diag, npc.value, tab.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), ctypes.byref(npc),
                                ctypes.byref(ncta)) == 0
    return tab[: npc.value], int(ncta.value)


def test_syrk_schedules_cover_every_stage_once():
    """Both SYRK schedules (host tables the kernel walks) give every pipeline stage of every work unit to exactly one CTA; the partial
    slots of a unit are contiguous and in row order (the second-stage sum is a fixed-order walk over them); CTA piece lists are
    contiguous; the stream-K cut is balanced in cost, aligned across units of one kind and issued in row order."""
    for R, K in [(0, 500), (1, 8), (15, 128), (16, 129), (1000, 500), (5000, 257), (361_250, 500), (2_890_000, 500), (1_250_000, 1000), (100_000, 1500)]:
        T = (K + 127) // 128
        nunits = T * (T - 1) // 2 + (T + 1) // 2
        noff = T * (T - 1) // 2
        nst = (R + 15) // 16
        for mode, waves in [(0, 0), (0, 7), (1, 0), (1, 1), (1, 2), (1, 4), (1, 8)]:
            tab, nctas = _syrk_schedule(R, K, 148, mode, waves, 1.125)
            if nst == 0:
                assert len(tab) == 0
                continue
            assert np.all(tab[:, 3] > 0) and tab[:, 0].max() < nctas
            assert np.all(np.diff(tab[:, 0]) >= 0), "a CTA's pieces are adjacent rows of the table"
            slots = np.sort(tab[:, 4])
            assert np.array_equal(slots, np.arange(len(tab))), "slots are a permutation of 0..npieces-1"
            by_slot = tab[np.argsort(tab[:, 4])]
            assert np.all(np.diff(by_slot[:, 1]) >= 0), "slots are unit-major"
            for u in range(nunits):
                rows = by_slot[by_slot[:, 1] == u]
                assert len(rows) >= 1 and rows[0, 2] == 0
                assert np.array_equal(rows[1:, 2], (rows[:, 2] + rows[:, 3])[:-1]), "row order, no gap, no overlap"
                assert rows[-1, 2] + rows[-1, 3] == nst
            if mode == 1 and nst >= 4096:
                assert nctas <= 148 * (waves or 20) or nctas == nunits or (noff == 0 and nctas == 1), (nctas, waves)
                assert len(tab) == nctas, "one piece per CTA"
                cuts = [tuple(map(tuple, by_slot[by_slot[:, 1] == u][:, 2:4])) for u in range(nunits)]
                assert len(set(cuts[:noff])) <= 1 and len(set(cuts[noff:])) == 1, "units of one kind are cut at identical rows"
                assert np.all(np.diff(tab[:, 2]) >= 0), "CTAs are issued in row order"
                cost = np.where(tab[:, 1] < noff, 1.0, 1.125) * tab[:, 3]
                n0 = len(cuts[0]) if noff else len(cuts[-1])
                assert cost.max() <= cost.min() * (1.0 + 1.3 / n0) + 2, (R, K, waves, cost.min(), cost.max())
                rounds = -(-nctas // 148)
                eff = (noff + 1.125 * (nunits - noff)) * nst / (148 * rounds * cost.max())
                if waves in (0, 8):  # (small caps are experiments: one round of 78 CTAs cannot fill 148 SMs)
                    assert eff > (0.97 if waves == 0 else 0.90), (R, K, waves, nctas, eff)
                if noff:
                    assert cost[tab[:, 1] >= noff].max() <= cost[tab[:, 1] < noff].max() + 1.125, "diagonal CTAs are never the longer ones"
