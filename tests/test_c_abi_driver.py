"""tests/c_abi_driver.c: the boundary exercised without Python marshalling — plain C, dlopen, the exact argument lists of the Julia
shim's ccalls, buffers in Julia's column-major layout, compared with the oracle in Julia's own indexing."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi_driver.c")
EXE = os.path.join(ROOT, "tests", "_build", "c_abi_driver")


def build_driver():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O1", "-Wall", "-o", EXE, SRC, "-ldl", "-lm"])
    return EXE


def libs():
    sys.path.insert(0, ROOT)
    from oracle import oracle
    from plb200 import _lib

    return _lib.LIB_PATH, oracle.build_c()


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_driver_builds_and_fails_loudly_without_a_gpu():
    exe = build_driver()
    plb, orc = libs()
    p = subprocess.run([exe, plb, orc], capture_output=True, text=True, timeout=120)
    assert p.returncode == 1 and "pl_init_all" in p.stderr and "OK" not in p.stdout, (p.stdout, p.stderr)


@pytest.mark.gpu
def test_c_abi_driver_matches_the_oracle():
    exe = build_driver()
    plb, orc = libs()
    p = subprocess.run([exe, plb, orc, "1"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.startswith("OK"), (p.stdout, p.stderr)
    assert int(p.stdout.split()[1]) >= 9
