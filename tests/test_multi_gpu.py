"""Several GPUs behind the C ABI (csrc/multi.inl): ONE process driving all devices (pl_init_all — what a Julia session does) and one
process per GPU (pl_comm_init_rank + the *_sharded entry points), NCCL issued by the library itself. The checks live in
scripts/multi_gpu_selftest.py (numpy fp64 / oracle references, 1e-11) and run in fresh processes because pl_init_all needs an
uninitialised library. With one GPU only the one-device legs run; `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu` runs all."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "scripts", "multi_gpu_selftest.py")


def ngpus():
    import torch

    return torch.cuda.device_count()


def run(cmd, timeout=900, extra_env=None):
    env = dict(os.environ)
    env.setdefault("MASTER_ADDR", "127.0.0.1")
    env.update(extra_env or {})
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert p.returncode == 0, f"{' '.join(cmd)}\nSTDOUT:\n{p.stdout[-3000:]}\nSTDERR:\n{p.stderr[-3000:]}"
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, p.stdout[-2000:]
    return json.loads(lines[-1])


def test_init_all_with_one_device_is_the_plain_path():
    out = run([sys.executable, SCRIPT, "single", "1"])
    assert out["ok"] and out["devices"] == 1


# peer = "1": the K x K exchange over peer memory (NVLink loads / stores fused behind the SYRK's reduce kernel); "0": NCCL all-reduce
@pytest.mark.skipif("ngpus() < 2")
@pytest.mark.parametrize("peer", ["1", "0"])
def test_single_process_drives_all_gpus(peer):
    n = min(ngpus(), 8)
    out = run([sys.executable, SCRIPT, "single", str(n)], extra_env={"PLB200_PEER_ALLREDUCE": peer})
    assert out["ok"] and out["devices"] == n and out["nccl"]["nranks"] == n
    if peer == "0":
        assert not out["nccl"]["peer_exchange"]


@pytest.mark.skipif("ngpus() < 2")
@pytest.mark.parametrize("peer", ["1", "0"])
def test_one_process_per_gpu_sharded_entry_points(peer):
    n = min(ngpus(), 8)
    out = run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port",
               "29613" if peer == "1" else "29614", SCRIPT, "ranks"], extra_env={"PLB200_PEER_ALLREDUCE": peer})
    assert out["ok"] and out["world"] == n
    if peer == "0":
        assert not out["nccl"]["peer_exchange"]
