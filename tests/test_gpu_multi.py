"""Multi-GPU paths, run as one process per GPU under torchrun (skipped on a single-GPU box): the single-lattice CMC / SA
driver with its in-kernel peer-memory exchange must give the world-size-independent trajectory."""
import os
import subprocess
import sys

import pytest

from latticemontecarlo_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_multi_gpu_single_lattice_cmc_is_world_size_independent():
    n = capi.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "cmc_multi_gpu.py"), "28", "60000"]
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=540)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if "identical_to_world1" in l]
    assert len(lines) == 2 and all("identical_to_world1=True" in l for l in lines), res.stdout[-2000:]


@pytest.mark.timeout(600)
def test_multi_gpu_domain_decomposed_cmc_equals_the_single_gpu_run():
    """lmc_cmc_domain_run over several GPUs (x slabs of the domain grid, in-kernel halo / migration exchange through peer
    memory): all ranks end identical and equal to the world-1 run of the same seed (tools/domain_multi_gpu.py)."""
    n = capi.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else 4
    for args in (["24", "1500000"], ["32", "3000000", "sa"]):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", "29543", os.path.join(ROOT, "tools", "domain_multi_gpu.py")] + args
        res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=280)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        lines = [l for l in res.stdout.splitlines() if "equals world-1 run" in l]
        assert len(lines) == 1 and "ranks identical True, equals world-1 run True" in lines[0], res.stdout[-2000:]
