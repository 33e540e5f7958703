"""The real N-rank path on real GPUs (skipped on a box with fewer than two): torchrun starts one process per GPU, each checks
its share of the distributed apply against the GLOBAL oracle mesh (tests/multirank_worker.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,p,r,mode", [(2, 4, 2, "weak"), (2, 3, 3, "weak"), (2, 4, 3, "strong"), (4, 4, 2, "weak"), (8, 4, 1, "weak"), (8, 4, 3, "strong")])
def test_distributed_apply_matches_global_oracle(world, p, r, mode):
    if n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29600 + (os.getpid() % 300) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_worker.py"), str(p), str(r), mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.count("MULTIRANK_OK") == world, out.stdout[-3000:] + out.stderr[-3000:]
