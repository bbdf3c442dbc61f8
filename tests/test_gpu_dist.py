"""Multi-GPU path on real devices: the row-sharded data-parallel step with the CUDA backend (NCCL) against the float64
oracle on the global batch.  world_size 1 always runs (exercises fetch / emit-only / owner-side apply on one GPU);
world_size 2 and 4 run when the box has that many GPUs."""
import os
import subprocess
import sys

import pytest

from tests.test_dist_cpu import ROOT, _free_port

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("model", ["rescal", "sp", "rescal+sp"])
def test_sharded_step_matches_oracle(model, world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), model, "cuda"]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-3000:]
