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


def test_sharded_step_with_nccl_dense_allreduce_aside():
    """The dense gradient too large for the peer-read update (forced here: threshold 0) goes through NCCL's all-reduce on
    its own stream, beside the flag barrier and the sparse-row applies; the dense update waits for its event."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), "rescal+sp", "cuda"]
    env = dict(os.environ, RAE_TEST_PEER_DENSE_MAX="0")
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-3000:]


def test_peer_barrier_timeout_is_fatal_and_skips_the_pulls():
    """A rank that never arrives (here: world = 2 with nobody behind rank 1's flag word) makes the bounded barrier time
    out (~2 s).  From then on the handle is dead: the owner-side pull kernels leave the tables untouched (the peers'
    gradient rows may be stale), rae_peer_status and every later rae_dist_* call report the failure."""
    import ctypes as C

    import torch

    from relation_autoencoder_b200 import _lib as L
    from relation_autoencoder_b200.engine import Engine

    def pull(eng, table, acc, grad):
        rows = torch.tensor([0, 1], dtype=torch.int32, device="cuda")
        off = torch.tensor([0, 1, 2], dtype=torch.int32, device="cuda")
        src = torch.zeros(2, dtype=torch.int32, device="cuda")
        slot = torch.tensor([0, 1], dtype=torch.int32, device="cuda")
        grads = (C.c_void_p * 1)(grad.data_ptr())
        rc = eng.lib.rae_pull_apply(eng._h, table.data_ptr(), acc.data_ptr(), 8, rows.data_ptr(), off.data_ptr(), src.data_ptr(),
                                    slot.data_ptr(), 2, grads, 1, eng._stream)
        torch.cuda.synchronize()
        return rc

    eng = Engine("rescal+sp", 5, 6, 2, 8, 20, 10, 8)
    table = torch.ones(4, 8, device="cuda")
    acc = torch.zeros(4, 8, device="cuda")
    grad = torch.ones(2, 8, device="cuda")
    assert pull(eng, table, acc, grad) == 0
    assert not torch.equal(table[:2], torch.ones(2, 8, device="cuda"))         # healthy handle: the rows were updated
    flags = []
    for _ in range(2):
        p = C.c_void_p(0)
        assert eng.lib.rae_peer_alloc(4096, C.byref(p), None) == 0
        flags.append(p)
    arr = (C.c_void_p * 2)(flags[0].value, flags[1].value)
    assert eng.lib.rae_peer_barrier(eng._h, arr, 2, 0, eng._stream) == 0         # launches; rank 1 never publishes
    torch.cuda.synchronize()
    assert eng.lib.rae_peer_status(eng._h, eng._stream) != 0
    assert b"timed out" in eng.lib.rae_last_error(eng._h)
    before = table.clone()
    pull(eng, table, acc, grad)
    assert torch.equal(table, before)                                            # skipped: bit-identical
    step = L.RaeDistStep()
    assert eng.lib.rae_dist_step_begin(eng._h, C.byref(step), eng._stream) != 0   # every later call fails on the host side
    for p in flags:
        eng.lib.rae_peer_free(p)
    eng.close()
