"""N>1 path on CPU: world_size-2 gloo run of the row-sharded data-parallel step (routing plans, all-to-all row exchange,
dense all-reduce, owner-side segmented update) against a single-process float64 oracle on the global batch."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("model", ["rescal", "sp", "rescal+sp"])
def test_two_rank_gloo_matches_single_process_oracle(model):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), model]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-3000:]


def test_owned_entries_vectorised_equals_per_source_loop():
    """dist._owned_entries (one pass over the concatenated keys of all ranks) against the per-source loop it replaced."""
    import numpy as np
    import torch

    from relation_autoencoder_b200.dist import _excl_cumsum, _owned_entries

    rng = np.random.RandomState(3)
    for world in (1, 2, 3, 8):
        nb, stride = 5, 97
        keys = []
        for s in range(world):
            k = np.unique(rng.randint(0, nb, size=rng.randint(0, 60)) * stride + rng.randint(0, stride, size=1)[0] +
                          rng.randint(0, 7, size=1)[0] * 0)
            k = np.unique(np.concatenate([k, rng.randint(0, nb * stride, size=rng.randint(1, 80))]))
            keys.append(torch.from_numpy(k.astype(np.int64)))
        for rank in range(world):
            eb, erow, esrc, eslot = [], [], [], []
            for s in range(world):
                k = keys[s]
                b = torch.div(k, stride, rounding_mode="floor")
                idd = k - b * stride
                off = _excl_cumsum(torch.bincount(b, minlength=nb))
                slot = torch.arange(k.numel(), dtype=torch.int64) - off[b]
                own = (idd % world) == rank
                eb.append(b[own]); erow.append(torch.div(idd[own], world, rounding_mode="floor")); eslot.append(slot[own])
                esrc.append(torch.full((int(own.sum().item()),), s, dtype=torch.int64))
            want = [torch.cat(x) for x in (eb, erow, esrc, eslot)]
            got = _owned_entries(keys, nb, stride, rank, world)
            for w, g in zip(want, got):
                assert torch.equal(w, g)
