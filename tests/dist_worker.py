"""Worker for tests/test_dist_cpu.py: run under torch.distributed.run with the gloo backend (world_size 2).
Each rank owns B examples of every global batch; after a few steps the gathered parameters must equal a single-process
float64 oracle run on the global batches."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import rae_oracle as O  # noqa: E402
from relation_autoencoder_b200.dist import DistributedEngine  # noqa: E402
from tests.dist_numpy_backend import NumpyBackend  # noqa: E402
from tests.helpers import make_problem  # noqa: E402


def main():
    model = sys.argv[1]
    cuda = len(sys.argv) > 2 and sys.argv[2] == "cuda"
    if cuda:
        lr_ = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(lr_)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
    else:
        dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    B, K, d, S, F, N, nb = (64, 100, 32, 5, 3000, 900, 3) if cuda else (6, 5, 4, 3, 40, 23, 3)
    Bg = B * world
    pr = make_problem(model, B=Bg * nb, K=K, d=d, S=S, F=F, N=N, seed=5, dup_heavy=True)
    p0 = {k: v.astype(np.float32).astype(np.float64) for k, v in pr["p"].items()}
    # local rows of global batch g: [g*Bg + rank*B, g*Bg + (rank+1)*B)
    rows = np.concatenate([np.arange(g * Bg + rank * B, g * Bg + (rank + 1) * B) for g in range(nb)])
    ip = pr["indptr"]
    loc_ip = [0]
    loc_ix = []
    for r in rows:
        loc_ix.extend(pr["indices"][ip[r]:ip[r + 1]].tolist())
        loc_ip.append(len(loc_ix))
    de = DistributedEngine(model, K, d, S, B, F, N, n_train=Bg * nb, lr=0.1, alpha=0.7, rank=rank, world=world,
                           device=int(os.environ.get("LOCAL_RANK", "0")),
                           backend_factory=None if cuda else NumpyBackend,
                           peer_dense_max_bytes=int(os.environ.get("RAE_TEST_PEER_DENSE_MAX", str(16 << 20))))
    tol = 1e-5 if cuda else 1e-12
    de.set_params_numpy(p0)
    de.bind_split("train", np.asarray(loc_ip), np.asarray(loc_ix), pr["a1"][rows], pr["a2"][rows])
    de.bind_epoch_negatives(pr["neg1"][:, rows], pr["neg2"][:, rows])
    om = O.OracleModel(model, {k: v.copy() for k, v in p0.items()}, K=K, d=d, S=S, B=Bg, lr=0.1, alpha=0.7)
    om.bind_split("train", pr["indptr"], pr["indices"], pr["a1"], pr["a2"])
    ok = True
    for g in range(nb):
        if cuda:     # fp32 path: every step is checked from the GPUs' current state (see tests/test_gpu_parity.py)
            om.params = {k: v.astype(np.float64) for k, v in de.get_params_numpy().items()}
            om.acc = {k: v.astype(np.float64) for k, v in de.get_acc_numpy().items()}
        if g == 1:   # the reference-style call with host negatives
            c = de.train(g, pr["neg1"][:, rows[g * B:(g + 1) * B]], pr["neg2"][:, rows[g * B:(g + 1) * B]])
        else:
            c = de.train_device(g)
        c_ref = om.train(g, pr["neg1"][:, g * Bg:(g + 1) * Bg], pr["neg2"][:, g * Bg:(g + 1) * Bg])
        ok &= abs(c - c_ref) < tol * max(1.0, abs(c_ref))
        if cuda:
            got = de.get_params_numpy()
            acc = de.get_acc_numpy()
            for n in O.param_names(model):
                # accumulators are sums of squared gradients: a direct check of the summed row gradients
                ok &= np.abs(acc[n] - om.acc[n]).max() <= 1e-4 * max(np.abs(om.acc[n]).max(), 1e-30)
                # parameters: AdaGrad's first steps are sign-like for tiny gradients, so allow a few lr-sized outliers
                diff = np.abs(got[n] - om.params[n])
                ok &= np.mean(diff <= 1e-5 * max(1.0, np.abs(om.params[n]).max())) > 0.999 and diff.max() < 5e-3
            if not ok and rank == 0:
                print("mismatch at step", g, c, c_ref)
    got = de.get_params_numpy()
    acc = de.get_acc_numpy()
    if not cuda:
        for n in O.param_names(model):
            ok &= np.abs(got[n] - om.params[n]).max() < 1e-12
            ok &= np.abs(acc[n] - om.acc[n]).max() < 1e-12
    else:
        om.params = {k: v.astype(np.float64) for k, v in got.items()}
    # labelling through the sharded W
    lab, probs = de.label("train", 1)
    lab_ref, q_ref = om.label("train", 1)
    ok &= np.array_equal(lab, lab_ref[rank * B:(rank + 1) * B])
    ok &= np.abs(probs - q_ref[rank * B:(rank + 1) * B]).max() < max(tol, 1e-12) * q_ref.max()
    if cuda:
        # negatives that are not the bound epoch's columns are refused (the routing plan was built from the bound ones)
        bad = np.array(pr["neg1"][:, rows[:B]], dtype=np.int32)
        bad[0, 0] = (bad[0, 0] + 1) % N
        try:
            de.train(0, bad, pr["neg2"][:, rows[:B]])
            ok = False
        except RuntimeError:
            pass
        c = de.train(0, pr["neg1"][:, rows[:B]], pr["neg2"][:, rows[:B]])       # the flag is cleared by the report
        ok &= bool(np.isfinite(c))
    flag = torch.tensor([1 if ok else 0], device="cuda" if cuda else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_OK" if int(flag) == 1 else "DIST_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
