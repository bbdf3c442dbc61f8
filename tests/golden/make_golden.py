#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory (run from the repo root: ``python tests/golden/make_golden.py``).

The reference's arithmetic lives in Theano, which is not installable here ("parity unpinned"), and the
reference ships no golden vector for this path.  These fixtures therefore freeze, per decoder (A / C / AC):

  * the inputs: parameters in the reference's shapes, one binary-CSR batch, args, INJECTED negative ids;
  * ``cost_torch`` / ``g_torch_*``: cost and gradients from torch-float64 AUTOGRAD of an op-by-op transcription of the
    Theano graph (``tests/helpers.torch_reference_cost``) - independent of the oracle's closed forms;
  * ``cost`` / ``q`` / ``g_*`` / ``p1_*`` / ``acc1_*``: the oracle's cost, q(r|x), gradients and the parameters and
    accumulators after ONE dense AdaGrad step (Optimizers.py:29-32);
  * ``labels``: argmax labels of the label callable (RelationClassifier.py:47).

and, in ``negatives.npz``, the legacy-RandomState negative-sample ids of NegativeExampleGenerator.py:14-32 for a fixed
frequency table (bit-exact contract), generated once with the vectorised searchsorted and once element-wise as the
reference's ``map`` does.

``tests/test_golden.py`` checks the oracle against these files on the CPU and the CUDA path against them on the GPU.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import rae_oracle as O  # noqa: E402
from tests.helpers import make_problem, torch_reference_cost  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (model, shape, seed, options)
    "A_tiny": ("rescal", dict(B=12, K=5, d=6, S=3, F=40, N=25, fbar=4), 21, dict()),
    "C_tiny": ("sp", dict(B=12, K=5, d=6, S=3, F=40, N=25, fbar=4), 22, dict()),
    "AC_tiny": ("rescal+sp", dict(B=12, K=5, d=6, S=3, F=40, N=25, fbar=4), 23, dict()),
    "AC_dups_l2": ("rescal+sp", dict(B=20, K=10, d=10, S=5, F=60, N=30, fbar=6), 24,
                   dict(dup_heavy=True, l2=0.1, alpha=0.1)),                     # README flags: l2 0.1, alpha 0.1
    "A_odd": ("rescal", dict(B=9, K=7, d=33, S=2, F=50, N=20, fbar=5), 25, dict(empty_rows=True)),
    "C_l1": ("sp", dict(B=10, K=6, d=8, S=2, F=30, N=16, fbar=4), 26, dict(l1=0.02, l2=0.05)),
}
ADJ = 0.37
LR = 0.1


def make_case(name):
    model, sh, seed, opt = CASES[name]
    l1, l2, alpha = opt.get("l1", 0.0), opt.get("l2", 0.0), opt.get("alpha", 1.0)
    pr = make_problem(model, seed=seed, dup_heavy=opt.get("dup_heavy", False), empty_rows=opt.get("empty_rows", False), **sh)
    # fp32-representable parameters so the GPU (fp32 storage) starts from identical values
    p0 = {k: v.astype(np.float32).astype(np.float64) for k, v in pr["p"].items()}
    out = dict(model=np.array(O.MODEL_ALIASES[model]), K=sh["K"], d=sh["d"], S=sh["S"], B=sh["B"], F=sh["F"], N=sh["N"],
               l1=l1, l2=l2, alpha=alpha, adj=ADJ, lr=LR,
               indptr=pr["indptr"], indices=pr["indices"], a1=pr["a1"], a2=pr["a2"], neg1=pr["neg1"], neg2=pr["neg2"])
    for k, v in p0.items():
        out["p0_" + k] = v
    # independent pin: torch autograd of the op-by-op graph
    cost_t, leaves, q_t = torch_reference_cost(model, p0, pr["indptr"], pr["indices"], pr["a1"], pr["a2"], pr["neg1"],
                                               pr["neg2"], alpha, l1, l2, ADJ, True)
    cost_t.backward()
    out["cost_torch"] = float(cost_t.detach())
    out["q_torch"] = q_t.detach().numpy()
    for k in p0:
        out["g_torch_" + k] = leaves[k].grad.numpy() if leaves[k].grad is not None else np.zeros_like(p0[k])
    # oracle: one reference-faithful (dense AdaGrad) step
    om = O.OracleModel(model, {k: v.copy() for k, v in p0.items()}, K=sh["K"], d=sh["d"], S=sh["S"], B=sh["B"], lr=LR,
                       l1=l1, l2=l2, alpha=alpha)
    labels, _ = O.label_batch(p0["W"], p0["Wb"], pr["indptr"], pr["indices"])
    out["labels"] = labels
    out["cost"] = om.train_explicit(pr["indptr"], pr["indices"], pr["a1"], pr["a2"], pr["neg1"], pr["neg2"], adj=ADJ)
    out["q"] = om.last_q
    for k in p0:
        out["g_" + k] = om.last_grads[k]
        out["p1_" + k] = om.params[k]
        out["acc1_" + k] = om.acc[k]
    return out


def make_negatives():
    rs = np.random.RandomState(7)
    freqs = rs.zipf(1.6, size=97).astype(np.float64)
    cum = O.neg_sampling_cum(freqs)
    rng = np.random.RandomState(2)                   # seed 2 = the CLI default (OieInduction.py:476)
    smp = O.NegativeSampler(rng, cum)
    n1 = smp.get_negative_samples(50, 5)             # side 1 first, then side 2 (OieInduction.py:183-184)
    n2 = smp.get_negative_samples(50, 5)
    # element-wise restatement of the reference's map(lambda x: cum.searchsorted(x), u) on a fresh, identical stream
    rng2 = np.random.RandomState(2)
    u1 = rng2.uniform(0, cum[-1], 250)
    m1 = np.array([cum.searchsorted(x) for x in u1], dtype=np.int32).reshape(5, 50)
    assert np.array_equal(m1, n1)
    return dict(freqs=freqs, cum=cum, neg1=n1, neg2=n2)


def main():
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **make_case(name))
        print("wrote", name)
    np.savez_compressed(os.path.join(HERE, "negatives.npz"), **make_negatives())
    print("wrote negatives")


if __name__ == "__main__":
    main()
