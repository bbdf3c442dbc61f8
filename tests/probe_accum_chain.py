"""TEST INFRASTRUCTURE (not collected by pytest): per-parameter gradient error of ONE full-size step against the float64
oracle as a function of how many tensor-core MMAs accumulate into one TMEM accumulator, plus the device time of a step.

    python tests/probe_accum_chain.py T "default" "RAE_TC_DC_SPLITS=9" "RAE_TC_DC_SPLITS=9,RAE_TC_DQ_SPLITS=9"

Writes one JSON line per variant (stdout and gpurun_out/accum_chain.jsonl).  The error is read from the AdaGrad
accumulators after a first step from zero (acc' = g*g), see tests/test_scale_fullsize.py.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import rae_oracle as O                      # noqa: E402
from tests import test_scale_fullsize as TF                   # noqa: E402


def errors(p1, acc1, g_ref, touched):
    out = {}
    for n, g in g_ref.items():
        a1 = acc1[n]
        if n in touched:
            g, a1 = g[touched[n]], a1[touched[n]]
        gmax = float(np.abs(g).max())
        out[n] = float(np.abs(np.sqrt(a1.astype(np.float64)) - np.abs(g)).max() / gmax)
    return out


def main():
    import torch
    from relation_autoencoder_b200.engine import Engine
    name = sys.argv[1]
    variants = sys.argv[2:] or ["default"]
    model, K, d, S, B, F, N, fbar = TF.FULL[name]
    nb = 24
    data, p, neg1, neg2 = TF.make_full_problem(model, K, d, S, B, F, N, fbar, n_batches=nb)
    p64 = {k: v.astype(np.float64) for k, v in p.items()}
    ip = data.indptr[:B + 1]
    c_ref, q_ref, g = O.cost_and_grads(model, p64, ip, data.indices[:ip[-1]], data.args1[:B], data.args2[:B], neg1[:, :B],
                                       neg2[:, :B], alpha=1.0)
    del p64
    touched = TF.touched_rows(data, neg1, neg2, B)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "accum_chain.jsonl"), "a")
    for var in variants:
        env = dict(kv.split("=") for kv in var.split(",")) if var != "default" else {}
        for k in ("RAE_TC_DC_SPLITS", "RAE_TC_DQ_SPLITS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        eng = Engine(model, K, d, S, B, F, N, data.n, lr=TF.LR, alpha=1.0, flags=0)
        eng.set_params_numpy(p)
        eng.bind_split("train", data.indptr, data.indices, data.args1, data.args2)
        eng.bind_epoch_negatives(neg1, neg2)
        cost = eng.train_device(0)
        err = errors(eng.get_params_numpy(), eng.get_acc_numpy(), g, touched)
        for b in range(1, 4):
            eng.train_device(b, want_cost=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in range(4, nb):
            eng.train_device(b, want_cost=False)
        e1.record()
        torch.cuda.synchronize()
        line = {"workload": name, "variant": var, "cost_err": abs(cost - c_ref), "grad_err_over_gmax": err,
                "ms_per_step": e0.elapsed_time(e1) / (nb - 4)}
        eng.close()
        print(json.dumps(line), flush=True)
        log.write(json.dumps(line) + "\n")
        log.flush()


if __name__ == "__main__":
    main()
