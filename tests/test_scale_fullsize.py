"""Full-size parity: one production step (no debug flags, bound split, device-resident negatives) at the sizes
BASELINE.json names - batch 4096, K=100, d=30 / 128, 5 / 20 negatives, 1 M features, 500 k entities - against the
float64 oracle run on the SAME inputs, plus the size-independent properties of the path.

What a single step from ZERO AdaGrad accumulators exposes without dense gradient buffers (Optimizers.py:12-15,29-32):
  * acc' = g*g                      -> |g| of EVERY parameter element is readable from the accumulators:
                                       || sqrt(acc') - |g_ref| ||_inf <= 1e-5 * ||g_ref||_inf      (gradient tolerance)
  * p' = p - lr*g/(|g| + 1e-6)      -> the move has the gradient's sign (checked where |g_ref| > 1e-3 ||g_ref||_inf) and
                                       the rule's magnitude: | |p - p'| - lr|g|/(|g|+1e-6) | <= 2e-6 * max(|p|, lr)
  * rows no example touched         -> accumulator rows are exactly 0 and parameter rows are BIT-identical
  * sort-by-row layout at full size -> np.argsort(kind='stable') / np.unique, bit-exact
  * same step twice                 -> bit-identical cost and parameters (no atomics anywhere)
  * labels after the step           -> argmax of the float64 scores wherever the top-2 gap is not a rounding tie
cost and q(r|x): |x - ref| <= 1e-5 * max|ref| as everywhere else.

The checker itself is exercised on the CPU tier against an fp32 NumPy emulation of the update rule (and must reject a
corrupted row), so a red GPU run points at the kernels, not at this file.
"""
import numpy as np
import pytest

from oracle import rae_oracle as O
from relation_autoencoder_b200 import synthetic as SY
from tests.helpers import record_err, rel_err

TOL_COST = 1e-5
TOL_Q = 1e-5
TOL_GRAD = 1e-5
TOL_RULE = 2e-6
LR = 0.1
SPARSE = ("W", "A", "Ab")          # row-sparse tables: only touched rows may change

# name -> (decoder, K, d, S, B, F, N, fbar): BASELINE.json configs[1], configs[2], the north-star target, and model C at
# the target's per-example shape (configs[3]'s 10 M-entity table is a sharding case, test_gpu_dist.py)
FULL = {
    "cfg2": ("rescal+sp", 100, 30, 5, 4096, 1_000_000, 500_000, 30),
    "cfg3": ("rescal", 100, 128, 20, 4096, 1_000_000, 500_000, 30),
    "T": ("rescal+sp", 100, 128, 20, 4096, 1_000_000, 500_000, 30),
    "C": ("sp", 100, 128, 20, 4096, 1_000_000, 500_000, 30),
}


def make_full_problem(model, K, d, S, B, F, N, fbar, n_batches=2, seed=77):
    """Synthetic data of bench.py's recipe (Zipf feature / entity ids, freq**0.75 negatives) with parameters scaled away
    from the near-zero reference initialisation so that q(r|x) is not uniform and the sigmoids are not all at 0."""
    data = SY.make_dataset(n_batches * B, F, N, fbar, seed=seed)
    rng = np.random.RandomState(seed)
    p = SY.init_params(rng, model, F, K, N, d)
    p["W"] = (p["W"] * 50.0).astype(np.float32)                  # U(-0.05, 0.05)
    p["A"] = (p["A"] * 50.0).astype(np.float32)                  # U(-0.5, 0.5)
    p["Wb"] = rng.uniform(-0.3, 0.3, size=K).astype(np.float32)
    p["Ab"] = rng.uniform(-0.3, 0.3, size=N).astype(np.float32)
    neg1, neg2 = SY.draw_negatives(rng, data.neg_cum, data.n, S)
    return data, p, neg1, neg2


def touched_rows(data, neg1, neg2, B, batch=0):
    lo, hi = batch * B, (batch + 1) * B
    feats = np.unique(data.indices[data.indptr[lo]:data.indptr[hi]])
    ents = np.unique(np.concatenate([data.args1[lo:hi], data.args2[lo:hi], neg1[:, lo:hi].reshape(-1), neg2[:, lo:hi].reshape(-1)]))
    return {"W": feats, "A": ents, "Ab": ents}


def check_first_step(p0, p1, acc1, g_ref, touched, lr=LR):
    """The property list of the module docstring for one AdaGrad step from zero accumulators.  ``p0, p1, acc1`` are the
    fp32 arrays of the path under test, ``g_ref`` the oracle's float64 gradients, ``touched`` the row ids per sparse
    table.  Raises AssertionError naming the parameter and the property."""
    for n, g in g_ref.items():
        g = np.asarray(g, dtype=np.float64)
        a1 = np.asarray(acc1[n])
        q0, q1 = np.asarray(p0[n]), np.asarray(p1[n])
        assert a1.dtype == np.float32 and q1.dtype == np.float32, n
        if n in touched:
            rows = touched[n]
            mask = np.ones(g.shape[0], dtype=bool)
            mask[rows] = False
            assert not a1[mask].any(), "%s: accumulator of an untouched row changed" % n
            assert np.array_equal(q1[mask], q0[mask]), "%s: untouched row is not bit-identical" % n
            assert not g[mask].any(), "%s: oracle gradient outside the touched rows" % n
            g, a1, q0, q1 = g[rows], a1[rows], q0[rows], q1[rows]
        gmax = float(np.abs(g).max())
        g_abs = np.sqrt(a1.astype(np.float64))
        e = float(np.abs(g_abs - np.abs(g)).max())
        record_err("fullsize", n, e / max(gmax, 1e-300))
        assert e <= TOL_GRAD * gmax, "%s: |g| from the accumulators off by %.3g of ||g||_inf" % (n, e / max(gmax, 1e-300))
        moved = q0.astype(np.float64) - q1.astype(np.float64)
        rule = lr * g_abs / (g_abs + 1e-6)
        scale = max(float(np.abs(q0).max()), lr)
        e = float(np.abs(np.abs(moved) - rule).max())
        assert e <= TOL_RULE * scale, "%s: update magnitude off the AdaGrad rule by %.3g" % (n, e / scale)
        sig = np.abs(g) > 1e-3 * gmax
        assert np.array_equal(np.sign(moved[sig]), np.sign(g[sig])), "%s: update sign differs from the gradient's" % n


def labels_agree(lab, z_ref, rel_gap=1e-5):
    """argmax parity except where the float64 top-2 gap is inside fp32 rounding of the scores."""
    top2 = np.partition(z_ref, -2, axis=1)[:, -2:]
    decisive = (top2[:, 1] - top2[:, 0]) > rel_gap * np.abs(z_ref).max()
    ref = np.argmax(z_ref, axis=1)
    assert decisive.mean() > 0.99
    return np.array_equal(lab[decisive], ref[decisive])


# ----------------------------------------------------------------------------------------------------------------------
# CPU tier: the checker against an fp32 emulation of the rule, and against corrupted results
# ----------------------------------------------------------------------------------------------------------------------
def _emulated_fp32_step(p0, g_ref, lr=LR):
    p1, acc1 = {}, {}
    for n, g in g_ref.items():
        g32 = g.astype(np.float32)
        acc1[n] = g32 * g32
        p1[n] = (p0[n] - np.float32(lr) * g32 / (np.sqrt(acc1[n]) + np.float32(1e-6))).astype(np.float32)
    return p1, acc1


@pytest.mark.parametrize("model", O.MODELS)
def test_first_step_checker_accepts_fp32_rule_and_rejects_corruption(model):
    K, d, S, B, F, N, fbar = 12, 9, 3, 64, 4000, 900, 6
    data, p, neg1, neg2 = make_full_problem(model, K, d, S, B, F, N, fbar, seed=5)
    p64 = {k: v.astype(np.float64) for k, v in p.items()}
    ip = data.indptr[:B + 1]
    _, _, g = O.cost_and_grads(model, p64, ip, data.indices[:ip[-1]], data.args1[:B], data.args2[:B], neg1[:, :B], neg2[:, :B],
                               alpha=1.0)
    touched = touched_rows(data, neg1, neg2, B)
    p1, acc1 = _emulated_fp32_step(p, g)
    check_first_step(p, p1, acc1, g, touched)
    # (a) a touched row whose gradient is 1e-3 off
    bad = {k: v.copy() for k, v in acc1.items()}
    r = touched["A"][0]
    bad["A"][r] *= np.float32(1.002)
    with pytest.raises(AssertionError, match="A: "):
        check_first_step(p, p1, bad, g, touched)
    # (b) an untouched row that moved by one ulp
    untouched = np.setdiff1d(np.arange(F), touched["W"])[0]
    bad = {k: v.copy() for k, v in p1.items()}
    bad["W"][untouched, 0] = np.nextafter(bad["W"][untouched, 0], np.float32(1.0))
    with pytest.raises(AssertionError, match="W: untouched row"):
        check_first_step(p, bad, acc1, g, touched)
    # (c) an update with the wrong sign
    bad = {k: v.copy() for k, v in p1.items()}
    k = int(np.argmax(np.abs(g["Wb"])))
    bad["Wb"][k] = p["Wb"][k] + (p["Wb"][k] - p1["Wb"][k])
    with pytest.raises(AssertionError, match="Wb: update sign"):
        check_first_step(p, bad, acc1, g, touched)


def test_label_comparison_ignores_only_rounding_ties():
    z = np.tile(np.array([[0.0, 1.0, 0.5]]), (400, 1))
    z[0] = [2.0, 2.0 + 1e-9, 0.0]    # a tie within rounding: either label is accepted on this row
    lab = np.ones(400, dtype=np.int64)
    lab[0] = 0
    assert labels_agree(lab, z)
    lab[7] = 2                       # a decisive row with the wrong label
    assert not labels_agree(lab, z)


# ----------------------------------------------------------------------------------------------------------------------
# GPU tier
# ----------------------------------------------------------------------------------------------------------------------
def _gpu_first_step(model, K, d, S, B, F, N, data, p, neg1, neg2, want_layout=False):
    from relation_autoencoder_b200.engine import Engine
    eng = Engine(model, K, d, S, B, F, N, data.n, lr=LR, alpha=1.0, flags=0)
    eng.set_params_numpy(p)
    eng.bind_split("train", data.indptr, data.indices, data.args1, data.args2)
    eng.bind_epoch_negatives(neg1, neg2)
    cost = eng.train_device(0)
    out = dict(cost=cost, q=eng.last_probs(), p1=eng.get_params_numpy(), acc1=eng.get_acc_numpy(), stats=eng.stats())
    if want_layout:
        out["layout"] = eng.entity_segments()
        out["labels"] = eng.label("train", 1)[0]
    eng.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(FULL))
def test_full_size_step_matches_oracle(name):
    model, K, d, S, B, F, N, fbar = FULL[name]
    data, p, neg1, neg2 = make_full_problem(model, K, d, S, B, F, N, fbar)
    out = _gpu_first_step(model, K, d, S, B, F, N, data, p, neg1, neg2, want_layout=True)
    p64 = {k: v.astype(np.float64) for k, v in p.items()}
    ip = data.indptr[:B + 1]
    n1, n2 = neg1[:, :B], neg2[:, :B]
    c_ref, q_ref, g = O.cost_and_grads(model, p64, ip, data.indices[:ip[-1]], data.args1[:B], data.args2[:B], n1, n2, alpha=1.0)
    del p64
    assert abs(out["cost"] - c_ref) <= TOL_COST * max(1.0, abs(c_ref)), (out["cost"], c_ref)
    assert rel_err(out["q"], q_ref) <= TOL_Q
    touched = touched_rows(data, neg1, neg2, B)
    assert out["stats"]["unique_w_rows"] == len(touched["W"])
    assert out["stats"]["unique_e_rows"] == len(touched["A"])
    check_first_step(p, out["p1"], out["acc1"], g, touched)
    # sort-by-row layout of the (2+2S)*B entity occurrences, bit-exact at full size
    rows, occ, seg = out["layout"]
    keys = np.concatenate([data.args1[:B], data.args2[:B], n1.reshape(-1), n2.reshape(-1)])
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(occ, order.astype(np.int32)) and np.array_equal(rows, keys[order])
    _, first = np.unique(keys[order], return_index=True)
    assert np.array_equal(seg[:-1], first.astype(np.int32)) and seg[-1] == len(keys)
    # labels of the NEXT batch from the updated parameters (func['label_train'](1), RelationClassifier.py:45-47)
    ip2 = data.indptr[B:2 * B + 1]
    z = O.encoder_scores_fast(out["p1"]["W"].astype(np.float64), out["p1"]["Wb"].astype(np.float64), ip2 - ip2[0],
                              data.indices[ip2[0]:ip2[-1]])
    assert labels_agree(out["labels"], z)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg2", "T"])
def test_full_size_step_is_bitwise_reproducible(name):
    model, K, d, S, B, F, N, fbar = FULL[name]
    data, p, neg1, neg2 = make_full_problem(model, K, d, S, B, F, N, fbar, seed=78)
    a = _gpu_first_step(model, K, d, S, B, F, N, data, p, neg1, neg2)
    b = _gpu_first_step(model, K, d, S, B, F, N, data, p, neg1, neg2)
    assert a["cost"] == b["cost"]
    for n in a["p1"]:
        assert np.array_equal(a["p1"][n], b["p1"][n]) and np.array_equal(a["acc1"][n], b["acc1"][n]), n
