"""Pins the NumPy float64 oracle (the reference ships no golden vectors; Theano is not installable here):
 (i)   torch-float64 autograd of an independent op-by-op transcription of the Theano graph,
 (ii)  central finite differences,
 (iii) dense (reference, Optimizers.py:29-32) vs sparse-row AdaGrad equivalence,
 plus the op-by-op forward vs the closed-form forward and the negative sampler recipe."""
import numpy as np
import pytest

from oracle import rae_oracle as O
from tests.helpers import make_problem, rel_err, torch_reference_cost

MODELS = [O.MODEL_A, O.MODEL_C, O.MODEL_AC]


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("reg", [(0.0, 0.0, True), (0.01, 0.1, True), (0.0, 0.1, False)])
def test_closed_form_matches_autograd(model, reg):
    l1, l2, ext_reg = reg
    pr = make_problem(model, seed=3, dup_heavy=True)
    alpha, adj = 0.7, 0.25
    cost, q, g = O.cost_and_grads(model, pr["p"], pr["indptr"], pr["indices"], pr["a1"], pr["a2"], pr["neg1"],
                                  pr["neg2"], alpha, l1, l2, adj, ext_reg)
    tcost, t, tq = torch_reference_cost(model, pr["p"], pr["indptr"], pr["indices"], pr["a1"], pr["a2"],
                                        pr["neg1"], pr["neg2"], alpha, l1, l2, adj, ext_reg)
    tcost.backward()
    assert abs(cost - tcost.item()) <= 1e-13 * max(1.0, abs(cost))
    assert rel_err(q, tq.detach().numpy()) < 1e-13
    for n in O.param_names(model):
        tg = t[n].grad.numpy() if t[n].grad is not None else np.zeros_like(pr["p"][n])
        assert rel_err(g[n], tg) < 1e-11, n


@pytest.mark.parametrize("model", MODELS)
def test_forward_transcription_matches_closed_form(model):
    pr = make_problem(model, seed=5)
    c1, _ = O.train_cost(model, pr["p"], pr["indptr"], pr["indices"], pr["a1"], pr["a2"], pr["neg1"], pr["neg2"],
                         0.3, 0.02, 0.05, 0.5, True)
    c2, _, _ = O.cost_and_grads(model, pr["p"], pr["indptr"], pr["indices"], pr["a1"], pr["a2"], pr["neg1"],
                                pr["neg2"], 0.3, 0.02, 0.05, 0.5, True)
    assert abs(c1 - c2) < 1e-13


@pytest.mark.parametrize("model", MODELS)
def test_finite_differences(model):
    pr = make_problem(model, B=6, K=4, d=3, S=2, F=12, N=9, seed=11, dup_heavy=True)
    args = (pr["indptr"], pr["indices"], pr["a1"], pr["a2"], pr["neg1"], pr["neg2"], 1.0, 0.0, 0.03, 0.5, True)
    _, _, g = O.cost_and_grads(model, pr["p"], *args)
    rng = np.random.RandomState(0)
    eps = 1e-6
    for n in O.param_names(model):
        flat = pr["p"][n].reshape(-1)
        for idx in rng.choice(flat.size, size=min(6, flat.size), replace=False):
            old = flat[idx]
            flat[idx] = old + eps
            cp, _ = O.train_cost(model, pr["p"], *args)
            flat[idx] = old - eps
            cm, _ = O.train_cost(model, pr["p"], *args)
            flat[idx] = old
            fd = (cp - cm) / (2 * eps)
            assert abs(fd - g[n].reshape(-1)[idx]) < 1e-7 * max(1.0, abs(fd)), (n, idx)


def test_model_c_quirk_a2_embedding_gets_no_gradient():
    """SelectionalPreferences.py:35 uses A[args1] on the right side: rows only reachable through args2 keep a zero
    embedding gradient while their bias still receives one."""
    pr = make_problem(O.MODEL_C, seed=2)
    pr["a2"][:] = 24          # an id that appears nowhere else
    pr["a1"][pr["a1"] == 24] = 0
    pr["neg1"][pr["neg1"] == 24] = 1
    pr["neg2"][pr["neg2"] == 24] = 1
    _, _, g = O.cost_and_grads(O.MODEL_C, pr["p"], pr["indptr"], pr["indices"], pr["a1"], pr["a2"], pr["neg1"],
                               pr["neg2"], 1.0)
    assert np.all(g["A"][24] == 0.0)
    assert g["Ab"][24] != 0.0
    _, _, gfix = O.cost_and_grads(O.MODEL_C, pr["p"], pr["indptr"], pr["indices"], pr["a1"], pr["a2"], pr["neg1"],
                                  pr["neg2"], 1.0, fix_sp_quirk=True)
    assert np.any(gfix["A"][24] != 0.0)


@pytest.mark.parametrize("model", MODELS)
def test_dense_and_sparse_row_adagrad_agree_bitwise(model):
    """Rows with g = 0 are fixed points of Optimizers.py:29-32, so visiting only touched rows is exact."""
    pr = make_problem(model, seed=7, dup_heavy=True)
    kw = dict(K=5, d=6, S=3, B=12, lr=0.1, alpha=1.0)
    md = O.OracleModel(model, {k: v.copy() for k, v in pr["p"].items()}, sparse_rows=False, **kw)
    ms = O.OracleModel(model, {k: v.copy() for k, v in pr["p"].items()}, sparse_rows=True, **kw)
    for m in (md, ms):
        for step in range(3):
            m.train_explicit(pr["indptr"], pr["indices"], pr["a1"], pr["a2"], np.roll(pr["neg1"], step, 1),
                             pr["neg2"], adj=1.0)
    for n in O.param_names(model):
        assert np.array_equal(md.params[n], ms.params[n]), n
        assert np.array_equal(md.acc[n], ms.acc[n]), n


def test_negative_sampler_recipe_is_elementwise_searchsorted():
    """NegativeExampleGenerator.py:32 maps searchsorted element by element; the vectorised call must be identical,
    ids are int32, shape (S, n) row-major (:24), and the draw order is side 1 then side 2 from ONE RandomState."""
    freqs = np.array([5, 1, 1, 9, 2, 2, 40, 1], dtype=np.float64)
    cum = O.neg_sampling_cum(freqs)
    assert abs(cum[-1] - 1.0) < 1e-12
    r1, r2 = np.random.RandomState(2), np.random.RandomState(2)
    s = O.NegativeSampler(r1, cum)
    n1 = s.get_negative_samples(7, 3)
    n2 = s.get_negative_samples(7, 3)
    u = r2.uniform(0, cum[-1], 21)
    exp1 = np.array([cum.searchsorted(x) for x in u], dtype=np.int32).reshape(3, 7)
    u = r2.uniform(0, cum[-1], 21)
    exp2 = np.array([cum.searchsorted(x) for x in u], dtype=np.int32).reshape(3, 7)
    assert n1.dtype == np.int32 and n1.shape == (3, 7)
    assert np.array_equal(n1, exp1) and np.array_equal(n2, exp2)


def test_init_param_draw_order():
    """W -> A -> decoder normals from one RandomState (RelationClassifier.py:24, OieModel.py:105, BilinearPlusSP.py:14-17)."""
    r = np.random.RandomState(2)
    p = O.init_params(np.random.RandomState(2), "AC", 7, 3, 5, 4)
    W = r.uniform(-1e-3, 1e-3, size=(7, 3))
    A = r.uniform(-0.01, 0.01, size=(5, 4))
    C = r.normal(0, np.sqrt(0.1), size=(4, 4, 3))
    C1 = r.normal(0, np.sqrt(0.1), size=(4, 3))
    C2 = r.normal(0, np.sqrt(0.1), size=(4, 3))
    for a, b in ((p["W"], W), (p["A"], A), (p["C"], C), (p["C1"], C1), (p["C2"], C2)):
        assert np.array_equal(a, b)
    assert not p["Wb"].any() and not p["Ab"].any()


def test_label_is_argmax_of_scores_first_max_wins():
    W = np.zeros((3, 4))
    W[0] = [1.0, 2.0, 2.0, 0.0]
    indptr = np.array([0, 1, 1])
    indices = np.array([0])
    labels, q = O.label_batch(W, np.zeros(4), indptr, indices)
    assert labels.tolist() == [1, 0] and labels.dtype == np.int64      # tie -> first; empty row -> uniform -> 0
    assert np.allclose(q[1], 0.25)
