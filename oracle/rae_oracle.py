"""CPU oracle for the relation-autoencoder training hot path (TEST INFRASTRUCTURE ONLY).

NumPy float64 restatement of the reference's cost graph and optimiser, used as the checker for the
CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module; the product package must never route through it.

PARITY UNPINNED: the reference's arithmetic lives in Theano (un-vendored, version unpinned, ``README.md:8``),
which cannot be installed in this environment (no wheel, no network, Python-2 sources), and the reference's own
tests (``test.py:29-52``) hold no golden vector or numeric assertion for this path.  This oracle is therefore
pinned only by (i) a torch-float64 autograd run of an independent op-by-op transcription of the decoders,
(ii) central finite differences, (iii) dense-vs-sparse-row update equivalence (see ``tests/test_oracle.py``) and
(iv) the committed golden fixtures ``tests/golden/*.npz`` / ``readme_run*.json`` it must keep reproducing
(``tests/test_golden.py``, ``tests/test_driver.py``).

Every function cites the reference file:line it follows (paths relative to the reference root).

Notation: B=batch (l), K=relations (m), d=embed (r), S=negatives (s), F=feature dim, N=#entities.
Parameter shapes/orders are the reference's: W[F,K], Wb[K], A[N,d], Ab[N], C (or R)[d,d,K], C1[d,K], C2[d,K].
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

MODEL_A = "rescal"        # Bilinear.py:7
MODEL_C = "sp"            # SelectionalPreferences.py:6
MODEL_AC = "rescal+sp"    # BilinearPlusSP.py:7
MODELS = (MODEL_A, MODEL_C, MODEL_AC)

# README.md:44 spelling -> this fork's spelling (Decoder.py:85-93)
MODEL_ALIASES = {"A": MODEL_A, "C": MODEL_C, "AC": MODEL_AC,
                 MODEL_A: MODEL_A, MODEL_C: MODEL_C, MODEL_AC: MODEL_AC}

INIT_LOW, INIT_HIGH = -1.0e-3, 1.0e-3   # settings.py:23-24


# ----------------------------------------------------------------------------------------------------------
# parameter initialisation (rng draw order is part of the contract: one RandomState feeds everything)
# ----------------------------------------------------------------------------------------------------------
def param_names(model: str):
    """Parameter list order = encoder params then decoder ``get_parameters()``.

    RelationClassifier.py:26 ([W, Wb]); Bilinear.py:20 ([R, A, Ab]); SelectionalPreferences.py:22
    ([A, C1, C2, Ab]); BilinearPlusSP.py:32 ([C, A, Ab, C1, C2]).  ``R`` is stored under the key ``C``.
    """
    model = MODEL_ALIASES[model]
    if model == MODEL_A:
        return ["W", "Wb", "C", "A", "Ab"]
    if model == MODEL_C:
        return ["W", "Wb", "A", "C1", "C2", "Ab"]
    return ["W", "Wb", "C", "A", "Ab", "C1", "C2"]


def init_params(rng: np.random.RandomState, model: str, F: int, K: int, N: int, d: int) -> Dict[str, np.ndarray]:
    """Draw the parameters in the reference's order from ONE legacy RandomState.

    Order (OieModel.py:50,58-59): W uniform(low, high)[F,K] (RelationClassifier.py:24), Wb zeros (:25);
    A uniform(-0.01, 0.01)[N,d] (OieModel.py:105); then the decoder normals with sigma = sqrt(0.1):
    R[d,d,K] (Bilinear.py:14) | C1,C2[d,K] (SelectionalPreferences.py:13-14) | C[d,d,K],C1,C2 (BilinearPlusSP.py:14-17).
    Ab zeros[N].
    """
    model = MODEL_ALIASES[model]
    p: Dict[str, np.ndarray] = {}
    p["W"] = np.asarray(rng.uniform(low=INIT_LOW, high=INIT_HIGH, size=(F, K)), dtype=np.float64)
    p["Wb"] = np.zeros(K, dtype=np.float64)
    p["A"] = np.asarray(rng.uniform(-0.01, 0.01, size=(N, d)), dtype=np.float64)
    sd = math.sqrt(0.1)
    if model in (MODEL_A, MODEL_AC):
        p["C"] = np.asarray(rng.normal(0, sd, size=(d, d, K)), dtype=np.float64)
    if model in (MODEL_C, MODEL_AC):
        p["C1"] = np.asarray(rng.normal(0, sd, size=(d, K)), dtype=np.float64)
        p["C2"] = np.asarray(rng.normal(0, sd, size=(d, K)), dtype=np.float64)
    p["Ab"] = np.zeros(N, dtype=np.float64)
    return p


# ----------------------------------------------------------------------------------------------------------
# negative sampling  (NegativeExampleGenerator.py:14-32, OieData.py:57-59)
# ----------------------------------------------------------------------------------------------------------
def neg_sampling_cum(entity_freqs: np.ndarray, power: float = 0.75) -> np.ndarray:
    """freq**power normalised then ``np.cumsum`` in float64 (OieData.py:57-59, :117-118)."""
    f = np.asarray(entity_freqs, dtype=np.float64) ** power
    # the reference sums python floats left-to-right (OieData.py:57); np.cumsum accumulates in the same order
    # (np.sum would be pairwise and can differ in the last ulp)
    norm1 = float(np.cumsum(f)[-1])
    distr = f / norm1
    return np.cumsum(distr)


class NegativeSampler:
    """``ids = cum.searchsorted(rng.uniform(0, cum[-1], S*n))`` -> int32 -> reshape (S, n).

    NegativeExampleGenerator.py:24 (reshape) and :32 (element-wise ``searchsorted``, side='left', then int32).
    The vectorised searchsorted is element-wise identical to the reference's ``map``.
    """

    def __init__(self, rng: np.random.RandomState, cum: np.ndarray):
        assert abs(cum[-1] - 1) < 1.0e-4   # NegativeExampleGenerator.py:12
        self._rand = rng
        self._cum = np.asarray(cum, dtype=np.float64)

    def get_negative_samples(self, num_positive_entities: int, num_negative_samples: int) -> np.ndarray:
        n = num_positive_entities * num_negative_samples
        u = self._rand.uniform(0, self._cum[-1], n)
        ids = np.asarray(self._cum.searchsorted(u), dtype=np.int32)
        return ids.reshape((num_negative_samples, num_positive_entities))


# ----------------------------------------------------------------------------------------------------------
# encoder  (RelationClassifier.py:28-48)
# ----------------------------------------------------------------------------------------------------------
def encoder_scores(W, Wb, indptr, indices) -> np.ndarray:
    """``sparse.dot(x_feats, W) + Wb`` for a binary CSR batch (RelationClassifier.py:35,45; data all 1.0, OieData.py:88)."""
    B = len(indptr) - 1
    z = np.empty((B, W.shape[1]), dtype=np.float64)
    for b in range(B):
        z[b] = W[indices[indptr[b]:indptr[b + 1]]].sum(axis=0)
    return z + Wb


def encoder_scores_fast(W, Wb, indptr, indices) -> np.ndarray:
    """Same as :func:`encoder_scores` through scipy CSR (used where the python loop would be slow)."""
    import scipy.sparse as sp
    B = len(indptr) - 1
    x = sp.csr_matrix((np.ones(len(indices), dtype=np.float64), np.asarray(indices), np.asarray(indptr)),
                      shape=(B, W.shape[0]))
    return np.asarray(x @ W) + Wb


def softmax_rows(z: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """``T.nnet.softmax`` (RelationClassifier.py:36,46): returns (q, log q)."""
    zs = z - z.max(axis=1, keepdims=True)
    lse = np.log(np.exp(zs).sum(axis=1, keepdims=True))
    logq = zs - lse
    return np.exp(logq), logq


def label_batch(W, Wb, indptr, indices) -> Tuple[np.ndarray, np.ndarray]:
    """``comp_probs_and_labels`` (RelationClassifier.py:39-48): labels = argmax of the *scores*, first max wins."""
    z = encoder_scores_fast(W, Wb, indptr, indices)
    q, _ = softmax_rows(z)
    return np.argmax(z, axis=1).astype(np.int64), q


# ----------------------------------------------------------------------------------------------------------
# decoders: op-by-op forward (follows the Theano expressions one line at a time)
# ----------------------------------------------------------------------------------------------------------
def _log_sigmoid(x):
    """``T.log(T.nnet.sigmoid(x))`` evaluated stably as -softplus(-x) (Theano applies the same rewrite)."""
    return -np.logaddexp(0.0, -x)


def _sigmoid(x):
    return np.exp(-np.logaddexp(0.0, -x))


def decoder_scores(model, p, a1, a2, q, neg1, neg2, entropy) -> np.ndarray:
    """``get_scores`` of the selected decoder; returns the concatenated score vector of length 4B + 2BS.

    A : Bilinear.py:28-49 (+ helpers :51-79).  C : SelectionalPreferences.py:30-51 (incl. the ``args1`` quirk at :35).
    AC: BilinearPlusSP.py:34-57 (+ helpers :59-102).
    """
    model = MODEL_ALIASES[model]
    A, Ab = p["A"], p["Ab"]
    S, B = neg1.shape
    d = A.shape[1]
    if model == MODEL_A:
        e1 = A[a1]                                              # Bilinear.py:30
        e2 = A[a2]                                              # :31
        wR = np.tensordot(q, p["C"], axes=[[1], [2]])           # :33  (l,r,r)
        afirst = np.einsum("bij,bi->bj", wR, e1)                # :58
        one = np.einsum("bj,bj->b", afirst, e2)                 # :59
        u = np.concatenate([one + Ab[a1], one + Ab[a2]])        # :36
        all_scores = np.concatenate([_log_sigmoid(u), entropy, entropy])   # :38-39
        x = A[neg1.reshape(-1)].reshape(S, B, d)                # :41
        y = A[neg2.reshape(-1)].reshape(S, B, d)                # :42
        t = np.einsum("bij,bis->bjs", wR, x.transpose(1, 2, 0))  # :68
        neg_one = np.einsum("bjs,bj->bs", t, e2)                # :69  (l,s)
        neg_two = np.einsum("bj,bjs->bs", afirst, y.transpose(1, 2, 0))   # :78-79
        g = np.concatenate([neg_one + Ab[neg1].T, neg_two + Ab[neg2].T])  # :46  (2l,s)
        return np.concatenate([all_scores, _log_sigmoid(-g).reshape(-1)])  # :47-48
    if model == MODEL_C:
        wC1 = q @ p["C1"].T                                     # SelectionalPreferences.py:31
        wC2 = q @ p["C2"].T                                     # :32
        left = np.einsum("bj,bj->b", wC1, A[a1])                # :34
        right = np.einsum("bj,bj->b", wC2, A[a1])               # :35  (sic: args1)
        one = left + right                                      # :36
        u = np.concatenate([one + Ab[a1], one + Ab[a2]])        # :38
        all_scores = np.concatenate([_log_sigmoid(u), entropy, entropy])   # :39
        x = A[neg1.reshape(-1)].reshape(S, B, d)                # :41
        y = A[neg2.reshape(-1)].reshape(S, B, d)                # :42
        nl = np.einsum("bj,bjs->bs", wC1, x.transpose(1, 2, 0))  # :43
        nr = np.einsum("bj,bjs->bs", wC2, y.transpose(1, 2, 0))  # :44
        neg_one = nl.T + right                                  # :46  (s,l)
        neg_two = nr.T + left                                   # :47
        g = np.concatenate([neg_one + Ab[neg1], neg_two + Ab[neg2]])      # :48  (2s,l)
        return np.concatenate([all_scores, _log_sigmoid(-g).reshape(-1)])  # :49-50
    # AC
    wC1 = q @ p["C1"].T                                         # BilinearPlusSP.py:35
    wC2 = q @ p["C2"].T                                         # :36
    wC = np.tensordot(q, p["C"], axes=[[1], [2]])               # :37
    e1 = A[a1]                                                  # :39
    e2 = A[a2]                                                  # :40
    afirst = np.einsum("bij,bi->bj", wC, e1)                    # :68
    asecond = np.einsum("bj,bj->b", afirst, e2)                 # :69
    sp_first = np.einsum("bj,bj->b", wC1, e1)                   # :70
    sp_second = np.einsum("bj,bj->b", wC2, e2)                  # :71
    one = asecond + sp_first + sp_second                        # :72
    u = np.concatenate([one + Ab[a1], one + Ab[a2]])            # :44
    all_scores = np.concatenate([_log_sigmoid(u), entropy, entropy])      # :45-47
    x = A[neg1.reshape(-1)].reshape(S, B, d)                    # :49
    y = A[neg2.reshape(-1)].reshape(S, B, d)                    # :50
    xt = x.transpose(1, 2, 0)
    yt = y.transpose(1, 2, 0)
    t = np.einsum("bij,bis->bjs", wC, xt)                       # :83
    n1 = (np.einsum("bjs,bj->bs", t, e2)                        # :84
          + np.einsum("bj,bjs->bs", wC1, xt)                    # :85
          + sp_second.reshape(B, 1))                            # :86-87
    n2 = (np.einsum("bj,bjs->bs", afirst, yt)                   # :98-99
          + np.einsum("bj,bjs->bs", wC2, yt)                    # :101
          + sp_first.reshape(B, 1))                             # :100,102
    g = np.concatenate([n1 + Ab[neg1].T, n2 + Ab[neg2].T])      # :54
    return np.concatenate([all_scores, _log_sigmoid(-g).reshape(-1)])      # :55-56


def regulariser_terms(model, p, ext_reg: bool) -> Tuple[float, float]:
    """(L1, L2): W always (OieModel.py:54-56); decoder terms when ``ext_reg`` (OieModel.py:60-62;
    Bilinear.py:22-26, SelectionalPreferences.py:24-28, BilinearPlusSP.py:26-30).  A, Ab, Wb never."""
    model = MODEL_ALIASES[model]
    names = ["W"]
    if ext_reg:
        names += {MODEL_A: ["C"], MODEL_C: ["C1", "C2"], MODEL_AC: ["C1", "C2", "C"]}[model]
    l1 = float(sum(np.abs(p[n]).sum() for n in names))
    l2 = float(sum(np.square(p[n]).sum() for n in names))
    return l1, l2


def regularised_names(model, ext_reg: bool):
    model = MODEL_ALIASES[model]
    names = ["W"]
    if ext_reg:
        names += {MODEL_A: ["C"], MODEL_C: ["C1", "C2"], MODEL_AC: ["C1", "C2", "C"]}[model]
    return names


def train_cost(model, p, indptr, indices, a1, a2, neg1, neg2, alpha, l1=0.0, l2=0.0, adj=1.0, ext_reg=True):
    """Regularised batch cost: ``-mean(all_scores) + l1*L1*adj + l2*L2*adj``.

    OieModel.py:80-82,90 (encoder, entropy, -mean) and OieInduction.py:131,134-135 (adjust, regulariser).
    """
    z = encoder_scores_fast(p["W"], p["Wb"], indptr, indices)
    q, logq = softmax_rows(z)
    entropy = alpha * -(logq * q).sum(axis=1)                    # OieModel.py:81
    scores = decoder_scores(model, p, a1, a2, q, neg1, neg2, entropy)
    cost = -scores.mean()                                        # OieModel.py:90
    if l1 != 0.0 or l2 != 0.0:
        L1, L2 = regulariser_terms(model, p, ext_reg)
        cost = cost + l1 * L1 * adj + l2 * L2 * adj              # OieInduction.py:134-135
    return float(cost), q


# ----------------------------------------------------------------------------------------------------------
# closed-form forward + backward (what T.grad produces, Optimizers.py:27,48); dense gradients
# ----------------------------------------------------------------------------------------------------------
def cost_and_grads(model, p, indptr, indices, a1, a2, neg1, neg2, alpha, l1=0.0, l2=0.0, adj=1.0,
                   ext_reg=True, fix_sp_quirk=False, z_total: Optional[int] = None):
    """Returns (cost, q, grads) with dense float64 gradients shaped like the parameters.

    ``z_total`` overrides the mean's denominator Z = 4B + 2BS (used by the multi-GPU tests where each rank holds a
    slice of a global batch: the local sums are divided by the GLOBAL Z).  Duplicate rows accumulate
    (``AdvancedIncSubtensor1`` semantics) via ``np.add.at``.
    """
    model = MODEL_ALIASES[model]
    W, Wb, A, Ab = p["W"], p["Wb"], p["A"], p["Ab"]
    S, B = neg1.shape
    K = W.shape[1]
    d = A.shape[1]
    Z = float(4 * B + 2 * B * S) if z_total is None else float(z_total)
    has_M = model in (MODEL_A, MODEL_AC)
    has_sp = model in (MODEL_C, MODEL_AC)

    z = encoder_scores_fast(W, Wb, indptr, indices)
    q, logq = softmax_rows(z)
    ent = alpha * -(logq * q).sum(axis=1)

    e1 = A[a1]
    e2 = A[a2]
    Lv = e1
    Rv = e1 if (model == MODEL_C and not fix_sp_quirk) else e2     # SelectionalPreferences.py:35 quirk
    x = A[neg1]            # (S,B,d)
    y = A[neg2]
    zero = np.zeros((B, d))
    if has_M:
        M = np.tensordot(q, p["C"], axes=[[1], [2]])               # (B,d,d)
        v = np.einsum("bij,bj->bi", M, Rv)
        w = np.einsum("bij,bi->bj", M, Lv)
    else:
        M = None
        v = zero
        w = zero
    if has_sp:
        c1 = q @ p["C1"].T
        c2 = q @ p["C2"].T
    else:
        c1 = zero
        c2 = zero
    pos = (Lv * v).sum(1) + (c1 * Lv).sum(1) + (c2 * Rv).sum(1)
    u1 = pos + Ab[a1]
    u2 = pos + Ab[a2]
    n1 = np.einsum("sbi,bi->sb", x, v + c1) + (c2 * Rv).sum(1)[None, :] + Ab[neg1]
    n2 = np.einsum("sbj,bj->sb", y, w + c2) + (c1 * Lv).sum(1)[None, :] + Ab[neg2]
    total = (_log_sigmoid(u1).sum() + _log_sigmoid(u2).sum() + 2.0 * ent.sum()
             + _log_sigmoid(-n1).sum() + _log_sigmoid(-n2).sum())
    cost = -total / Z

    gu1 = -_sigmoid(-u1) / Z
    gu2 = -_sigmoid(-u2) / Z
    gp = gu1 + gu2
    gn1 = _sigmoid(n1) / Z          # (S,B)
    gn2 = _sigmoid(n2) / Z
    G1 = gn1.sum(0)
    G2 = gn2.sum(0)
    X1 = np.einsum("sb,sbi->bi", gn1, x)
    Y2 = np.einsum("sb,sbj->bj", gn2, y)
    a_ = gp[:, None] * Lv + X1
    c_ = gp[:, None] * Rv + Y2

    g: Dict[str, np.ndarray] = {}
    dq = (2.0 * alpha / Z) * (logq + 1.0)
    gA = np.zeros_like(A)
    gAb = np.zeros_like(Ab)
    gL = np.zeros((B, d))
    gR = np.zeros((B, d))
    gx = np.zeros((S, B, d))
    gy = np.zeros((S, B, d))
    if has_M:
        dM = np.einsum("bi,bj->bij", a_, Rv) + np.einsum("bi,bj->bij", Lv, Y2)
        g["C"] = np.einsum("bij,bk->ijk", dM, q)
        dq = dq + np.einsum("bij,ijk->bk", dM, p["C"])
        gL += np.einsum("bij,bj->bi", M, c_)
        gR += np.einsum("bij,bi->bj", M, a_)
        gx += gn1[:, :, None] * v[None, :, :]
        gy += gn2[:, :, None] * w[None, :, :]
    if has_sp:
        dc1 = a_ + G2[:, None] * Lv
        dc2 = c_ + G1[:, None] * Rv
        g["C1"] = dc1.T @ q
        g["C2"] = dc2.T @ q
        dq = dq + dc1 @ p["C1"] + dc2 @ p["C2"]
        gL += (gp + G2)[:, None] * c1
        gR += (gp + G1)[:, None] * c2
        gx += gn1[:, :, None] * c1[None, :, :]
        gy += gn2[:, :, None] * c2[None, :, :]
    np.add.at(gA, a1, gL)
    np.add.at(gA, a1 if (model == MODEL_C and not fix_sp_quirk) else a2, gR)
    np.add.at(gA, neg1.reshape(-1), gx.reshape(-1, d))
    np.add.at(gA, neg2.reshape(-1), gy.reshape(-1, d))
    np.add.at(gAb, a1, gu1)
    np.add.at(gAb, a2, gu2)
    np.add.at(gAb, neg1.reshape(-1), gn1.reshape(-1))
    np.add.at(gAb, neg2.reshape(-1), gn2.reshape(-1))
    g["A"] = gA
    g["Ab"] = gAb

    dz = q * (dq - (q * dq).sum(axis=1, keepdims=True))
    g["Wb"] = dz.sum(axis=0)
    gW = np.zeros_like(W)
    rows = np.repeat(np.arange(B), np.diff(indptr))
    np.add.at(gW, np.asarray(indices), dz[rows])
    g["W"] = gW

    if l1 != 0.0 or l2 != 0.0:
        L1, L2 = regulariser_terms(model, p, ext_reg)
        cost = cost + l1 * L1 * adj + l2 * L2 * adj
        for n in regularised_names(model, ext_reg):
            g[n] = g[n] + adj * (l1 * np.sign(p[n]) + 2.0 * l2 * p[n])
    return float(cost), q, g


# ----------------------------------------------------------------------------------------------------------
# optimisers (Optimizers.py) - DENSE, exactly as the reference executes them
# ----------------------------------------------------------------------------------------------------------
def adagrad_update(p, acc, g, lr, names):
    """``acc' = acc + g^2 ; p' = p - lr*g/(sqrt(acc') + 1e-6)`` over every element (Optimizers.py:29-32)."""
    for n in names:
        acc[n] = acc[n] + np.square(g[n])
        p[n] = p[n] - lr * g[n] / (np.sqrt(acc[n]) + 1e-6)


def sgd_update(p, g, lr, names):
    """``p' = p - lr*g`` (Optimizers.py:50-51)."""
    for n in names:
        p[n] = p[n] - lr * g[n]


def adagrad_update_sparse_rows(p, acc, g, lr, names, touched: Dict[str, np.ndarray]):
    """Row-sparse AdaGrad: only rows in ``touched[n]`` are visited (rows with g=0 are fixed points of the dense rule)."""
    for n in names:
        if n in touched:
            r = touched[n]
            gr = g[n][r]
            acc[n][r] = acc[n][r] + np.square(gr)
            p[n][r] = p[n][r] - lr * gr / (np.sqrt(acc[n][r]) + 1e-6)
        else:
            acc[n] = acc[n] + np.square(g[n])
            p[n] = p[n] - lr * g[n] / (np.sqrt(acc[n]) + 1e-6)


# ----------------------------------------------------------------------------------------------------------
# the reference's run-time boundary: func['train'] / func['label_<split>']  (OieInduction.py:146-155)
# ----------------------------------------------------------------------------------------------------------
@dataclass
class OracleSplit:
    indptr: np.ndarray
    indices: np.ndarray
    args1: np.ndarray
    args2: np.ndarray


@dataclass
class OracleModel:
    """float64 stand-in for the compiled Theano callables of ``ReconstructInducer`` (OieInduction.py:118-155)."""
    model: str
    params: Dict[str, np.ndarray]
    K: int
    d: int
    S: int
    B: int
    lr: float = 0.1
    l1: float = 0.0
    l2: float = 0.0
    alpha: float = 1.0
    optimizer: str = "adagrad"
    ext_reg: bool = True
    fix_sp_quirk: bool = False
    sparse_rows: bool = False          # False = reference-faithful dense update
    splits: Dict[str, OracleSplit] = field(default_factory=dict)
    acc: Dict[str, np.ndarray] = field(default_factory=dict)
    last_q: Optional[np.ndarray] = None
    last_grads: Optional[Dict[str, np.ndarray]] = None

    def __post_init__(self):
        self.model = MODEL_ALIASES[self.model]
        if not self.acc:
            self.acc = {n: np.zeros_like(v) for n, v in self.params.items()}   # Optimizers.py:12-15

    def bind_split(self, name, indptr, indices, args1, args2):
        self.splits[name] = OracleSplit(np.asarray(indptr), np.asarray(indices), np.asarray(args1), np.asarray(args2))

    @property
    def adj(self):
        return float(self.B) / float(len(self.splits["train"].args1))          # OieInduction.py:131

    def _slice(self, split, batch_index):
        sp_ = self.splits[split]
        lo, hi = batch_index * self.B, (batch_index + 1) * self.B               # OieInduction.py:147-149
        ip = sp_.indptr[lo:hi + 1]
        return ip - ip[0], sp_.indices[ip[0]:ip[-1]], sp_.args1[lo:hi], sp_.args2[lo:hi]

    def train_explicit(self, indptr, indices, a1, a2, neg1, neg2, adj=None, z_total=None):
        adj = self.adj if adj is None else adj
        cost, q, g = cost_and_grads(self.model, self.params, indptr, indices, a1, a2, neg1, neg2, self.alpha,
                                    self.l1, self.l2, adj, self.ext_reg, self.fix_sp_quirk, z_total)
        self.last_q, self.last_grads = q, g
        names = param_names(self.model)
        if self.optimizer == "adagrad":
            if self.sparse_rows and self.l1 == 0.0 and self.l2 == 0.0:
                touched = {"W": np.unique(indices),
                           "A": np.unique(np.concatenate([a1, a2, neg1.reshape(-1), neg2.reshape(-1)]))}
                touched["Ab"] = touched["A"]
                adagrad_update_sparse_rows(self.params, self.acc, g, self.lr, names, touched)
            else:
                adagrad_update(self.params, self.acc, g, self.lr, names)
        elif self.optimizer == "sgd":
            sgd_update(self.params, g, self.lr, names)
        else:
            raise Exception("Optimizer '{}' not implemented".format(self.optimizer))   # OieInduction.py:269
        return cost

    def train(self, batch_index, neg1, neg2):
        """``func['train'](batch_index, neg1[S,B], neg2[S,B]) -> cost`` with the updates as side effect."""
        ip, ix, a1, a2 = self._slice("train", batch_index)
        return self.train_explicit(ip, ix, a1, a2, np.asarray(neg1), np.asarray(neg2))

    def label(self, split, batch_index):
        """``func['label_'+split](batch_index) -> (labels int64[B], probs[B,K])``."""
        ip, ix, _, _ = self._slice(split, batch_index)
        return label_batch(self.params["W"], self.params["Wb"], ip, ix)
