"""Synthetic workloads of the shapes BASELINE.json names: there is no network for the NYT corpus.

Data: ``np.random.default_rng(1234)``; features per example ``nnz ~ clip(Poisson(fbar), 1, 4 fbar)`` unique ids drawn
from a Zipf(1.0)-over-rank distribution on [0, F), sorted within the row (binary CSR, OieData.py:83-90); entity ids from
Zipf(1.0) over [0, N).  Parameters and negatives: legacy ``np.random.RandomState(2)`` exactly as the reference draws
them (init order RelationClassifier.py:24 -> OieModel.py:105 -> decoder normals; negatives by the inverse-CDF recipe of
NegativeExampleGenerator.py:24,32 over freq**0.75 of the realised entity counts, OieData.py:57-59).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

# name -> sizes.  cfg2..cfg5 and T are BASELINE.json configs[1..4] + the north-star target.
WORKLOADS = {
    "cfg1": dict(model="rescal+sp", K=10, d=10, S=5, B=100, fbar=14, F=6276, N=1476, N_train=1000, l2=0.1, alpha=0.1,
                 desc="README run sizes on synthetic data (AC, K=10, d=10, B=100, S=5, l2=0.1)"),
    "cfg2": dict(model="rescal+sp", K=100, d=30, S=5, B=4096, fbar=30, F=1_000_000, N=500_000, N_train=2_000_000,
                 desc="synthetic NYT-scale AC: 2M examples, K=100, d=30, ~30 features/example, 1M feature vocab, 500k entities"),
    "cfg3": dict(model="rescal", K=100, d=128, S=20, B=4096, fbar=30, F=1_000_000, N=500_000, N_train=2_000_000,
                 desc="model A (bilinear only): K=100, d=128, 20 negatives, batch 4096"),
    "cfg4": dict(model="sp", K=100, d=128, S=20, B=4096, fbar=30, F=1_000_000, N=10_000_000, N_train=2_000_000,
                 desc="model C (selectional preferences): 10M-entity table, K=100, d=128, 20 negatives, batch 4096/GPU"),
    "cfg5": dict(model="rescal+sp", K=1000, d=256, S=50, B=2048, fbar=30, F=1_000_000, N=500_000, N_train=2_000_000,
                 desc="AC at scale: K=1000, d=256, 50 negatives, batch 16384 global (2048/GPU at 8 GPUs)"),
    "T": dict(model="rescal+sp", K=100, d=128, S=20, B=4096, fbar=30, F=1_000_000, N=500_000, N_train=2_000_000,
              desc="north-star target: AC, K=100, d=128, 20 negatives, batch 4096"),
}


@dataclass
class SyntheticData:
    indptr: np.ndarray      # int32 [n+1]
    indices: np.ndarray     # int32 [nnz]
    args1: np.ndarray       # int32 [n]
    args2: np.ndarray       # int32 [n]
    neg_cum: np.ndarray     # float64 [N] cumulative freq**0.75 distribution
    n: int


def _zipf_cdf(n: int) -> np.ndarray:
    w = 1.0 / np.arange(1, n + 1, dtype=np.float64)
    c = np.cumsum(w)
    return c / c[-1]


def make_dataset(n: int, F: int, N: int, fbar: int, seed: int = 1234, uniform: bool = False) -> SyntheticData:
    """``n`` examples.  Rows of the CSR hold unique, ascending feature ids."""
    rng = np.random.default_rng(seed)
    nnz_draw = np.clip(rng.poisson(fbar, size=n), 1, 4 * fbar).astype(np.int64)
    total = int(nnz_draw.sum())
    u = rng.random(total)
    if uniform:
        ids = np.minimum((u * F).astype(np.int64), F - 1)
    else:
        ids = np.searchsorted(_zipf_cdf(F), u).astype(np.int64)
        np.minimum(ids, F - 1, out=ids)
    rows = np.repeat(np.arange(n, dtype=np.int64), nnz_draw)
    key = np.unique(rows * F + ids)             # sorts by (row, id) and drops within-row duplicates
    rows_u = key // F
    indices = (key - rows_u * F).astype(np.int32)
    counts = np.bincount(rows_u, minlength=n)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    assert indptr[-1] < 2 ** 31
    ecdf = _zipf_cdf(N)
    a1 = np.minimum(np.searchsorted(ecdf, rng.random(n)), N - 1).astype(np.int32)
    a2 = np.minimum(np.searchsorted(ecdf, rng.random(n)), N - 1).astype(np.int32)
    freq = (np.bincount(a1, minlength=N) + np.bincount(a2, minlength=N)).astype(np.float64)
    f = freq ** 0.75                             # OieData.py:36,118
    cum = np.cumsum(f / float(np.cumsum(f)[-1]))  # OieData.py:57-59
    return SyntheticData(indptr.astype(np.int32), indices, a1, a2, cum, n)


def draw_negatives(rng: np.random.RandomState, cum: np.ndarray, n: int, S: int):
    """One epoch's negatives, side 1 then side 2 (OieInduction.py:183-184); int32 [S, n] each."""
    def one():
        uu = rng.uniform(0, cum[-1], S * n)                      # NegativeExampleGenerator.py:32
        return np.asarray(cum.searchsorted(uu), dtype=np.int32).reshape((S, n))   # :24
    return one(), one()


def init_params(rng: np.random.RandomState, model: str, F: int, K: int, N: int, d: int, dtype=np.float32) -> Dict[str, np.ndarray]:
    """Reference initialisation in the reference's RNG order (see module docstring); cast to ``dtype``."""
    p: Dict[str, np.ndarray] = {}
    p["W"] = rng.uniform(low=-1.0e-3, high=1.0e-3, size=(F, K)).astype(dtype)        # settings.py:23-24
    p["Wb"] = np.zeros(K, dtype=dtype)
    p["A"] = rng.uniform(-0.01, 0.01, size=(N, d)).astype(dtype)                     # OieModel.py:105
    sd = math.sqrt(0.1)
    if model in ("rescal", "rescal+sp"):
        p["C"] = rng.normal(0, sd, size=(d, d, K)).astype(dtype)                      # Bilinear.py:14 / BilinearPlusSP.py:14
    if model in ("sp", "rescal+sp"):
        p["C1"] = rng.normal(0, sd, size=(d, K)).astype(dtype)                        # SelectionalPreferences.py:13
        p["C2"] = rng.normal(0, sd, size=(d, K)).astype(dtype)                        # :14
    p["Ab"] = np.zeros(N, dtype=dtype)
    return p


def algorithmic_bytes(model: str, K: int, d: int, S: int, B: int, nnz: float, U_W: float, U_E: float, adagrad: bool = True) -> float:
    """SURVEY 8(d) bytes model for one step (fp32 params + fp32 accumulators, int32 ids)."""
    rmw = 16.0 if adagrad else 8.0
    hasM = model in ("rescal", "rescal+sp")
    hasSP = model in ("sp", "rescal+sp")
    p_dense = (d * d * K if hasM else 0) + (2 * d * K if hasSP else 0) + K
    gath = 4.0 * (2 + 2 * S) * B * (d + 1) if hasM else 4.0 * ((1 + 2 * S) * B * (d + 1) + B)
    return 4.0 * nnz * K + rmw * U_W * K + gath + rmw * U_E * (d + 1) + (4.0 + rmw) * p_dense + 4.0 * (nnz + 2 * B + 2 * S * B)


def algorithmic_flops(model: str, K: int, d: int, S: int, B: int, nnz: float) -> float:
    """6 K d^2 per example for the bilinear models, 12 K d for SP, 2 fbar K for the encoder."""
    hasM = model in ("rescal", "rescal+sp")
    hasSP = model in ("sp", "rescal+sp")
    return (6.0 * K * d * d * B if hasM else 0.0) + (12.0 * K * d * B if hasSP else 0.0) + 2.0 * nnz * K
