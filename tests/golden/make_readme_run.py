#!/usr/bin/env python
"""Regenerates ``data_sample_indexed.npz`` and ``readme_run.json`` (run from the repo root, in the build container -
it reads /root/reference/data-sample.txt, which does not exist on the GPU box).

README run (README.md:33-44): the sample corpus is preprocessed three times (train / dev / test) into one pickle, then
``--model AC --epochs 10 --batch_size 100 --relations_number 10 --negative_samples_number 5 --l2_regularization 0.1
--alpha 0.1 --seed 2 --embed_size 10 --learning_rate 0.1`` (AdaGrad).

* ``data_sample_indexed.npz``: the INDEXED dataset (feature-id CSR, entity ids, gold first tokens, negative-sampling
  cumulative distribution) produced by this repository's preprocessor + DatasetManager.  No text of the corpus is stored.
* ``readme_run.json``: the float64 ORACLE's end-to-end trace of that run through the product's own driver
  (``ReconstructInducer`` with the oracle bound as backend): per-batch costs of epoch 1, per-epoch training error, per-epoch
  B-cubed (F1, P, R) on valid/test, final cluster sizes.  Theano is not installable, so this is the oracle's trace, not the
  reference's (parity unpinned).
"""
import io
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from relation_autoencoder_b200 import data as D, induction as I, preprocess as P  # noqa: E402
from tests.oracle_backend import oracle_backend  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/data-sample.txt"
README_ARGS = dict(nb_epochs=10, learning_rate=0.1, batch_size=100, embed_size=10, nb_relations=10, nb_neg_samples=5,
                   lambda1=0.0, lambda2=0.1, optimization='adagrad', model_name='discrete-autoencoder',
                   decoder_model='rescal+sp', external_embeddings=False, extended_regularizer=True, frequent_eval=False,
                   alpha=0.1)
SEED = 2


def run_readme(indexed, backend, epochs=None, **override):
    rng = np.random.RandomState(seed=SEED)
    args = dict(README_ARGS)
    args.update(override)
    if epochs is not None:
        args['nb_epochs'] = epochs
    ind = I.ReconstructInducer(indexed, indexed.goldStandard, rng, backend=backend, out=io.StringIO(), **args)
    ind.compile_function()
    batch_costs = []
    train = ind.func['train']

    def traced(b, n1, n2):
        c = train(b, n1, n2)
        batch_costs.append(float(c))
        return c
    ind.func['train'] = traced
    ind.learn(debug=False)
    sizes = ind.get_clusters_size(ind.func['label_train'], ind.batch_reps['train'])
    return dict(batch_costs=batch_costs, train_error=ind.train_error_series,
                metrics={s: [list(m) for m in v] for s, v in ind.metrics.items()},
                cluster_sizes={int(k): int(v) for k, v in sorted(sizes.items())}), ind


def main():
    with tempfile.TemporaryDirectory() as tmp:
        pk = os.path.join(tmp, "sample.pk")
        for batch in ("train", "dev", "test"):
            P.preprocess(SRC, pk, batch=batch, verbose=False)
        dm, gold = D.load_data(pk, np.random.RandomState(SEED))
    indexed = D.IndexedDataset.from_manager(dm, gold)
    indexed.save_npz(os.path.join(HERE, "data_sample_indexed.npz"))
    indexed = D.IndexedDataset.load_npz(os.path.join(HERE, "data_sample_indexed.npz"))
    # (1) the README flags verbatim; (2) the same run without the L2 term - with l2 = 0.1 the regulariser drives W to
    # zero on this 1000-sentence sample and every example lands in ONE cluster, which checks little about the clustering
    for name, override in (("readme_run.json", {}), ("readme_run_l2_0.json", {"lambda2": 0.0})):
        res, _ = run_readme(indexed, oracle_backend, **override)
        cfg = dict(README_ARGS)
        cfg.update(override)
        res["config"] = cfg
        res["seed"] = SEED
        res["F"] = indexed.get_dimensionality()
        res["N"] = indexed.get_arg_voc_size()
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(res, f, indent=1)
        print(name, "F=%d N=%d" % (res["F"], res["N"]))
        print("  train error per epoch:", ["%.4f" % e for e in res["train_error"]])
        print("  test (f1, pre, rec) last epoch:", ["%.4f" % x for x in res["metrics"]["test"][-1]])
        print("  cluster sizes:", res["cluster_sizes"])


if __name__ == "__main__":
    main()
