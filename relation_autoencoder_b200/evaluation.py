"""B-cubed clustering metrics of the reference's evaluator, same class / method names and call protocol
(``evaluation/OieEvaluation.py:5-209``): ``construct_split_evaluator(gold, split)`` -> object with
``feed_induced_clusters(response)`` and ``compute_metrics() -> (F1, precision, recall)``.

Definitions kept exactly (OieEvaluation.py:66-96,121-127):
  * gold clusters come from the FIRST label token of each example; '' = unlabelled, not assessable (:185-209);
  * per assessable member m of induced cluster c with gold cluster g(m):
        precision(m) = |c & g| / |c & assessable|        (:66-76, the denominator ignores unlabelled members)
        recall(m)    = |c & g| / |g|                      (:78-88)
  * totals are averaged over the assessable members (:90-96,121-127); F1 is the harmonic mean, 0 when both are 0.

The reference walks every gold cluster for every member (``_find_cluster``, O(n * #gold)); here each induced cluster is
reduced once to its (gold label -> count) histogram, so a split is evaluated in O(n) with the same result (the sums
are over integers-ratios grouped per (cluster, gold) pair; agreement with the per-member loop is tested to 1e-12).
"""
from __future__ import annotations

from collections import Counter
from typing import Dict, Iterable, List, Set, Tuple

SPLIT_LABELS = ['train', 'valid', 'test']           # settings.py:26


class SingleLabelClusterEvaluation(object):
    def __init__(self, split_goldstandard: Dict[int, List[str]], split_label: str):
        assert split_label in SPLIT_LABELS
        self.split_label = split_label
        self.numberOfElements = 0
        self.induced_clusters: Dict[int, Set[int]] = {}
        self.gold_clusters, self.assessableElemSet = self._parse_first_relation_label(split_goldstandard)
        self._gold_of: Dict[int, str] = {}
        for label, members in self.gold_clusters.items():
            for m in members:
                self._gold_of[m] = label

    # ------------------------------------------------------------------ protocol used by the driver
    def feed_induced_clusters(self, response: Dict[int, Iterable[int]]):
        """Copies the clustering, dropping empty clusters (OieEvaluation.py:23-34)."""
        self.numberOfElements = 0
        self.induced_clusters = {}
        for cluster_id, example_set in response.items():
            if len(example_set) > 0:
                self.numberOfElements += len(example_set)
                self.induced_clusters[cluster_id] = set(example_set)

    def compute_metrics(self) -> Tuple[float, float, float]:
        """(F1, precision, recall), B-cubed per element (OieEvaluation.py:36-44)."""
        rec = self.b3_total_element_recall()
        pre = self.b3_total_element_precision()
        if rec == 0.0 and pre == 0.0:
            return 0.0, pre, rec
        return (2 * rec * pre) / (rec + pre), pre, rec

    def get_f1(self) -> float:
        return self.compute_metrics()[0]

    def get_f_n(self, n) -> float:
        """F-beta with beta = n (OieEvaluation.py:54-64)."""
        if n == 1:
            return self.get_f1()
        rec = self.b3_total_element_recall()
        pre = self.b3_total_element_precision()
        if rec == 0.0 and pre == 0.0:
            return 0.0
        b2 = n ** 2
        return ((1 + b2) * rec * pre) / ((b2 * pre) + rec)

    # ------------------------------------------------------------------ B-cubed
    def _histograms(self):
        for members in self.induced_clusters.values():
            hist = Counter(self._gold_of[m] for m in members if m in self._gold_of)
            yield hist, sum(hist.values())

    def b3_total_element_precision(self) -> float:
        total = 0.0
        for hist, assessable in self._histograms():
            for label, n_cg in hist.items():
                total += n_cg * (n_cg / float(assessable))
        return total / float(len(self.assessableElemSet))

    def b3_total_element_recall(self) -> float:
        total = 0.0
        for hist, _ in self._histograms():
            for label, n_cg in hist.items():
                total += n_cg * (n_cg / float(len(self.gold_clusters[label])))
        return total / len(self.assessableElemSet)

    def b3_total_cluster_precision(self) -> float:
        """Cluster-weighted variant (OieEvaluation.py:98-105)."""
        total = 0.0
        n_clusters = len(self.induced_clusters)
        for members, (hist, assessable) in zip(self.induced_clusters.values(), self._histograms()):
            for label, n_cg in hist.items():
                total += n_cg * (n_cg / float(assessable)) / (n_clusters * len(members))
        return total

    def b3_total_cluster_recall(self) -> float:
        """Cluster-weighted variant (OieEvaluation.py:129-135)."""
        total = 0.0
        n_clusters = len(self.induced_clusters)
        for members, (hist, _) in zip(self.induced_clusters.values(), self._histograms()):
            for label, n_cg in hist.items():
                total += n_cg * (n_cg / float(len(self.gold_clusters[label]))) / (n_clusters * len(members))
        return total

    # ------------------------------------------------------------------ single-pair helpers (same names as the reference)
    def precision(self, retrieved_members: Set[int], true_members: Set[int]) -> float:
        return len(retrieved_members & true_members) / float(len(retrieved_members & self.assessableElemSet))

    def b3recall(self, retrieved_members: Set[int], true_members: Set[int]) -> float:
        return len(retrieved_members & true_members) / float(len(true_members))

    @staticmethod
    def _parse_first_relation_label(relations: Dict[int, List[str]]):
        """gold label -> set(example ids), and the set of labelled ids (OieEvaluation.py:185-209)."""
        gold_relation2ids: Dict[str, Set[int]] = {}
        labeled = set()
        for example_id, label_list in relations.items():
            first = label_list[0]
            if first != '':
                labeled.add(example_id)
                gold_relation2ids.setdefault(first, set()).add(example_id)
        return gold_relation2ids, labeled


def construct_split_evaluator(split_goldstandard, split_label):
    """OieEvaluation.py:227-238."""
    assert split_label in SPLIT_LABELS
    return SingleLabelClusterEvaluation(split_goldstandard, split_label)


def construct_split_evaluator_from_file(goldstandard_file, split_label):
    """Gold labels straight from a tab-separated file: last field, split on blanks (OieEvaluation.py:241-256, subset=None)."""
    relations = {}
    with open(goldstandard_file, 'r') as ref_set:
        for example_id, line in enumerate(ref_set):
            relations[example_id] = line.split('\t')[-1].strip().split(' ')
    return SingleLabelClusterEvaluation(relations, split_label)


def get_clusters_sets(labeling_func, nb_batches, nb_relations):
    """cluster id -> set of example indices from ``labeling_func(i)[0]`` over all batches (OieInduction.py:321-340)."""
    clusters = {i: set() for i in range(nb_relations)}
    idx = 0
    for i in range(nb_batches):
        for pred in labeling_func(i)[0]:
            clusters[int(pred)].add(idx)
            idx += 1
    return clusters
