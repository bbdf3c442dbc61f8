"""Host-side negative sampler: same class name, constructor and call as the reference's
``learning/NegativeExampleGenerator.py:4-32`` so the driver's two calls per epoch (``OieInduction.py:183-184``) are
unchanged.

The ids are an integer function of a float64 uniform stream and MUST be bit-exact: the draw stays on the
host with the caller's legacy ``numpy.random.RandomState`` (frozen stream), and the inverse-CDF lookup is one vectorised
``searchsorted`` (side='left'), element-wise identical to the reference's per-element ``map`` (NegativeExampleGenerator.py:32).
The resulting [S, n] int32 arrays are what ``Engine.bind_epoch_negatives`` / ``Engine.train`` consume.
"""
from __future__ import annotations

import numpy as np


class NegativeExampleGenerator(object):
    def __init__(self, rand, neg_sampling_cum):
        """
        :param rand: numpy.random.RandomState (the run's single generator, OieInduction.py:262,296)
        :param neg_sampling_cum: ascending cumulative distribution, ``neg_sampling_cum[-1] == 1`` (OieData.py:57-59)
        """
        self._rand = rand
        self._negSamplingCum = np.asarray(neg_sampling_cum, dtype=np.float64)
        assert abs(self._negSamplingCum[-1] - 1) < 1.e-4, \
            'Negative example generator initialized with a cumulative distribution derived from a non-normalized one'

    def get_negative_samples(self, num_positive_entities, num_negative_samples):
        """(s, l) int32 array: column j holds the ``s`` sampled entity ids for example j (NegativeExampleGenerator.py:14-24)."""
        return self._get_sample(num_positive_entities * num_negative_samples).reshape(
            (num_negative_samples, num_positive_entities))

    def _get_sample(self, num_samples):
        u = self._rand.uniform(0, self._negSamplingCum[-1], num_samples)
        return np.asarray(self._negSamplingCum.searchsorted(u), dtype=np.int32)


def neg_sampling_cum(powered_frequencies):
    """Cumulative distribution of already-powered entity frequencies, summed left to right like the reference's python
    ``sum`` over a generator (OieData.py:57) and ``np.cumsum`` (:59)."""
    f = np.asarray(powered_frequencies, dtype=np.float64)
    norm1 = float(np.cumsum(f)[-1]) if f.size else 1.0
    return np.cumsum(f / norm1)
