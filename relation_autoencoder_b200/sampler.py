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


class DeviceNegativeExampleGenerator(object):
    """Same constructor and call as :class:`NegativeExampleGenerator`, but the draw and the inverse-CDF lookup run on the
    GPU (``rae_sample_negatives``) and the result stays there: an int32 CUDA tensor [s, l] ready for
    ``Engine.bind_epoch_negatives``.  The ids are BIT-IDENTICAL to the host generator's: the kernel advances the very same
    MT19937 stream (the state of ``rand`` is uploaded before and written back after every call, so host and device draws
    can be interleaved) and compares in float64.  At NYT scale one epoch needs 2 x 20 x 2M draws: ~4 s of host
    ``searchsorted`` versus a few milliseconds here."""

    def __init__(self, rand, neg_sampling_cum, device=0):
        import torch
        from . import _lib as L
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: use NegativeExampleGenerator (host) instead")
        self._rand = rand
        cum = np.ascontiguousarray(neg_sampling_cum, dtype=np.float64)
        assert abs(cum[-1] - 1) < 1.e-4, \
            'Negative example generator initialized with a cumulative distribution derived from a non-normalized one'
        self._torch = torch
        self._lib = L.load()
        self._dev = torch.device("cuda", device)
        self._cum_last = float(cum[-1])
        self._cum = torch.from_numpy(cum).to(self._dev)

    def get_negative_samples(self, num_positive_entities, num_negative_samples):
        import ctypes as C
        torch = self._torch
        n = int(num_positive_entities) * int(num_negative_samples)
        kind, key, pos, has_gauss, cached = self._rand.get_state()
        if kind != 'MT19937':
            raise ValueError("the legacy RandomState (MT19937) is required for bit-exact ids")
        state = np.empty(625, dtype=np.uint32)
        state[:624] = key
        state[624] = pos
        st_dev = torch.from_numpy(state.view(np.int32)).to(self._dev)
        words = torch.empty(max(2 * n, 1), dtype=torch.int32, device=self._dev)
        out = torch.empty(max(n, 1), dtype=torch.int32, device=self._dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(self._dev).cuda_stream)
        rc = self._lib.rae_sample_negatives(None, p(st_dev), p(self._cum), self._cum.numel(), self._cum_last, p(out), n, p(words), stream)
        if rc != 0:
            raise RuntimeError("rae_sample_negatives failed (%d): %s" % (rc, self._lib.rae_last_error(None).decode()))
        new_state = st_dev.cpu().numpy().view(np.uint32)
        self._rand.set_state((kind, new_state[:624].copy(), int(new_state[624]), has_gauss, cached))
        return out[:n].reshape(int(num_negative_samples), int(num_positive_entities))
