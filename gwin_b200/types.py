"""Minimal frequency-series container exposing exactly the attributes the
reference's hot-path files touch on ``pycbc.types.FrequencySeries``
(SURVEY.md appendix C): ``len``, ``delta_f``, ``epoch``/``start_time``,
``copy()``, slicing, ``numpy()``/``data`` and ``inner``.

``inner`` here exists for API completeness on *host* data (it is what
``lognl`` docstrings and users call); the likelihood itself never uses it --
the inner products of the hot path are computed by the CUDA kernel.
"""
import numpy


class FrequencySeries(object):
    def __init__(self, initial_array, delta_f, epoch=0., dtype=None, copy=True):
        arr = numpy.array(initial_array, dtype=dtype, copy=copy)
        if arr.ndim != 1:
            raise ValueError("FrequencySeries must be one-dimensional")
        self._data = arr
        self._delta_f = float(delta_f)
        self._epoch = float(epoch)

    # -- metadata ---------------------------------------------------------
    @property
    def delta_f(self):
        return self._delta_f

    @property
    def epoch(self):
        return self._epoch

    start_time = epoch

    @property
    def delta_t(self):
        return 1.0 / ((len(self) - 1) * 2 * self._delta_f)

    @property
    def sample_frequencies(self):
        return numpy.arange(len(self)) * self._delta_f

    @property
    def dtype(self):
        return self._data.dtype

    # -- data -------------------------------------------------------------
    @property
    def data(self):
        return self._data

    def numpy(self):
        return self._data

    def __array__(self, dtype=None, copy=None):
        return numpy.asarray(self._data, dtype=dtype)

    def __len__(self):
        return len(self._data)

    def __getitem__(self, idx):
        out = self._data[idx]
        if isinstance(idx, slice):
            return out
        return out

    def __setitem__(self, idx, val):
        self._data[idx] = val

    def copy(self):
        return FrequencySeries(self._data, self._delta_f, self._epoch, copy=True)

    def inner(self, other):
        """sum(conj(self) * other) -- pycbc ``Array.inner``."""
        return numpy.vdot(self._data, numpy.asarray(other))

    def __repr__(self):
        return "FrequencySeries(n=%d, delta_f=%r, epoch=%r)" % (
            len(self), self._delta_f, self._epoch)
