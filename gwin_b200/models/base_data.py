"""Models that need data and a waveform generator.

Public surface of the reference's ``BaseDataModel`` (``gwin/models/base_data.py:34-161``):
the model owns a copy of the data, exposes the generator, adds the cached stats ``lognl`` and
``loglr`` plus the prior-weighted likelihood ratio ``logplr``, and applies waveform transforms
after the sampling transforms / boundary conditions.
"""
from abc import abstractmethod

import numpy

from .. import transforms as _transforms
from .base import BaseModel


def _cached(statname):
    """Property that returns ``current_stats.<statname>``, computing it once per ``update`` with
    the subclass hook ``_<statname>`` (what the reference spells out per stat with
    ``_trytoget``)."""
    def getter(self):
        return self._trytoget(statname, getattr(self, '_' + statname))
    getter.__doc__ = "The %s at the current parameters (cached until the next update)." % statname
    return property(getter)


class BaseDataModel(BaseModel):

    def __init__(self, variable_params, data, waveform_generator,
                 waveform_transforms=None, **kwargs):
        # private copy: the caller's series are never modified (reference :87)
        self._data = dict((ifo, series.copy()) for ifo, series in data.items())
        self._waveform_generator = waveform_generator
        self._waveform_transforms = waveform_transforms
        BaseModel.__init__(self, variable_params, **kwargs)

    lognl = _cached('lognl')
    loglr = _cached('loglr')

    @abstractmethod
    def _lognl(self):
        """log likelihood of the data being noise only."""

    @abstractmethod
    def _loglr(self):
        """log likelihood ratio at the current parameters."""

    @property
    def _extra_stats(self):
        return ['loglr', 'lognl']

    @property
    def logplr(self):
        """``logprior + loglr``; a point excluded by the prior never reaches the likelihood
        (reference :128-141)."""
        logp = self.logprior
        return logp if logp == -numpy.inf else logp + self.loglr

    data = property(lambda self: self._data, doc="The (whitened) copy of the data.")
    waveform_generator = property(lambda self: self._waveform_generator,
                                  doc="The waveform generator handed to the constructor.")

    def _transform_params(self, **params):
        params = BaseModel._transform_params(self, **params)
        if self._waveform_transforms is None:
            return params
        return _transforms.apply_transforms(params, self._waveform_transforms, inverse=False)
