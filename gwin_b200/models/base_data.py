"""Models that need data and a waveform generator -- same interface as the
reference ``BaseDataModel`` (``gwin/models/base_data.py:34-161``): ``data``,
``waveform_generator``, ``lognl``, ``loglr``, ``logplr``, waveform transforms
applied after the sampling transforms."""
from abc import abstractmethod

import numpy

from .. import transforms as _transforms
from .base import BaseModel


class BaseDataModel(BaseModel):

    def __init__(self, variable_params, data, waveform_generator,
                 waveform_transforms=None, **kwargs):
        # the model keeps its own copy of the data (reference base_data.py:87)
        self._data = {ifo: d.copy() for (ifo, d) in data.items()}
        self._waveform_generator = waveform_generator
        self._waveform_transforms = waveform_transforms
        super(BaseDataModel, self).__init__(variable_params, **kwargs)

    @property
    def _extra_stats(self):
        return ['loglr', 'lognl']

    @property
    def lognl(self):
        return self._trytoget('lognl', self._lognl)

    @abstractmethod
    def _lognl(self):
        pass

    @property
    def loglr(self):
        return self._trytoget('loglr', self._loglr)

    @abstractmethod
    def _loglr(self):
        pass

    @property
    def logplr(self):
        """logprior + loglr, skipping loglr when the prior is -inf
        (reference base_data.py:128-141)."""
        logp = self.logprior
        if logp == -numpy.inf:
            return logp
        return logp + self.loglr

    @property
    def waveform_generator(self):
        return self._waveform_generator

    @property
    def data(self):
        return self._data

    def _transform_params(self, **params):
        params = super(BaseDataModel, self)._transform_params(**params)
        if self._waveform_transforms is not None:
            params = _transforms.apply_transforms(
                params, self._waveform_transforms, inverse=False)
        return params
