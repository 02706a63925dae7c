"""``TestNormal``: the analytic multivariate-normal model of the reference
(``gwin/models/analytic.py:27-85``), re-created as a sampler self-test fixture.
It has no waveform and stays on the host."""
from collections import OrderedDict

import numpy
from scipy import stats

from .base import BaseModel


class TestNormal(BaseModel):
    __test__ = False   # not a pytest class
    name = "test_normal"

    def __init__(self, variable_params, mean=None, cov=None, **kwargs):
        super(TestNormal, self).__init__(variable_params, **kwargs)
        n = len(self.variable_params)
        if mean is None:
            mean = [0.] * n
        if cov is None:
            cov = [1.] * n
        self._dist = stats.multivariate_normal(mean=mean, cov=cov)
        if self._dist.dim != n:
            raise ValueError("dimension mis-match between variable_params "
                             "and mean and/or cov")

    def _loglikelihood(self):
        return float(self._dist.logpdf([self.current_params[p]
                                        for p in self.variable_params]))

    def _batch_loglikelihood(self, params, W, skip):
        x = numpy.stack([numpy.broadcast_to(params[p], (W,))
                         for p in self.variable_params], axis=-1)
        return numpy.atleast_1d(self._dist.logpdf(x)), OrderedDict()
