"""The analytic models of the reference (``gwin/models/analytic.py``: ``TestNormal`` :27-85,
``TestEggbox`` :88-117, ``TestRosenbrock`` :120-152, ``TestVolcano`` :155-193), re-created as sampler
self-test fixtures.  They have no waveform and stay on the host; each
evaluates a whole batch of points with one numpy expression (``_batch_loglikelihood``)."""
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


class _PointFunction(BaseModel):
    """A model whose log likelihood is a closed-form function of the point: subclasses give
    ``_logl(x)`` for ``x`` of shape [W, ndim] (columns in ``variable_params`` order)."""

    def _loglikelihood(self):
        x = numpy.array([[self.current_params[p] for p in self.variable_params]], dtype=float)
        return float(self._logl(x)[0])

    def _batch_loglikelihood(self, params, W, skip):
        x = numpy.stack([numpy.broadcast_to(numpy.asarray(params[p], dtype=float), (W,))
                         for p in self.variable_params], axis=-1)
        return numpy.atleast_1d(self._logl(x)), OrderedDict()


class TestEggbox(_PointFunction):
    """log L = (2 + prod_i cos(theta_i / 2))^5, any number of dimensions."""
    __test__ = False
    name = "test_eggbox"

    @staticmethod
    def _logl(x):
        return (2.0 + numpy.cos(0.5 * x).prod(axis=-1)) ** 5


class TestRosenbrock(_PointFunction):
    """log L = -sum_i [(1 - theta_i)^2 + 100 (theta_{i+1} - theta_i^2)^2]."""
    __test__ = False
    name = "test_rosenbrock"

    @staticmethod
    def _logl(x):
        a, b = x[..., :-1], x[..., 1:]
        return -((1.0 - a) ** 2 + 100.0 * (b - a * a) ** 2).sum(axis=-1)


class TestVolcano(_PointFunction):
    """Two dimensions, r = |theta|: log L = 25 (exp(-r / 35) + N(r; 5, 2))."""
    __test__ = False
    name = "test_volcano"

    def __init__(self, variable_params, **kwargs):
        super(TestVolcano, self).__init__(variable_params, **kwargs)
        if len(self.variable_params) != 2:
            raise ValueError("TestVolcano distribution requires exactly "
                             "two variable args")

    @staticmethod
    def _logl(x):
        r = numpy.hypot(x[..., 0], x[..., 1])
        return 25.0 * (numpy.exp(-r / 35.0) + stats.norm.pdf(r, loc=5.0, scale=2.0))
