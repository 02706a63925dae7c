"""Likelihood models (the namespace mirrors ``gwin.models``) and the
sampler-facing ``CallModel`` adapter (reference ``gwin/models/__init__.py``:
``CallModel`` :37-137, ``_call_global_model`` :28-34, ``models`` :164-171).

``CallModel`` keeps the scalar call ``call_model(param_values) -> (val, blob)``
that emcee / kombine issue through ``pool.map`` and adds ``batch(points)``, which
``gwin_b200.pool.BatchPool.map`` uses to turn the whole map into one kernel
launch.
"""
from .analytic import TestNormal
from .base import BaseModel, BatchResult, ModelStats, SamplingTransforms
from .base_data import BaseDataModel
from .gaussian_noise import GaussianNoise
from .marginalized_gaussian_noise import (MarginalizedGaussianNoise,
                                          MarginalizedPhaseGaussianNoise)

# Used to manage a model instance across processes (reference :28-34).
_global_instance = None


def _call_global_model(*args, **kwds):
    return _global_instance(*args, **kwds)  # pylint:disable=not-callable


class CallModel(object):
    """Wraps a model so that it can be called like a function of the
    parameter values, in ``model.sampling_params`` order."""

    def __init__(self, model, callstat, return_all_stats=True):
        self.model = model
        self.callstat = callstat
        self.return_all_stats = return_all_stats

    def __getattr__(self, attr):
        # only reached for attributes CallModel does not define itself
        if attr == 'model':
            raise AttributeError(attr)
        return getattr(self.model, attr)

    def __call__(self, param_values, callstat=None, return_all_stats=None):
        if callstat is None:
            callstat = self.callstat
        if return_all_stats is None:
            return_all_stats = self.return_all_stats
        params = dict(zip(self.model.sampling_params, param_values))
        self.model.update(**params)
        val = getattr(self.model, callstat)
        if return_all_stats:
            return val, self.model.get_current_stats()
        return val

    def batch(self, points, callstat=None):
        """All points at once: returns a ``BatchResult``."""
        if callstat is None:
            callstat = self.callstat
        return self.model.evaluate_batch(points, callstat=callstat)


models = {_cls.name: _cls for _cls in (
    TestNormal,
    GaussianNoise,
    MarginalizedGaussianNoise,
    MarginalizedPhaseGaussianNoise,
)}
