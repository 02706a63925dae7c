"""Likelihood models (the namespace mirrors ``gwin.models``) and the
sampler-facing ``CallModel`` adapter (reference ``gwin/models/__init__.py``:
``CallModel`` :37-137, ``_call_global_model`` :28-34, ``models`` :164-171).

``CallModel`` keeps the scalar call ``call_model(param_values) -> (val, blob)``
that emcee / kombine issue through ``pool.map`` and adds ``batch(points)``, which
``gwin_b200.pool.BatchPool.map`` uses to turn the whole map into one kernel
launch.
"""
from .analytic import TestEggbox, TestNormal, TestRosenbrock, TestVolcano
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
    """Callable view of a model for samplers: ``call_model(values)`` updates the model with
    ``values`` (ordered like ``model.sampling_params``) and returns ``callstat`` -- together with
    the tuple of the model's default stats when ``return_all_stats`` is set.  Every other
    attribute is looked up on the wrapped model.  ``batch(points)`` is the vectorised form."""

    def __init__(self, model, callstat, return_all_stats=True):
        self.model = model
        self.callstat = callstat
        self.return_all_stats = return_all_stats

    def __getattr__(self, attr):
        if attr == 'model':                 # not yet set (e.g. during unpickling)
            raise AttributeError(attr)
        return getattr(self.model, attr)

    def __call__(self, param_values, callstat=None, return_all_stats=None):
        stat = self.callstat if callstat is None else callstat
        with_stats = self.return_all_stats if return_all_stats is None else return_all_stats
        model = self.model
        model.update(**dict(zip(model.sampling_params, param_values)))
        value = getattr(model, stat)
        return (value, model.get_current_stats()) if with_stats else value

    def batch(self, points, callstat=None):
        """All points at once: returns a ``BatchResult``."""
        return self.model.evaluate_batch(points, callstat=callstat or self.callstat,
                                         with_stats=self.return_all_stats)

    def batch_async(self, points, callstat=None):
        """``batch`` that only queues the evaluation: returns a handle whose ``result()`` is the
        ``BatchResult``.  Up to two may be pending per model (``gwin_loglr_batch_points_submit``)."""
        return self.model.evaluate_batch(points, callstat=callstat or self.callstat,
                                         with_stats=self.return_all_stats, defer=True)


models = {_cls.name: _cls for _cls in (
    TestEggbox,
    TestNormal,
    TestRosenbrock,
    TestVolcano,
    GaussianNoise,
    MarginalizedGaussianNoise,
    MarginalizedPhaseGaussianNoise,
)}
