"""Model base class: same public surface as ``gwin/models/base.py`` (reference
``BaseModel`` :354-621, ``ModelStats`` :58-106, ``SamplingTransforms`` :109-220,
``_NoPrior`` :44-55) plus a *batched* evaluation entry point.

The reference evaluates one walker per call (``update(**params)`` followed by a
lazily cached stat property).  That scalar protocol is kept so the class drops
in under ``CallModel``; in addition ``evaluate_batch(points)`` runs the same
pipeline -- inverse sampling transforms, prior boundary conditions, waveform
transforms, log-Jacobian, log-prior, prior short-circuit, likelihood -- for W
walkers at once with array arithmetic, so that the likelihood sees the whole
ensemble half-step in one kernel launch.
"""
from abc import ABCMeta, abstractmethod
from collections import OrderedDict

import numpy

from .. import transforms as _transforms
from ..distributions import FieldArray


class _NoPrior(object):
    """Dummy prior returning 0 for the log prior (reference base.py:44-55)."""

    variable_args = ()

    @staticmethod
    def apply_boundary_conditions(**params):
        return params

    def __call__(self, **params):
        return 0.


class ModelStats(object):
    """Holds the stats computed for the current parameters."""

    @property
    def statnames(self):
        return list(self.__dict__.keys())

    def getstats(self, names, default=numpy.nan):
        return tuple(getattr(self, n, default) for n in names)

    def getstatsdict(self, names, default=numpy.nan):
        return OrderedDict(zip(names, self.getstats(names, default=default)))


class SamplingTransforms(object):
    """Maps between the variable parameters and the parameters the sampler
    walks in, with the log-Jacobian (reference base.py:109-220).

    ``replace_parameters`` are the variable params that are replaced;
    ``sampling_params`` what they are replaced with.
    """

    def __init__(self, variable_params, sampling_params, replace_parameters,
                 sampling_transforms):
        assert len(replace_parameters) == len(sampling_params), (
            "number of sampling parameters must be the same as the number of "
            "replace parameters")
        self.sampling_transforms = list(sampling_transforms)
        self.sampling_params = tuple(
            [p for p in variable_params if p not in replace_parameters] +
            list(sampling_params))
        self.variable_params = tuple(variable_params)

    def logjacobian(self, **params):
        """log |d(variable)/d(sampling)| at ``params`` (which must hold the
        variable params; sampling params are derived if absent)."""
        j = _transforms.compute_jacobian(params, self.sampling_transforms,
                                         inverse=True)
        with numpy.errstate(divide="ignore"):
            return numpy.log(numpy.abs(j))

    def apply(self, samples, inverse=False):
        return _transforms.apply_transforms(samples, self.sampling_transforms,
                                            inverse=inverse)


class BatchResult(object):
    """Result of ``evaluate_batch``: ``values`` [W] of the call stat and
    ``stats`` (OrderedDict name -> array [W]) for the model's default stats.
    ``rows()`` yields what W scalar ``CallModel`` calls would have returned."""

    def __init__(self, values, stats, names=None):
        self.values = values
        self._stats = stats          # dict, or a callable building it on first use
        self._names = names

    @property
    def stats(self):
        if callable(self._stats):
            self._stats = self._stats()
        if self._names is not None:
            W = len(self.values)
            full = OrderedDict()
            for n in self._names:
                v = self._stats.get(n)
                full[n] = numpy.full(W, numpy.nan) if v is None else v
            self._stats, self._names = full, None
        return self._stats

    def __len__(self):
        return len(self.values)

    def rows(self, return_all_stats=True):
        vals = self.values.tolist()
        if not return_all_stats:
            return vals
        cols = []
        for v in self.stats.values():
            v = numpy.asarray(v)
            cols.append(v.tolist())
        blobs = list(zip(*cols)) if cols else [()] * len(vals)
        return list(zip(vals, blobs))


class Deferred(object):
    """What ``_batch_callstat`` returns for a queued evaluation: ``finish()`` waits for it and
    returns the ``(values, stats)`` pair the synchronous form returns."""

    def __init__(self, finish):
        self.finish = finish


class PendingBatch(object):
    """Handle of ``evaluate_batch(..., defer=True)`` / ``CallModel.batch_async``."""

    def __init__(self, finish):
        self._finish, self._res = finish, None

    def result(self):
        if self._finish is not None:
            self._res, self._finish = self._finish(), None
        return self._res


class BaseModel(object, metaclass=ABCMeta):
    name = None

    def __init__(self, variable_params, static_params=None, prior=None,
                 sampling_transforms=None):
        if isinstance(variable_params, str):
            variable_params = (variable_params,)
        if not isinstance(variable_params, tuple):
            variable_params = tuple(variable_params)
        self._variable_params = variable_params
        if static_params is None:
            static_params = {}
        self._static_params = static_params
        if prior is None:
            self.prior_distribution = _NoPrior()
        else:
            assert tuple(prior.variable_args) == variable_params, (
                "variable params of prior and model must be the same")
            self.prior_distribution = prior
        self.sampling_transforms = sampling_transforms
        self._current_params = None
        self._current_stats = ModelStats()

    # -- parameters -------------------------------------------------------
    @property
    def variable_params(self):
        return self._variable_params

    @property
    def static_params(self):
        return self._static_params

    @property
    def sampling_params(self):
        if self.sampling_transforms is None:
            return self.variable_params
        return self.sampling_transforms.sampling_params

    def update(self, **params):
        """Sets the current parameters (after transforms) and resets stats."""
        self._current_params = self._transform_params(**params)
        self._current_stats = ModelStats()

    @property
    def current_params(self):
        if self._current_params is None:
            raise ValueError("no parameters values currently stored; "
                             "run update to add some")
        return self._current_params

    # -- stats ------------------------------------------------------------
    @property
    def default_stats(self):
        return ['logjacobian', 'logprior', 'loglikelihood'] + self._extra_stats

    @property
    def _extra_stats(self):
        return []

    def get_current_stats(self, names=None):
        if names is None:
            names = self.default_stats
        return self._current_stats.getstats(names)

    @property
    def current_stats(self):
        return self._current_stats.getstatsdict(self.default_stats)

    def _trytoget(self, statname, fallback, **kwargs):
        try:
            return getattr(self._current_stats, statname)
        except AttributeError:
            val = fallback(**kwargs)
            setattr(self._current_stats, statname, val)
            return val

    @property
    def loglikelihood(self):
        return self._trytoget('loglikelihood', self._loglikelihood)

    @abstractmethod
    def _loglikelihood(self):
        pass

    @property
    def logjacobian(self):
        return self._trytoget('logjacobian', self._logjacobian)

    def _logjacobian(self):
        if self.sampling_transforms is None:
            return 0.
        return float(self.sampling_transforms.logjacobian(**self.current_params))

    @property
    def logprior(self):
        return self._trytoget('logprior', self._logprior)

    def _logprior(self):
        logj = self.logjacobian
        logp = float(self.prior_distribution(**self.current_params)) + logj
        if numpy.isnan(logp):
            logp = -numpy.inf
        return logp

    @property
    def logposterior(self):
        """logprior first; if it is -inf the likelihood is not evaluated
        (reference base.py:556-568)."""
        logp = self.logprior
        if logp == -numpy.inf:
            return logp
        return logp + self.loglikelihood

    def prior_rvs(self, size=1, prior=None):
        if prior is None:
            prior = self.prior_distribution
        p0 = prior.rvs(size=size)
        if self.sampling_transforms is not None:
            ptrans = self.sampling_transforms.apply(
                {n: p0[n] for n in p0.dtype.names})
            p0 = FieldArray.from_arrays(
                [ptrans[arg] for arg in self.sampling_params],
                names=self.sampling_params)
        return p0

    def _transform_params(self, **params):
        if self.sampling_transforms is not None:
            params = self.sampling_transforms.apply(params, inverse=True)
        params = self.prior_distribution.apply_boundary_conditions(**params)
        return params

    # -- batched evaluation ------------------------------------------------
    def _batch_logprior(self, params, W):
        """Vectorised ``_logjacobian`` + ``_logprior`` for W points."""
        if self.sampling_transforms is None:
            logj = numpy.zeros(W)
        else:
            logj = numpy.broadcast_to(numpy.asarray(
                self.sampling_transforms.logjacobian(**params), dtype=float),
                (W,)).copy()
        try:
            lp = self.prior_distribution(**params)
            lp = numpy.broadcast_to(numpy.asarray(lp, dtype=float), (W,))
        except (TypeError, ValueError):
            # a user-supplied scalar-only prior: fall back to one call per point
            lp = numpy.array([
                self.prior_distribution(**{k: (v[i] if numpy.ndim(v) else v)
                                           for k, v in params.items()})
                for i in range(W)], dtype=float)
        logp = lp + logj
        logp = numpy.where(numpy.isnan(logp), -numpy.inf, logp)
        return logj, logp

    def _batch_loglikelihood(self, params, W, skip):
        """Returns (loglikelihood [W], extra stats dict).  Points with
        ``skip`` set need not be evaluated.  Default: loop the scalar path."""
        out = numpy.full(W, numpy.nan)
        for i in range(W):
            if skip is not None and skip[i]:
                continue
            self._current_params = {k: (v[i] if numpy.ndim(v) else v)
                                    for k, v in params.items()}
            self._current_stats = ModelStats()
            out[i] = self.loglikelihood
        return out, OrderedDict()

    def evaluate_batch(self, points, callstat='logposterior', with_stats=True, defer=False):
        """Evaluate ``callstat`` for W points given in ``sampling_params`` order.

        Returns a ``BatchResult`` whose ``stats`` are materialised on first use.
        Equivalent to W calls of ``CallModel.__call__`` (reference
        ``models/__init__.py:101-137``) including the ``logprior == -inf``
        short-circuit of ``logposterior``/``logplr``; stats that the scalar
        path would not have computed for a point are NaN for that point.
        ``with_stats=False`` tells the model that only ``values`` will be read (the
        per-detector statistics are then not brought back from the device; asking
        for them anyway evaluates the batch again).
        ``defer=True`` returns a ``PendingBatch`` at once -- a model with a device behind it has only
        queued the work (copies in, kernels, copies out) -- whose ``result()`` is the ``BatchResult``;
        a caller with several batches keeps two in flight so that one's copies overlap the other's kernels.
        """
        points = numpy.asarray(points, dtype=numpy.float64)
        if points.ndim == 1:
            points = points.reshape(1, -1)
        W = points.shape[0]
        names = self.sampling_params
        if points.shape[1] != len(names):
            raise ValueError("points must have %d columns (%s)"
                             % (len(names), ", ".join(names)))
        samples = {p: points[:, i] for i, p in enumerate(names)}
        params = self._transform_params(**samples)
        stats = OrderedDict()
        need_prior = callstat in ('logposterior', 'logplr', 'logprior',
                                  'logjacobian')
        logp = None
        if need_prior:
            logj, logp = self._batch_logprior(params, W)
            stats['logjacobian'] = logj
            stats['logprior'] = logp
        all_names = list(self.default_stats)
        if callstat in ('logprior', 'logjacobian'):
            res = BatchResult(logp if callstat == 'logprior' else stats['logjacobian'], stats, all_names)
            return PendingBatch(lambda: res) if defer else res
        skip = None
        if callstat in ('logposterior', 'logplr'):
            skip = numpy.isneginf(logp)
        # hints for models that can hand the sampled columns to a kernel as they are
        self._batch_hint = (points, samples, bool(with_stats), bool(defer))
        try:
            out = self._batch_callstat(callstat, params, W, skip, logp, stats)
        finally:
            self._batch_hint = None
        if isinstance(out, Deferred):
            return PendingBatch(lambda: self._wrap_batch(out.finish(), stats, all_names))
        res = self._wrap_batch(out, stats, all_names)
        return PendingBatch(lambda: res) if defer else res

    @staticmethod
    def _wrap_batch(out, stats, all_names):
        values, more = out
        if callable(more):
            base = stats

            def build(base=base, more=more):
                out = OrderedDict(base)
                out.update(more())
                return out
            return BatchResult(values, build, all_names)
        stats.update(more)
        return BatchResult(values, stats, all_names)

    def _batch_callstat(self, callstat, params, W, skip, logp, stats):
        if callstat not in ('logposterior', 'loglikelihood'):
            raise ValueError("unsupported callstat %r for this model" % callstat)
        logl, extra = self._batch_loglikelihood(params, W, skip)
        extra = OrderedDict(extra)
        extra['loglikelihood'] = logl
        if skip is not None and skip.any():
            # the scalar path never evaluates the likelihood of a prior-excluded point
            for k in extra:
                extra[k] = numpy.where(skip, numpy.nan, extra[k])
        if callstat == 'loglikelihood':
            return logl, extra
        return numpy.where(skip, -numpy.inf, logp + numpy.where(skip, 0., logl)), extra
