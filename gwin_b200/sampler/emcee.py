"""``EmceeEnsembleSampler`` / ``EmceePTSampler``: the reference's sampler wrappers
(``gwin/sampler/emcee.py:46-194, 253-432``) on top of the native steppers in
``ensemble.py`` instead of the emcee package.  Names, constructor arguments and
the ``from_cli / set_p0 / run / chain / lnpost / model_stats`` surface are kept;
every ensemble half-step becomes one batched likelihood launch (optionally
sharded over GPUs through the pool)."""
import numpy

from .. import models as _models
from ..distributions import FieldArray
from ..pool import BatchPool
from .base import BaseMCMCSampler
from .ensemble import EnsembleStepper, PTStepper


def _as_batch(model_call, pool):
    """points -> BatchResult-like evaluation through the pool's sharding."""
    shard = getattr(pool, "shard", None)

    def run(points, callstat=None):
        target = model_call
        if target is _models._call_global_model:
            target = _models._global_instance
        if shard is not None:
            values, stats, names = shard.evaluate(target, points, callstat=callstat)
            return values, stats
        res = target.batch(points, callstat=callstat) if callstat is not None \
            else target.batch(points)
        return res.values, res.stats
    return run


class _callprior(object):
    """Calls the model's prior only (reference ``sampler/emcee.py:230-240``); handed to a
    parallel-tempered sampler that wants prior and likelihood separately."""

    def __init__(self, model_call):
        self.callable = model_call
        self.return_all_stats = False

    def __call__(self, args):
        return self.callable(args, callstat='logprior', return_all_stats=False)

    def batch(self, points, callstat=None):
        return self.callable.batch(points, callstat='logprior')


class _callloglikelihood(object):
    """Calls the model's log likelihood only (reference ``sampler/emcee.py:243-250``)."""

    def __init__(self, model_call):
        self.callable = model_call
        self.return_all_stats = False

    def __call__(self, args):
        return self.callable(args, callstat='loglikelihood', return_all_stats=False)

    def batch(self, points, callstat=None):
        return self.callable.batch(points, callstat='loglikelihood')


class EmceeEnsembleSampler(BaseMCMCSampler):
    name = "emcee"

    def __init__(self, model, nwalkers, pool=None, model_call=None):
        if model_call is None:
            model_call = _models.CallModel(model, 'logposterior')
        self._pool = pool if pool is not None else BatchPool()
        self._model_call = model_call
        evaluate = _as_batch(model_call, self._pool)
        stat_names = list(model.default_stats)

        def lnpostfn(points):
            values, stats = evaluate(points)
            cols = [numpy.asarray(stats[n]).tolist() for n in stat_names]
            blobs = list(zip(*cols)) if cols else None
            return values, blobs

        ndim = len(model.variable_params)
        sampler = EnsembleStepper(nwalkers, ndim, lnpostfn)
        # the reference syncs emcee's RNG with numpy's global state (:79-80)
        sampler.random_state = numpy.random.get_state()
        super(EmceeEnsembleSampler, self).__init__(sampler, model)
        self._nwalkers = nwalkers

    @classmethod
    def from_cli(cls, opts, model, pool=None, model_call=None):
        return cls(model, opts.nwalkers, pool=pool, model_call=model_call)

    @property
    def lnpost(self):
        """nwalkers x niterations log posterior."""
        return self._sampler.lnprobability

    @property
    def chain(self):
        """nwalkers x niterations x ndim."""
        return self._sampler.chain

    def clear_chain(self):
        self.lastclear = self.niterations
        self._sampler.reset()
        self._sampler.clear_blobs()

    def set_p0(self, samples_file=None, prior=None):
        p0 = super(EmceeEnsembleSampler, self).set_p0(samples_file=samples_file,
                                                      prior=prior)
        self._sampler.random_state = numpy.random.get_state()
        self._pos = None
        return p0

    def run(self, niterations, **kwargs):
        pos = self._pos
        if pos is None:
            pos = self.p0
        res = self._sampler.run_mcmc(pos, niterations, **kwargs)
        p, lnpost, rstate = res[0], res[1], res[2]
        self._pos = p
        return p, lnpost, rstate


class EmceePTSampler(BaseMCMCSampler):
    name = "emcee_pt"

    def __init__(self, model, ntemps, nwalkers, pool=None, model_call=None,
                 betas=None):
        if model_call is None:
            model_call = _models.CallModel(model, 'logposterior')
        self._pool = pool if pool is not None else BatchPool()
        self._model_call = model_call
        evaluate = _as_batch(model_call, self._pool)

        def logl_logp(points):
            # one launch gives both: the reference pays two update() calls per
            # point here (_callloglikelihood + _callprior, :230-250)
            values, stats = evaluate(points, callstat='logposterior')
            return stats['loglikelihood'], stats['logprior']

        ndim = len(model.variable_params)
        sampler = PTStepper(ntemps, nwalkers, ndim, logl_logp, betas=betas)
        sampler.random_state = numpy.random.get_state()
        super(EmceePTSampler, self).__init__(sampler, model)
        self._nwalkers = nwalkers
        self._ntemps = ntemps

    @classmethod
    def from_cli(cls, opts, model, pool=None, model_call=None):
        return cls(model, opts.ntemps, opts.nwalkers, pool=pool,
                   model_call=model_call)

    @property
    def ntemps(self):
        return self._ntemps

    @property
    def chain(self):
        """ntemps x nwalkers x niterations x ndim."""
        return self._sampler.chain

    @property
    def lnpost(self):
        return self._sampler.lnprobability

    def clear_chain(self):
        self.lastclear = self.niterations
        self._sampler.reset()

    @property
    def model_stats(self):
        """loglr and logprior reconstructed as the reference does (:337-356)."""
        logl = self._sampler.lnlikelihood
        logp = self._sampler.lnprobability - logl * self._sampler.betas.reshape(-1, 1, 1)
        lognl = getattr(self.model, 'lognl', 0.)
        return FieldArray.from_kwargs(loglr=logl - lognl, logprior=logp)

    def set_p0(self, samples_file=None, prior=None):
        ntemps, nwalkers = self.ntemps, self.nwalkers
        p0 = numpy.ones((ntemps, nwalkers, len(self.variable_params)))
        if samples_file is not None:
            samples = samples_file.read_samples(self.sampling_params, iteration=-1,
                                                flatten=False, group='sampling_samples')
        else:
            samples = self.model.prior_rvs(size=nwalkers * ntemps, prior=prior)
        for i, param in enumerate(self.sampling_params):
            p0[..., i] = numpy.asarray(samples[param]).reshape((ntemps, nwalkers))
        self._p0 = p0
        self._pos = None
        self._sampler.random_state = numpy.random.get_state()
        return p0

    def run(self, niterations, **kwargs):
        pos = self._pos
        if pos is None:
            pos = self.p0
        p, lnpost, logl = self._sampler.run_mcmc(pos, niterations, **kwargs)
        self._pos = p
        return p, lnpost, self._sampler.random_state

    def calculate_logevidence(self, fburnin=0.1):
        """Thermodynamic-integration evidence (reference :902-954)."""
        return self._sampler.thermodynamic_integration_log_evidence(fburnin=fburnin)
