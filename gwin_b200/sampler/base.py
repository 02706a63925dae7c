"""Shared sampler-wrapper behaviour: initial positions from the prior, samples
and model stats as field arrays (reference ``gwin/sampler/base.py``: ``set_p0``
:285-319, ``samples`` :338-364, ``model_stats`` :366-382).  Result-file I/O and
autocorrelation analysis of the reference base class are outside the hot path
 and not provided."""
import numpy

from ..distributions import FieldArray


class BaseMCMCSampler(object):
    name = None

    def __init__(self, sampler, model):
        self._sampler = sampler
        self.model = model
        self._pos = None
        self._p0 = None
        self._nwalkers = None
        self.lastclear = 0

    @property
    def sampler(self):
        return self._sampler

    @property
    def variable_params(self):
        return self.model.variable_params

    @property
    def sampling_params(self):
        return self.model.sampling_params

    @property
    def nwalkers(self):
        return self._nwalkers

    @property
    def pos(self):
        return self._pos

    @property
    def p0(self):
        if self._p0 is None:
            raise ValueError("initial positions not set; run set_p0")
        return self._p0

    @property
    def niterations(self):
        """Iterations advanced since the start (chain clears included)."""
        return self._sampler.iterations + self.lastclear

    @property
    def acceptance_fraction(self):
        return self._sampler.acceptance_fraction

    def set_p0(self, samples_file=None, prior=None):
        """nwalkers x ndim initial positions drawn from the (given) prior, in
        sampling-parameter space."""
        if samples_file is not None:
            raise NotImplementedError("result files are outside the hot path")
        nwalkers = self.nwalkers
        p0 = numpy.ones((nwalkers, len(self.variable_params)))
        samples = self.model.prior_rvs(size=nwalkers, prior=prior)
        for i, param in enumerate(self.sampling_params):
            p0[:, i] = samples[param]
        self._p0 = p0
        return p0

    @property
    def samples(self):
        """Chain as a FieldArray ([extra dims x] nwalkers x niterations), boundary
        conditions applied, variable params added when sampling transforms are
        in use."""
        chain = self.chain
        names = list(self.sampling_params)
        samples = {p: chain[..., i] for i, p in enumerate(names)}
        samples = self.model.prior_distribution.apply_boundary_conditions(**samples)
        if self.model.sampling_transforms is not None:
            samples = self.model.sampling_transforms.apply(samples, inverse=True)
            names = list(samples.keys())
        return FieldArray.from_arrays([numpy.asarray(samples[p]) for p in names],
                                      names=names)

    @property
    def model_stats(self):
        """Model stats carried as blobs, nwalkers x niterations; ``None`` if the
        model returned none.  Forced to float like the reference (:377-381), so
        complex per-detector stats keep only their real part."""
        blobs = self._sampler.blobs
        if len(blobs) == 0:
            return None
        stats = numpy.array(blobs, dtype=complex)          # [niter, nwalkers, nstats]
        arrays = {field: stats[..., fi].real.astype(float)
                  for fi, field in enumerate(self.model.default_stats)}
        return FieldArray.from_kwargs(**arrays).transpose()
