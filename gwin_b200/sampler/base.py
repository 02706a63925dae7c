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
        """nwalkers x ndim initial positions in sampling-parameter space: the last iteration
        of ``samples_file`` if given (resume), else draws from the (given) prior."""
        nwalkers = self.nwalkers
        p0 = numpy.ones((nwalkers, len(self.variable_params)))
        if samples_file is not None:
            samples = samples_file.read_samples(self.sampling_params, iteration=-1,
                                                flatten=False, group='sampling_samples')
        else:
            samples = self.model.prior_rvs(size=nwalkers, prior=prior)
        for i, param in enumerate(self.sampling_params):
            p0[:, i] = samples[param]
        self._p0 = p0
        return p0

    # -- results / checkpoint (layout of the reference's InferenceFile) -----------------
    def write_metadata(self, fp, **kwargs):
        """reference sampler/base.py:152-186, 385-399"""
        fp.attrs['sampler'] = self.name
        fp.attrs['model'] = self.model.name
        fp.attrs['variable_params'] = list(self.variable_params)
        fp.attrs['sampling_params'] = list(self.sampling_params)
        fp.attrs['niterations'] = int(self.niterations)
        fp.attrs['nwalkers'] = int(self.nwalkers)
        if getattr(self, 'ntemps', None):
            fp.attrs['ntemps'] = int(self.ntemps)
        lognl = getattr(self.model, 'lognl', None)
        if lognl is not None:
            fp.attrs['lognl'] = float(lognl)
        fp.attrs['static_params'] = {k: v for k, v in self.model.static_params.items()}
        for k, v in kwargs.items():
            fp.attrs[k] = v

    def write_results(self, fp, **metadata):
        """Appends everything since the last ``clear_chain`` (reference
        ``write_results`` sampler/emcee.py:196-225): samples in variable-parameter space,
        the raw sampling-space positions (for resuming), model stats, acceptance fraction and
        the sampler state."""
        samples = self.samples
        for p in samples.dtype.names:
            fp.append('%s/%s' % (fp.samples_group, p), numpy.asarray(samples[p]))
        chain = self.chain
        for i, p in enumerate(self.sampling_params):
            fp.append('sampling_samples/%s' % p, chain[..., i])
        stats = self.model_stats
        if stats is not None:
            for p in stats.dtype.names:
                fp.append('%s/%s' % (fp.stats_group, p), numpy.asarray(stats[p]))
        fp['acceptance_fraction'] = numpy.asarray(self.acceptance_fraction)
        self.write_metadata(fp, **metadata)
        self.write_state(fp)
        fp.flush()

    def write_state(self, fp):
        """RNG state of the stepper (reference sampler/emcee.py:156-158)."""
        st = self._sampler.random_state
        fp['%s/random_state/keys' % fp.sampler_group] = st[1]
        fp.attrs['random_state'] = [st[0], int(st[2]), int(st[3]), float(st[4])]

    def set_state_from_file(self, fp):
        """reference sampler/emcee.py:160-167"""
        a = fp.attrs['random_state']
        state = (a[0], fp['%s/random_state/keys' % fp.sampler_group], int(a[1]), int(a[2]), float(a[3]))
        numpy.random.set_state(state)
        self._sampler.random_state = state
        self.lastclear = int(fp.attrs.get('niterations', 0))      # iteration numbering continues

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
