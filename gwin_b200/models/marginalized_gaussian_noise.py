"""Gaussian-noise likelihood marginalised over phase and/or distance.

Reference: ``gwin/models/marginalized_gaussian_noise.py`` -- constructor
:232-272, ``_setup_prior`` :355-389, ``_margdistphase_loglr`` :431-439,
``_margdist_loglr`` :441-448, ``_margphase_loglr`` :456-459, ``_loglr`` :461-511.
The phase, distance and distance+phase branches run on the CUDA path (the
distance grid is a second small kernel after the inner products, one warp per
walker over the reference's 10**4-point grid).  Time marginalisation needs an
FFT matched filter -- a different hot path -- and raises.

``MarginalizedPhaseGaussianNoise`` is the later ``pycbc.inference`` spelling of
the same model; it is the phase branch without the (unused, see reference
:382-389) ``marg_prior`` requirement.
"""
from collections import OrderedDict

import numpy

from .. import _lib
from .gaussian_noise import GaussianNoise


class MarginalizedGaussianNoise(GaussianNoise):
    name = 'marginalized_gaussian_noise'
    _mode = _lib.MODE_MARG_PHASE

    def __init__(self, variable_params, data, waveform_generator,
                 f_lower, psds=None, f_upper=None, norm=None,
                 time_marginalization=False, distance_marginalization=False,
                 phase_marginalization=False, marg_prior=None, **kwargs):
        super(MarginalizedGaussianNoise, self).__init__(
            variable_params, data, waveform_generator, f_lower, psds, f_upper,
            norm, **kwargs)
        self._margtime = time_marginalization
        self._margdist = distance_marginalization
        self._margphase = phase_marginalization
        self._args = (int(self._margdist), int(self._margphase),
                      int(self._margtime))
        if self._args == (0, 0, 0):
            raise AttributeError("This class requires that you marginalize "
                                 "over at least one parameter. You have not "
                                 "marginalized over any.")
        if self._margtime:
            raise NotImplementedError(
                "time marginalization (an FFT matched filter per point) is not "
                "on the CUDA path; phase and distance marginalization are")
        self._mode = {(1, 1, 0): _lib.MODE_MARG_DIST_PHASE,
                      (1, 0, 0): _lib.MODE_MARG_DIST,
                      (0, 1, 0): _lib.MODE_MARG_PHASE}[self._args]
        if marg_prior is None:
            raise AttributeError("No priors are specified for the "
                                 "marginalization. This is needed to "
                                 "calculated the marginalized likelihood")
        self._marg_prior = dict(zip([i.params[0] for i in marg_prior],
                                    marg_prior))
        self._setup_prior()

    def _setup_prior(self):
        """Reference :355-389.  The distance grid and ``deltad * dist_prior``
        go to the plan; the phase prior grid is built for parity of attributes
        only -- the reference never uses it (flat prior assumed)."""
        if len(self._marg_prior) == 0:
            raise ValueError("A prior must be specified for the parameters "
                             "that you wish to marginalize the likelihood "
                             "over")
        marg_number = len([i for i in self._args if i != 0])
        if len(self._marg_prior) != marg_number:
            raise AttributeError("There is not a prior for each keyword "
                                 "argument")
        if self._margdist:
            bounds = self._marg_prior["distance"].bounds
            self._dist_array = numpy.linspace(bounds["distance"].min,
                                              bounds["distance"].max, 10**4)
            self._deltad = self._dist_array[1] - self._dist_array[0]
            self.dist_prior = numpy.asarray(
                self._marg_prior["distance"].pdf(distance=self._dist_array),
                dtype=float)
            self._plan.set_distance_marg(self._dist_array,
                                         self._deltad * self.dist_prior)
        if self._margphase:
            bounds = self._marg_prior["phase"].bounds
            self._phase_array = numpy.linspace(bounds["phase"].min,
                                               bounds["phase"].max, 10**4)
            self._deltap = self._phase_array[1] - self._phase_array[0]
            self.phase_prior = numpy.asarray(
                self._marg_prior["phase"].pdf(phase=self._phase_array))

    @property
    def _extra_stats(self):
        return ['loglr'] + \
               ['{}_optimal_snrsq'.format(det) for det in self._data] + \
               ['{}_matchedfilter_snrsq'.format(det) for det in self._data]

    def _nowaveform_loglr(self):
        for det in self._data:
            setattr(self._current_stats, 'loglikelihood', -numpy.inf)
            setattr(self._current_stats, '{}_matchedfilter_snrsq'.format(det),
                    -numpy.inf)
            setattr(self._current_stats, '{}_optimal_snrsq'.format(det), 0.)
        return -numpy.inf

    def _loglr(self):
        lr, hd, hh = self._kernel_call(self.current_params, 1)
        if lr[0] == -numpy.inf:
            return self._nowaveform_loglr()
        for i, det in enumerate(self._dets):
            setattr(self._current_stats, '{}_optimal_snrsq'.format(det),
                    float(hh[0, i]))
            setattr(self._current_stats, '{}_matchedfilter_snrsq'.format(det),
                    complex(hd[0, i]))
        self._current_stats.loglikelihood = float(lr[0]) + self.lognl
        return float(lr[0])

    def _det_stats(self, hd, hh, dead):
        out = OrderedDict()
        for i, det in enumerate(self._dets):
            out['{}_optimal_snrsq'.format(det)] = hh[:, i]
            c = hd[:, i].copy()
            c[dead] = -numpy.inf
            out['{}_matchedfilter_snrsq'.format(det)] = c
        return out


class MarginalizedPhaseGaussianNoise(MarginalizedGaussianNoise):
    """``north_star`` spelling: phase marginalisation, no ``marg_prior``
    needed."""
    name = 'marginalized_phase'

    def __init__(self, variable_params, data, waveform_generator,
                 f_lower, psds=None, f_upper=None, norm=None, **kwargs):
        GaussianNoise.__init__(self, variable_params, data, waveform_generator,
                               f_lower, psds, f_upper, norm, **kwargs)
        self._margtime = self._margdist = False
        self._margphase = True
        self._args = (0, 1, 0)
        self._marg_prior = {}
