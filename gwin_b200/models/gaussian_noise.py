"""``GaussianNoise``: stationary-Gaussian-noise likelihood, evaluated for whole
batches of walkers by the sm_100a kernel.

Constructor, error behaviour, stat names and properties follow the reference
class (``gwin/models/gaussian_noise.py:219-434``).  What differs is *where* the
arithmetic runs: ``__init__`` builds a CUDA plan (the reference's whitening of
the data, :247-265, becomes a one-off table build), ``_loglr`` /
``evaluate_batch`` call ``gwin_loglr_batch`` through the C ABI.  There is no CPU
implementation of the likelihood in this package.
"""
import os
from collections import OrderedDict

import numpy

from .. import _lib
from ..calibration import spline_frequencies, values_matrix
from ..waveform import WAVEFORM_PARAMS, check_waveform_params, param_rows
from .base import Deferred
from .base_data import BaseDataModel


def _default_device():
    return int(os.environ.get("LOCAL_RANK", "0"))


class GaussianNoise(BaseDataModel):
    name = 'gaussian_noise'
    _mode = _lib.MODE_GAUSSIAN

    def __init__(self, variable_params, data, waveform_generator,
                 f_lower, psds=None, f_upper=None, norm=None, device=None,
                 **kwargs):
        super(GaussianNoise, self).__init__(variable_params, data,
                                            waveform_generator, **kwargs)
        # same detectors in data and generator (reference :227-233)
        if (sorted(waveform_generator.detectors.keys()) !=
                sorted(self._data.keys())):
            raise ValueError(
                "waveform generator's detectors ({0}) does not "
                "match data ({1})".format(
                    ','.join(sorted(waveform_generator.detector_names)),
                    ','.join(sorted(self._data.keys()))))
        # same epoch (reference :235-238)
        if any(waveform_generator.epoch != d.epoch
               for d in self._data.values()):
            raise ValueError("waveform generator does not have the same epoch "
                             "as all of the data sets.")
        # same lengths (reference :240-242)
        dlens = numpy.array([len(d) for d in data.values()])
        if not all(dlens == dlens[0]):
            raise ValueError("all data must be of the same length")
        # everything that reaches the generator must be something the kernel evaluates
        recalib = getattr(waveform_generator, 'recalib', None)
        recalib_names = []
        if recalib:
            for rec in recalib.values():
                recalib_names += rec.param_names
        check_waveform_params(self.variable_params, self.static_params,
                              self._waveform_transforms, recalib_names)
        self._dets = list(self._data.keys())
        d = self._data[self._dets[0]]
        N = len(d)
        delta_f = d.delta_f
        if norm is None:
            norm = 4 * delta_f
        psd_arr = None
        if psds is not None:
            psd_arr = numpy.stack([numpy.asarray(psds[det], dtype=numpy.float64)
                                   for det in self._dets])
            if psd_arr.shape[1] != N:
                raise ValueError("psds must have the same length as the data")
        gen = waveform_generator
        self._device = _default_device() if device is None else int(device)
        # plan build == kmin/kmax, weight, whitening (reference :247-265)
        self._plan = _lib.Plan(
            self._device, delta_f, gen.epoch, f_lower, f_upper, gen.f_lower,
            numpy.stack([gen.detectors[det].response for det in self._dets]),
            numpy.stack([gen.detectors[det].location for det in self._dets]),
            numpy.stack([numpy.asarray(self._data[det].numpy(),
                                       dtype=numpy.complex128)
                         for det in self._dets]),
            psd_arr, norm)
        # calibration model held by the generator (reference base_data.py:225-230)
        self._recalib = getattr(gen, 'recalib', None)
        if self._recalib:
            self._plan.set_calibration(
                spline_frequencies(self._recalib, self._dets))
        self._kmin = self._plan.kmin
        self._kmax = self._plan.kmax
        self._f_lower = f_lower
        self._f_upper = f_upper
        self._norm = norm
        # keep the reference's visible side effect: ``model.data`` is whitened
        kmin, kmax = self._kmin, self._kmax
        if psds is None:
            self._weight = {det: numpy.sqrt(norm) * numpy.ones(N)
                            for det in self._dets}
        else:
            with numpy.errstate(divide='ignore'):
                self._weight = {det: numpy.sqrt(norm / psd_arr[i])
                                for i, det in enumerate(self._dets)}
        for det in self._dets:
            self._data[det][kmin:kmax] = self._data[det][kmin:kmax] * \
                self._weight[det][kmin:kmax]

    # -- introspection ------------------------------------------------------
    @property
    def plan(self):
        """The CUDA plan (device tables + workspace) behind this model."""
        return self._plan

    @property
    def detectors(self):
        return list(self._dets)

    @property
    def _extra_stats(self):
        return ['loglr'] + \
               ['{}_cplx_loglr'.format(det) for det in self._data] + \
               ['{}_optimal_snrsq'.format(det) for det in self._data]

    # -- lognl ------------------------------------------------------------
    def _lognl(self):
        return float(self._plan.lognl)

    def det_lognl(self, det):
        return float(self._plan.det_lognl[self._dets.index(det)])

    # -- scalar protocol ----------------------------------------------------
    def _static_for_kernel(self):
        static = dict(self._waveform_generator.frozen_params)
        static.update(self._static_params)
        return static

    def _kernel_call(self, params, W, skip=None, stats=True, defer=False):
        """One launch for W points.  The kernel's 13-parameter rows are assembled on the device
        (``gwin_loglr_batch_points_host``): when the sampled columns reach the waveform unchanged
        -- no transforms, no boundary condition touched them -- the sampler's own ``points`` array
        is handed over as it is; otherwise the parameters that vary go through a page-locked
        block, one row each.  Parameters that do not vary travel as 13 scalars."""
        static = self._static_for_kernel()
        if self._recalib:
            P = param_rows(params, W, static=static)
            calib = values_matrix(self._recalib, self._dets, params, W, static=static)
            return self._plan.loglr_host(P, mode=self._mode, skip=skip,
                                         param_major=True, calib=calib)
        hint = getattr(self, '_batch_hint', None)
        points, samples = (hint[0], hint[1]) if hint is not None else (None, {})
        defer = bool(defer and hint is not None)
        colmap = numpy.full(len(WAVEFORM_PARAMS), -1, dtype=numpy.int32)
        fixed = numpy.zeros(len(WAVEFORM_PARAMS))
        varying = []
        direct = points is not None
        for i, (name, default) in enumerate(WAVEFORM_PARAMS.items()):
            if name in params and numpy.ndim(params[name]) > 0:
                varying.append((i, name))
                direct = direct and params[name] is samples.get(name)
                continue
            v = params.get(name, static.get(name, default))
            if v is None:
                raise ValueError("waveform parameter %r was not provided" % name)
            fixed[i] = v
        if direct:
            order = {name: j for j, name in enumerate(self.sampling_params)}
            for i, name in varying:
                colmap[i] = order[name]
            return self._plan.loglr_points_host(points, colmap, fixed, mode=self._mode,
                                                skip=skip, stats=stats, defer=defer)
        if defer:
            # a deferred call keeps its (page-locked) rows until it has been waited for
            rows = _lib.pinned_empty((len(varying), W))
            for j, (i, name) in enumerate(varying):
                rows[j] = params[name]
                colmap[i] = j
            return self._plan.loglr_points_host(rows, colmap, fixed, mode=self._mode,
                                                skip=skip, stats=stats, col_major=True, defer=True)
        key = (len(varying), W)
        if getattr(self, '_rows_key', None) != key:
            self._rows_key, self._rows = key, _lib.pinned_empty(key)
        for j, (i, name) in enumerate(varying):
            self._rows[j] = params[name]
            colmap[i] = j
        return self._plan.loglr_points_host(self._rows, colmap, fixed, mode=self._mode,
                                            skip=skip, stats=stats, col_major=True)

    def _nowaveform_loglr(self):
        for det in self._data:
            setattr(self._current_stats, 'loglikelihood', -numpy.inf)
            setattr(self._current_stats, '{}_cplx_loglr'.format(det),
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
            setattr(self._current_stats, '{}_cplx_loglr'.format(det),
                    complex(hd[0, i] - 0.5 * hh[0, i]))
        self._current_stats.loglikelihood = float(lr[0]) + self.lognl
        return float(lr[0])

    def _loglikelihood(self):
        return self.loglr + self.lognl

    def det_cplx_loglr(self, det):
        try:
            return getattr(self._current_stats, '{}_cplx_loglr'.format(det))
        except AttributeError:
            self.loglr
            return getattr(self._current_stats, '{}_cplx_loglr'.format(det))

    def det_optimal_snrsq(self, det):
        try:
            return getattr(self._current_stats, '{}_optimal_snrsq'.format(det))
        except AttributeError:
            self.loglr
            return getattr(self._current_stats, '{}_optimal_snrsq'.format(det))

    # -- batched protocol -----------------------------------------------------
    def _det_stats(self, hd, hh, dead):
        out = OrderedDict()
        for i, det in enumerate(self._dets):
            c = hd[:, i] - 0.5 * hh[:, i]
            c[dead] = -numpy.inf
            out['{}_cplx_loglr'.format(det)] = c
            out['{}_optimal_snrsq'.format(det)] = hh[:, i]
        return out

    def _batch_callstat(self, callstat, params, W, skip, logp, stats):
        if callstat not in ('logposterior', 'loglikelihood', 'loglr', 'logplr'):
            raise ValueError("unsupported callstat %r" % callstat)
        hint = getattr(self, '_batch_hint', None)
        want = hint[2] if hint is not None else True
        defer = bool(hint[3]) if hint is not None and len(hint) > 3 else False
        res = self._kernel_call(params, W, skip=skip, stats=want, defer=defer)
        if defer and not isinstance(res, tuple):
            # queued: the rest runs when the caller asks for the result
            return Deferred(lambda: self._finish_callstat(callstat, params, W, skip, logp, res.wait()))
        return self._finish_callstat(callstat, params, W, skip, logp, res)

    def _finish_callstat(self, callstat, params, W, skip, logp, res):
        lr, hd, hh = res
        lognl = self.lognl
        logl = lr + lognl
        held = {'hd': hd, 'hh': hh}

        def extra():
            if held['hd'] is None:
                # the caller said it would not read the statistics, and does: evaluate again
                _, held['hd'], held['hh'] = self._kernel_call(params, W, skip=skip, stats=True)
            dead = numpy.isneginf(lr)
            out = self._det_stats(held['hd'], held['hh'], dead)
            out['loglr'] = lr
            out['loglikelihood'] = logl
            if skip is not None and skip.any():
                # the scalar path never evaluates these for prior-excluded points
                for k in out:
                    out[k] = numpy.where(skip, numpy.nan, out[k])
            return out
        if callstat == 'loglr':
            return lr, extra
        if callstat == 'loglikelihood':
            return logl, extra
        base = lr if callstat == 'logplr' else logl
        return numpy.where(skip, -numpy.inf,
                           logp + numpy.where(skip, 0., base)), extra
