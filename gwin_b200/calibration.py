"""Calibration models for the template (reference ``gwin/calibration.py``).

The reference multiplies every detector's template by
``(1 + dA(f)) (2 + i dphi(f)) / (2 - i dphi(f))`` inside the waveform generator
(``Recalibrate.map_to_adjust`` :50-76, ``CubicSpline.apply_calibration``
:142-173; handed to the generator at ``models/base_data.py:225-230``), with
``dA`` and ``dphi`` ``scipy.interpolate.UnivariateSpline`` fits through
``n_points`` log-spaced values named ``recalib_amplitude_<ifo>_<i>`` and
``recalib_phase_<ifo>_<i>``.

Here the classes only *describe* the model -- the spline frequencies and the
parameter names; the factor itself is applied on the GPU, per walker, inside the
likelihood kernels and ``generate`` (``gwin_plan_set_calibration``,
``gwin_loglr_batch_calib``).  With scipy's default smoothing (``s = n_points``)
the fit through calibration-sized values (sum of squares <= ``n_points``; any
set of values below 1 in magnitude) is the least-squares cubic polynomial, and
that is what the kernels evaluate; ``values_matrix`` refuses values outside
that regime instead of silently differing from the reference.
"""
from abc import ABCMeta, abstractmethod

import numpy


class Recalibrate(object):
    __metaclass__ = ABCMeta
    name = None

    def __init__(self, ifo_name):
        self.ifo_name = ifo_name
        self.params = dict()

    @abstractmethod
    def apply_calibration(self, strain):
        return

    def map_to_adjust(self, strain, prefix='recalib_', **params):
        """Reference :50-76.  The CUDA generator applies the factor itself; a
        host-side adjustment of an arbitrary series is not part of this package."""
        raise NotImplementedError(
            "calibration is applied on the GPU inside "
            "FDomainDetFrameGenerator.generate and the likelihood kernels; "
            "there is no host implementation")

    @classmethod
    def from_config(cls, cp, ifo, section):
        """Reference :78-101.  ``section`` holds ``<ifo>_model = cubic_spline`` and the model's
        constructor arguments as ``<ifo>_<argument>``; the model is built for that detector."""
        tag = ifo.lower()
        kwargs = {}
        for key, value in cp.items(section):
            if tag in key:
                kwargs[key[len(ifo) + 1:]] = value
        kind = kwargs.pop('model')
        return all_models[kind](ifo_name=tag, **kwargs)


class CubicSpline(Recalibrate):
    """Reference :104-173."""
    name = 'cubic_spline'

    def __init__(self, minimum_frequency, maximum_frequency, n_points,
                 ifo_name):
        Recalibrate.__init__(self, ifo_name=ifo_name)
        self.n_points = int(n_points)
        if self.n_points < 4:
            raise ValueError(
                'Use at least 4 spline points for calibration model')
        # log-spaced nodes between the two frequencies (config values arrive as strings)
        lo, hi = numpy.log10(float(minimum_frequency)), numpy.log10(float(maximum_frequency))
        self.spline_points = numpy.logspace(lo, hi, self.n_points)

    def apply_calibration(self, strain):
        return self.map_to_adjust(strain)

    @property
    def param_names(self):
        """``recalib_*`` parameter names in kernel order: amplitudes, then phases."""
        return ['recalib_{}_{}_{}'.format(which, self.ifo_name, ii)
                for which in ('amplitude', 'phase')
                for ii in range(self.n_points)]


all_models = {
    CubicSpline.name: CubicSpline
}


def spline_frequencies(recalib, detectors):
    """[n_det, n_points] spline points in plan detector order."""
    models = [recalib[det] for det in detectors]
    for m in models:
        if not isinstance(m, CubicSpline):
            raise NotImplementedError("only the cubic_spline calibration model "
                                      "runs on the CUDA path")
    if len(set(m.n_points for m in models)) != 1:
        raise ValueError("all detectors must use the same number of "
                         "calibration spline points")
    return numpy.stack([m.spline_points for m in models])


def values_matrix(recalib, detectors, params, W, static=None):
    """Calibration values of ``W`` points as the kernels take them:
    [n_det * 2 * n_points, W], row ``(2 det + which) n_points + i``.  A missing
    ``recalib_*`` parameter is a ``KeyError``, as in the reference (:151-161)."""
    n = recalib[detectors[0]].n_points
    out = numpy.empty((len(detectors) * 2 * n, W), dtype=numpy.float64)
    r = 0
    for det in detectors:
        for name in recalib[det].param_names:
            if name in params:
                out[r] = params[name]
            elif static is not None and name in static:
                out[r] = static[name]
            else:
                raise KeyError(name)
            r += 1
    # the regime in which UnivariateSpline's default smoothing returns the
    # least-squares cubic: residual sum of squares <= s = n_points
    blocks = out.reshape(len(detectors) * 2, n, W)
    with numpy.errstate(invalid='ignore'):
        if numpy.any((blocks * blocks).sum(axis=1) > n):
            raise ValueError("calibration values this large (sum of squares > "
                             "n_points) leave the regime the CUDA path "
                             "reproduces the reference's spline fit in")
    return out
