"""Waveform-generator shim with the interface of
``pycbc.waveform.generator.FDomainDetFrameGenerator`` as the reference builds it
(``gwin/models/base_data.py:224-230``, ``test/utils/waveform.py:64-65``).

The object carries *configuration* (epoch, detectors, delta_f, f_lower,
approximant, frozen parameters).  On the likelihood path nothing here generates
an array: ``GaussianNoise`` hands the configuration to the CUDA plan, whose
kernel produces ``h_det(f_k)`` on the fly per frequency bin.  ``generate`` is
provided for the other users of the interface (building injections, plotting a
waveform): it runs the ``generate_kernel`` of the same CUDA library and copies
the detector-frame series back.
"""
import math
from collections import OrderedDict

import numpy

from . import _lib
from .calibration import spline_frequencies, values_matrix
from .detector import Detector
from .types import FrequencySeries

MTSUN_SI = 4.925491025543575903411922162094833998e-6

#: waveform parameter order of a row of the kernel's parameter matrix, with the
#: defaults pycbc's ``get_fd_waveform`` applies to parameters that are not given
WAVEFORM_PARAMS = OrderedDict([
    ("mass1", None), ("mass2", None), ("spin1z", 0.0), ("spin2z", 0.0),
    ("distance", 1.0), ("inclination", 0.0), ("coa_phase", 0.0),
    ("ra", None), ("dec", None), ("polarization", None), ("tc", None),
    ("lambda1", 0.0), ("lambda2", 0.0)])

SUPPORTED_APPROXIMANTS = ("TaylorF2",)

#: ``get_fd_waveform`` arguments that this path does not implement and that change a TaylorF2
#: waveform when they are not at the listed neutral value (pycbc's defaults): accepted as static
#: parameters at that value only, never as sampled ones
NEUTRAL_ONLY = {
    "f_ref": 0.0, "f_final": 0.0, "f_final_func": "", "phase_order": -1, "spin_order": -1,
    "tidal_order": -1, "amplitude_order": 0, "eccentricity": 0.0, "long_asc_nodes": 0.0,
    "mean_per_ano": 0.0, "spin1x": 0.0, "spin1y": 0.0, "spin2x": 0.0, "spin2y": 0.0,
    "dquad_mon1": 0.0, "dquad_mon2": 0.0, "lambda_octu1": None, "lambda_octu2": None,
    "quadfmode1": None, "quadfmode2": None, "octufmode1": None, "octufmode2": None,
    "frame_axis": 0, "modes_choice": 0, "side_bands": 0, "numrel_data": ""}
#: arguments that do not enter a frequency-domain TaylorF2 waveform at all
IGNORED = ("approximant", "delta_f", "delta_t", "f_lower", "taper", "fft_type", "t_gate_start",
           "t_gate_end")


def check_waveform_params(variable, static, transforms=None, recalib_names=()):
    """Every parameter that reaches the waveform generator must be one the kernel evaluates.

    In the reference any ``get_fd_waveform`` argument may be sampled or frozen
    (``models/base_data.py:84-161``; the generator forwards all of them).  Here the kernel reads
    the 13 TaylorF2 parameters of ``WAVEFORM_PARAMS`` (plus the ``recalib_*`` values); anything
    else would be ignored silently -- a flat marginal for a sampled parameter, a different
    likelihood than the reference's for a static one.  Raises ``ValueError`` instead."""
    consumed, produced = set(), set()
    for t in (transforms or ()):
        consumed.update(t.inputs)
        produced.update(t.outputs)
    known = set(WAVEFORM_PARAMS) | set(recalib_names)
    for name in (set(variable) - consumed) | produced:
        if name not in known:
            raise ValueError("variable parameter %r does not enter the TaylorF2 waveform of this "
                             "implementation (known: %s, recalib_*, or the inputs of a waveform "
                             "transform)" % (name, ", ".join(WAVEFORM_PARAMS)))
    for name, value in dict(static).items():
        if name in known or name in consumed or name in IGNORED:
            continue
        if name in NEUTRAL_ONLY:
            neutral = NEUTRAL_ONLY[name]
            if value is None or value == neutral or (neutral is None and not value):
                continue
            raise ValueError("static parameter %s=%r is not implemented by this TaylorF2 path "
                             "(only its neutral value %r is accepted)" % (name, value, neutral))
        raise ValueError("static parameter %r is not understood by the TaylorF2 waveform of this "
                         "implementation" % name)


class NoWaveformError(Exception):
    """No waveform could be generated for the given parameters
    (``pycbc.waveform.NoWaveformError``, caught at gaussian_noise.py:323)."""


class FDomainCBCGenerator(object):
    """Marker for "radiation-frame frequency-domain CBC generator"."""
    name = "fd_cbc"


def select_waveform_generator(approximant):
    """``pycbc.waveform.generator.select_waveform_generator``."""
    if approximant in SUPPORTED_APPROXIMANTS:
        return FDomainCBCGenerator
    raise ValueError("approximant %r is not supported by the CUDA path "
                     "(supported: %s)" % (approximant,
                                          ", ".join(SUPPORTED_APPROXIMANTS)))


def taylorf2_fisco(mass1, mass2):
    """f_ISCO = 6^(-3/2) / (pi M) as lalsimulation's TaylorF2 computes it."""
    piM = math.pi * ((mass1 + mass2) * MTSUN_SI)
    v_isco = 1.0 / math.sqrt(6.0)
    return v_isco * v_isco * v_isco / piM


def taylorf2_length(mass1, mass2, delta_f):
    """len(h) of the TaylorF2 series: ``int(f_ISCO/delta_f + 1)``."""
    return int(taylorf2_fisco(mass1, mass2) / delta_f + 1)


def param_matrix(params, W, static=None):
    """Assemble the kernel's [W, 13] parameter matrix from a dict of scalars
    and/or [W] arrays (plus ``static`` defaults)."""
    out = numpy.empty((W, len(WAVEFORM_PARAMS)), dtype=numpy.float64)
    for i, (name, default) in enumerate(WAVEFORM_PARAMS.items()):
        if name in params:
            v = params[name]
        elif static is not None and name in static:
            v = static[name]
        elif default is not None:
            v = default
        else:
            raise ValueError("waveform parameter %r was not provided" % name)
        out[:, i] = v
    return out


def param_rows(params, W, static=None):
    """Parameter-major [13, W] variant of ``param_matrix``: one contiguous row per
    waveform parameter (no transpose of the sampler's per-parameter arrays)."""
    out = numpy.empty((len(WAVEFORM_PARAMS), W), dtype=numpy.float64)
    for i, (name, default) in enumerate(WAVEFORM_PARAMS.items()):
        if name in params:
            v = params[name]
        elif static is not None and name in static:
            v = static[name]
        elif default is not None:
            v = default
        else:
            raise ValueError("waveform parameter %r was not provided" % name)
        out[i] = v
    return out


class FDomainDetFrameGenerator(object):
    location_args = set(['tc', 'ra', 'dec', 'polarization'])

    def __init__(self, rframe_generator_class, epoch, detectors=None,
                 variable_args=(), recalib=None, gates=None, **frozen_params):
        if rframe_generator_class is not FDomainCBCGenerator:
            raise ValueError("only FDomainCBCGenerator is supported")
        if gates:
            raise NotImplementedError("gating the data is outside the CUDA "
                                      "hot path")
        if not detectors:
            raise ValueError("at least one detector is required")
        approximant = frozen_params.pop("approximant", "TaylorF2")
        select_waveform_generator(approximant)
        try:
            self.delta_f = float(frozen_params.pop("delta_f"))
            self.f_lower = float(frozen_params.pop("f_lower"))
        except KeyError as e:
            raise ValueError("%s must be provided" % e)
        frozen_params.pop("delta_t", None)
        check_waveform_params((), frozen_params)
        self.approximant = approximant
        self._epoch = float(epoch)
        self.detectors = OrderedDict((d, Detector(d)) for d in detectors)
        self.detector_names = list(self.detectors.keys())
        self.variable_args = tuple(variable_args)
        self.frozen_params = dict(frozen_params)
        self.current_params = dict(frozen_params)
        self._plans = {}
        # dict detector -> calibration.Recalibrate ([UPSTREAM] pycbc generator
        # ``recalib=``; reference models/base_data.py:225-230)
        self.recalib = recalib if recalib else None
        if self.recalib:
            if sorted(self.recalib.keys()) != sorted(self.detector_names):
                raise ValueError("recalib needs one model per detector")
            self._spline_freqs = spline_frequencies(self.recalib,
                                                    self.detector_names)

    @property
    def epoch(self):
        return self._epoch

    @property
    def static_args(self):
        return self.frozen_params

    def _plan_for(self, n_f, device):
        key = (n_f, device)
        if key not in self._plans:
            nd = len(self.detectors)
            self._plans.clear()
            self._plans[key] = _lib.Plan(
                device, self.delta_f, self._epoch, self.f_lower, None,
                self.f_lower,
                numpy.stack([d.response for d in self.detectors.values()]),
                numpy.stack([d.location for d in self.detectors.values()]),
                numpy.zeros((nd, n_f), dtype=numpy.complex128), None, None)
            if self.recalib:
                self._plans[key].set_calibration(self._spline_freqs)
        return self._plans[key]

    def generate(self, device=0, **kwargs):
        """Returns ``{det: FrequencySeries}`` of length ``int(f_ISCO/df + 1)``
        (CUDA ``generate_kernel``)."""
        self.current_params.update(kwargs)
        row = param_matrix(self.current_params, 1)
        if not (row[0, 0] > 0 and row[0, 1] > 0) or \
                taylorf2_fisco(row[0, 0], row[0, 1]) <= self.f_lower:
            # lalsimulation: non-positive masses and "f_max <= fStart" are domain errors
            raise NoWaveformError("f_ISCO is not above f_lower for these masses")
        n = taylorf2_length(row[0, 0], row[0, 1], self.delta_f)
        n_f = max(((n + 1023) // 1024) * 1024 + 1,
                  int(self.f_lower / self.delta_f) + 9)
        calib = None
        if self.recalib:
            calib = values_matrix(self.recalib, self.detector_names,
                                  self.current_params, 1)[:, 0]
        h, length = self._plan_for(n_f, device).generate(row[0], calib=calib)
        return OrderedDict(
            (det, FrequencySeries(h[i, :length], self.delta_f, self._epoch))
            for i, det in enumerate(self.detector_names))
