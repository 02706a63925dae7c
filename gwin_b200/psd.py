"""Analytic PSD used to build synthetic data (host-side input preparation, not
on the likelihood path).  Mirrors ``pycbc.psd.aLIGOZeroDetHighPower(length,
delta_f, low_freq_cutoff)`` as used by the reference fixture
``test/utils/base.py:54-57``: the zero-detuned-high-power fit
``1e-48 (0.0152 x^-4 + 0.2935 x^(9/4) + 2.7951 x^(3/2) - 6.5080 x^(3/4) + 17.7622)``
with ``x = f/245.4`` (Ajith 2011), zero below the low-frequency cutoff.
"""
import numpy


def aLIGOZeroDetHighPower(length, delta_f, low_freq_cutoff):
    psd = numpy.zeros(int(length), dtype=numpy.float64)
    kmin = int(low_freq_cutoff / delta_f)
    x = numpy.arange(kmin, int(length), dtype=numpy.float64) * delta_f / 245.4
    psd[kmin:] = 1e-48 * (0.0152 * x ** -4.0 + 0.2935 * x ** 2.25 +
                          2.7951 * x ** 1.5 - 6.5080 * x ** 0.75 + 17.7622)
    return psd


def colored_noise(psd, delta_f, rng):
    """Frequency-domain Gaussian noise with one-sided PSD ``psd``:
    ``d_k = sqrt(S_k/(4 df)) (N(0,1) + i N(0,1))``; bins with ``S_k == 0`` are 0."""
    n = len(psd)
    sigma = numpy.sqrt(psd / (4.0 * delta_f))
    return sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
