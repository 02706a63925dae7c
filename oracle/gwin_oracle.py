"""CPU oracle for gwin's batched-loglr hot path.  TEST INFRASTRUCTURE ONLY.

This module is a float64 numpy restatement of what the reference does when
``GaussianNoise._loglr`` / ``MarginalizedGaussianNoise._loglr`` are called for
the TaylorF2 approximant.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product path (``gwin_b200``) never does and fails loudly without its CUDA
library.

PARITY UNPINNED.  The reference (``/root/reference``) is Python-2-only and its
arithmetic lives in un-vendored third-party packages -- ``pycbc>=1.12.1``
(``setup.py:44``) and unpinned ``lalsuite`` (``setup.py:50-53``) -- neither of
which is installed here, and the reference's own tests pin no numeric loglr
(only ``argmax_tc == tc_inj``, ``test/test_models.py:105-117``).  The
third-party arithmetic below (tagged [UPSTREAM]) is restated from the published
pycbc / lalsimulation algorithms; the gwin-level control flow follows the
reference line by line and each function cites the file:line it follows.

Two precision modes:

* ``faithful`` -- float64 everywhere, same pass structure and operation order
  as the reference stack (lalsim per-bin loop -> ``fp*hp+fc*hc`` -> phasor
  recurrence time shift -> whitening multiply -> two inner products).
* ``truth``    -- x87 80-bit ``numpy.longdouble`` phases with exact modular
  reduction of the time-shift term; used to show that the faithful mode (and the
  CUDA kernel) sit on the reference's own ~1e-9 rad noise floor.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

# ---------------------------------------------------------------------------
# [UPSTREAM] LAL constants (lal/LALConstants.h, CODATA-2014-era values)
# ---------------------------------------------------------------------------
LAL_PI = 3.141592653589793238462643383279502884
LAL_TWOPI = 6.283185307179586476925286766559005768
LAL_PI_4 = 0.785398163397448309615660845819875721
LAL_GAMMA = 0.577215664901532860606512090082402431
LAL_C_SI = 299792458.0
LAL_MSUN_SI = 1.988546954961461467461011951140572744e30
LAL_MTSUN_SI = 4.925491025543575903411922162094833998e-6
LAL_MRSUN_SI = 1.476625061404649406193430731479084713e3
LAL_PC_SI = 3.085677581491367278913937957796471611e16

XLAL_EPOCH_J2000_0_JD = 2451545.0

# GPS seconds at which GPS-UTC increased by one (lal/XLALLeapSeconds.h).
LEAP_SECONDS_GPS = (
    46828800, 78364801, 109900802, 173059203, 252028804, 315187205,
    346723206, 394156807, 425692808, 457228809, 504489610, 551750411,
    599184012, 820108813, 914803214, 1025136015, 1119744016, 1167264017,
)


# ---------------------------------------------------------------------------
# [UPSTREAM] pycbc.filter.get_cutoff_indices, called at gaussian_noise.py:247
# ---------------------------------------------------------------------------
def get_cutoff_indices(flow, fhigh, df, N):
    if flow:
        kmin = int(flow / df)
        if kmin < 0:
            raise ValueError("low frequency cutoff must be >= 0")
    else:
        kmin = 1
    if fhigh:
        kmax = int(fhigh / df)
        if kmax > int((N + 1) / 2.):
            kmax = int((N + 1) / 2.)
    else:
        kmax = int((N + 1) / 2.)
    if kmax <= kmin:
        raise ValueError("kmax <= kmin")
    return kmin, kmax


# ---------------------------------------------------------------------------
# [UPSTREAM] pycbc.psd.aLIGOZeroDetHighPower(length, delta_f, low_freq_cutoff)
# used by the reference fixture test/utils/base.py:54-57.  Analytic ZDHP fit
# (Ajith, arXiv:1107.1267 eq. 4.7).  pycbc fills bins k >= int(flow/df) and
# leaves zeros below.
# ---------------------------------------------------------------------------
def aligo_zdhp_psd(length, delta_f, low_freq_cutoff):
    psd = np.zeros(length, dtype=np.float64)
    kmin = int(low_freq_cutoff / delta_f)
    f = np.arange(kmin, length, dtype=np.float64) * delta_f
    x = f / 245.4
    psd[kmin:] = 1e-48 * (0.0152 * x ** -4.0 + 0.2935 * x ** 2.25 +
                          2.7951 * x ** 1.5 - 6.5080 * x ** 0.75 + 17.7622)
    return psd


# ---------------------------------------------------------------------------
# [UPSTREAM] lalsimulation XLALSimInspiralPNPhasing_F2 (3.5PN, aligned spins,
# spin order ALL, tides 0, qm_def = 1)
# ---------------------------------------------------------------------------
def taylorf2_phasing(m1, m2, chi1L=0.0, chi2L=0.0):
    """Returns dict with v[0..7], vlogv[5], vlogv[6] already multiplied by
    pfaN = 3/(128 eta)."""
    mtot = m1 + m2
    d = (m1 - m2) / (m1 + m2)
    eta = m1 * m2 / mtot / mtot
    m1M = m1 / mtot
    m2M = m2 / mtot
    pfaN = 3.0 / (128.0 * eta)
    v = [0.0] * 8
    vlogv = [0.0] * 8
    v[0] = 1.0
    v[1] = 0.0
    v[2] = 5.0 * (743.0 / 84.0 + 11.0 * eta) / 9.0
    v[3] = -16.0 * LAL_PI
    v[4] = 5.0 * (3058.673 / 7.056 + 5429.0 / 7.0 * eta +
                  617.0 * eta * eta) / 72.0
    v[5] = 5.0 / 9.0 * (7729.0 / 84.0 - 13.0 * eta) * LAL_PI
    vlogv[5] = 5.0 / 3.0 * (7729.0 / 84.0 - 13.0 * eta) * LAL_PI
    v[6] = (11583.231236531 / 4.694215680 - 640.0 / 3.0 * LAL_PI * LAL_PI -
            6848.0 / 21.0 * LAL_GAMMA) + \
        eta * (-15737.765635 / 3.048192 + 2255.0 / 12.0 * LAL_PI * LAL_PI) + \
        eta * eta * 76055.0 / 1728.0 - eta * eta * eta * 127825.0 / 1296.0
    v[6] += (-6848.0 / 21.0) * math.log(4.0)
    vlogv[6] = -6848.0 / 21.0
    v[7] = LAL_PI * (77096675.0 / 254016.0 + 378515.0 / 1512.0 * eta -
                     74045.0 / 756.0 * eta * eta)
    # aligned-spin terms
    qm1 = qm2 = 1.0
    chi1sq = chi1L * chi1L
    chi2sq = chi2L * chi2L
    chi1dotchi2 = chi1L * chi2L
    SL = m1M * m1M * chi1L + m2M * m2M * chi2L
    dSigmaL = d * (m2M * chi2L - m1M * chi1L)
    pn_sigma = eta * (721.0 / 48.0 * chi1L * chi2L -
                      247.0 / 48.0 * chi1dotchi2)
    pn_sigma += (720.0 * qm1 - 1.0) / 96.0 * m1M * m1M * chi1L * chi1L
    pn_sigma += (720.0 * qm2 - 1.0) / 96.0 * m2M * m2M * chi2L * chi2L
    pn_sigma -= (240.0 * qm1 - 7.0) / 96.0 * m1M * m1M * chi1sq
    pn_sigma -= (240.0 * qm2 - 7.0) / 96.0 * m2M * m2M * chi2sq
    pn_ss3 = (326.75 / 1.12 + 557.5 / 1.8 * eta) * eta * chi1L * chi2L
    pn_ss3 += ((4703.5 / 8.4 + 2935.0 / 6.0 * m1M - 120.0 * m1M * m1M) * qm1 +
               (-4108.25 / 6.72 - 108.5 / 1.2 * m1M +
                125.5 / 3.6 * m1M * m1M)) * m1M * m1M * chi1sq
    pn_ss3 += ((4703.5 / 8.4 + 2935.0 / 6.0 * m2M - 120.0 * m2M * m2M) * qm2 +
               (-4108.25 / 6.72 - 108.5 / 1.2 * m2M +
                125.5 / 3.6 * m2M * m2M)) * m2M * m2M * chi2sq
    pn_gamma = (554345.0 / 1134.0 + 110.0 * eta / 9.0) * SL + \
        (13915.0 / 84.0 - 10.0 * eta / 3.0) * dSigmaL
    v[7] += (-8980424995.0 / 762048.0 + 6586595.0 * eta / 756.0 -
             305.0 * eta * eta / 36.0) * SL - \
        (170978035.0 / 48384.0 - 2876425.0 * eta / 672.0 -
         4735.0 * eta * eta / 144.0) * dSigmaL
    v[6] += LAL_PI * (3760.0 * SL + 1490.0 * dSigmaL) / 3.0 + pn_ss3
    v[5] += -1.0 * pn_gamma
    vlogv[5] += -3.0 * pn_gamma
    v[4] += -10.0 * pn_sigma
    v[3] += 188.0 * SL / 3.0 + 25.0 * dSigmaL
    return {"v": [pfaN * x for x in v],
            "vlogv": [pfaN * x for x in vlogv],
            "eta": eta, "mtot": mtot}


def taylorf2_length(m1, m2, delta_f):
    """[UPSTREAM] XLALSimInspiralTaylorF2: n = (size_t)(fISCO/deltaF + 1)."""
    m_sec = (m1 + m2) * LAL_MTSUN_SI
    piM = LAL_PI * m_sec
    vISCO = 1.0 / math.sqrt(6.0)
    fISCO = vISCO * vISCO * vISCO / piM
    return int(fISCO / delta_f + 1)


def taylorf2_tidal(m1, m2, lambda1, lambda2):
    """[UPSTREAM] XLALSimInspiralTaylorF2Phasing_{10,12}PNTidalCoeff: returns (pft10, pft12)
    WITHOUT the pfaN prefactor."""
    mtot = m1 + m2
    out = []
    for coeff in (lambda x: (-288.0 + 264.0 * x) * x ** 4,
                  lambda x: (-15895.0 / 28.0 + 4595.0 / 28.0 * x + 5715.0 / 14.0 * x * x
                             - 325.0 / 7.0 * x * x * x) * x ** 4):
        out.append(lambda1 * coeff(m1 / mtot) + lambda2 * coeff(m2 / mtot))
    return tuple(out)


def taylorf2_hp_hc(m1, m2, s1z, s2z, distance, inclination, coa_phase,
                   delta_f, f_lower, mode="faithful", lambda1=0.0, lambda2=0.0):
    """[UPSTREAM] pycbc.waveform.get_fd_waveform(approximant='TaylorF2')
    = XLALSimInspiralChooseFDWaveform -> XLALSimInspiralTaylorF2 (+Core) with
    f_ref = 0, f_max = 0, 3.5PN phase, 0PN amplitude.  Returns complex128
    (hp, hc) of length n = int(fISCO/df + 1); bins below ceil(f_lower/df) are 0.
    masses in Msun, distance in Mpc.
    """
    ph = taylorf2_phasing(m1, m2, s1z, s2z)
    pv, pl, eta = ph["v"], ph["vlogv"], ph["eta"]
    pfaN = 3.0 / (128.0 * eta)
    t10, t12 = taylorf2_tidal(m1, m2, lambda1, lambda2)
    pft10, pft12 = pfaN * t10, pfaN * t12
    m = m1 + m2
    m_sec = m * LAL_MTSUN_SI
    piM = LAL_PI * m_sec
    n = taylorf2_length(m1, m2, delta_f)
    i_start = int(math.ceil(f_lower / delta_f))
    r = distance * 1e6 * LAL_PC_SI
    # lalsim is handed SI masses and divides back; m1*m2 here is in Msun^2
    amp0 = -4.0 * m1 * m2 / r * LAL_MRSUN_SI * LAL_MTSUN_SI * \
        math.sqrt(LAL_PI / 12.0)
    FTaN = 32.0 * eta * eta / 5.0
    dETaN = 2.0 * (-eta / 2.0)
    htilde = np.zeros(n, dtype=np.complex128)
    if i_start < n:
        k = np.arange(i_start, n)
        if mode == "faithful":
            f = k.astype(np.float64) * delta_f
            shft = LAL_TWOPI * (-1.0 / delta_f)   # epoch tC = -1/deltaF
            v = np.cbrt(piM * f)
            logv = np.log(v)
            v2 = v * v
            v3 = v * v2
            v4 = v * v3
            v5 = v * v4
            v6 = v * v5
            v7 = v * v6
            v10 = v5 * v5
            phasing = pv[7] * v7
            phasing = phasing + (pv[6] + pl[6] * logv) * v6
            phasing = phasing + (pv[5] + pl[5] * logv) * v5
            phasing = phasing + pv[4] * v4
            phasing = phasing + pv[3] * v3
            phasing = phasing + pv[2] * v2
            phasing = phasing + pv[1] * v
            phasing = phasing + pv[0]
            phasing = phasing + pft12 * (v10 * v2)      # tidal terms
            phasing = phasing + pft10 * v10
            phasing = phasing / v5
            flux = FTaN * v10
            dEnergy = dETaN * v
            phasing = phasing + (shft * f - 2.0 * coa_phase)
            amp = amp0 * np.sqrt(-dEnergy / flux) * v
            htilde[i_start:] = amp * np.cos(phasing - LAL_PI_4) - \
                1j * (amp * np.sin(phasing - LAL_PI_4))
        else:
            ld = np.longdouble
            f = k.astype(ld) * ld(delta_f)
            v = np.cbrt(ld(piM) * f)
            logv = np.log(v)
            poly = ld(pv[7])
            poly = poly * v + (ld(pv[6]) + ld(pl[6]) * logv)
            poly = poly * v + (ld(pv[5]) + ld(pl[5]) * logv)
            poly = poly * v + ld(pv[4])
            poly = poly * v + ld(pv[3])
            poly = poly * v + ld(pv[2])
            poly = poly * v + ld(pv[1])
            poly = poly * v + ld(pv[0])
            poly = poly + ld(pft10) * v ** 10 + ld(pft12) * v ** 12
            phasing = poly / v ** 5
            # shft*f = -2 pi k exactly: drop it
            phasing = phasing - ld(2.0) * ld(coa_phase) - ld(LAL_PI) / 4
            twopi = 2 * np.arctan2(ld(0), ld(-1))
            phasing = phasing - twopi * np.rint(phasing / twopi)
            amp = ld(amp0) * np.sqrt(ld(5.0) / (ld(32.0) * ld(eta))) * \
                v ** ld(-3.5)
            htilde = np.zeros(n, dtype=np.clongdouble)
            htilde[i_start:] = amp * np.cos(phasing) - \
                1j * (amp * np.sin(phasing))
    cfac = math.cos(inclination)
    pfac = 0.5 * (1.0 + cfac * cfac)
    hp = pfac * htilde
    hc = (-1j * cfac) * htilde
    return hp, hc


# ---------------------------------------------------------------------------
# [UPSTREAM] lal.GreenwichMeanSiderealTime
# ---------------------------------------------------------------------------
def gps_leap_seconds(gps_sec):
    """GPS-UTC in seconds at the given GPS second."""
    n = 0
    for t in LEAP_SECONDS_GPS:
        if gps_sec >= t:
            n += 1
    return n


def _civil_jd(gps_sec):
    """XLALGPSToUTC + XLALConvertCivilTimeToJD (integer seconds)."""
    unix = gps_sec - gps_leap_seconds(gps_sec) + 315964800
    import time as _time
    tm = _time.gmtime(unix)
    year, month, day = tm.tm_year, tm.tm_mon, tm.tm_mday
    sec = tm.tm_sec + 60 * (tm.tm_min + 60 * tm.tm_hour)
    jd = 367 * year - 7 * (year + (month + 9) // 12) // 4 + \
        275 * month // 9 + day + 1721014
    return jd + sec / 86400.0 - 0.5


def gmst(t_gps):
    """XLALGreenwichMeanSiderealTime of a float GPS time (radians, not
    reduced mod 2pi)."""
    sec = int(math.floor(t_gps))
    ns = int(round((t_gps - sec) * 1e9))
    if ns >= 1000000000:
        sec += 1
        ns -= 1000000000
    julian_day = _civil_jd(sec)
    t_hi = (julian_day - XLAL_EPOCH_J2000_0_JD) / 36525.0
    t_lo = ns / (1e9 * 36525.0 * 86400.0)
    t = t_hi + t_lo
    sidereal_time = (-6.2e-6 * t + 0.093104) * t * t + 67310.54841
    sidereal_time += 8640184.812866 * t_lo
    sidereal_time += 3155760000.0 * t_lo
    sidereal_time += 8640184.812866 * t_hi
    sidereal_time += 3155760000.0 * t_hi
    return sidereal_time * LAL_PI / 43200.0


# ---------------------------------------------------------------------------
# [UPSTREAM] pycbc.detector.Detector (LALDetectors.h cached detectors).
# LAL stores response = 0.5*(x (x) x - y (x) y) as REAL4.
# ---------------------------------------------------------------------------
_DETECTOR_GEOMETRY = {
    # vertex location [m], x-arm unit vector, y-arm unit vector (Earth-fixed)
    "H1": ((-2.16141492636e+06, -3.83469517889e+06, 4.60035022664e+06),
           (-0.22389266154, +0.79983062746, +0.55690487831),
           (-0.91397818574, +0.02609403989, -0.40492342125)),
    "L1": ((-7.42760447238e+04, -5.49628371971e+06, 3.22425701744e+06),
           (-0.95457412153, -0.14158077340, -0.26218911324),
           (+0.29774156894, -0.48791033647, -0.82054461286)),
    "V1": ((4.54637409900e+06, 8.42989697626e+05, 4.37857696241e+06),
           (-0.70045821479, +0.20848948619, +0.68256166277),
           (-0.05379255368, -0.96908180549, +0.24080451708)),
    # KAGRA (LAL_KAGRA_*), used for the 4-detector code path
    "K1": ((-3777336.024, 3484898.411, 3765313.697),
           (-0.3759040, -0.8361583, 0.3994189),
           (0.7164378, 0.01114076, 0.6975620)),
}


class Detector(object):
    def __init__(self, name):
        loc, xa, ya = _DETECTOR_GEOMETRY[name]
        self.name = name
        self.location = np.array(loc, dtype=np.float64)
        xa = np.array(xa)
        ya = np.array(ya)
        resp = 0.5 * (np.outer(xa, xa) - np.outer(ya, ya))
        self.response = resp.astype(np.float32).astype(np.float64)

    def antenna_pattern(self, ra, dec, pol, t_gps):
        gha = gmst(t_gps) - ra
        cosgha, singha = math.cos(gha), math.sin(gha)
        cosdec, sindec = math.cos(dec), math.sin(dec)
        cospsi, sinpsi = math.cos(pol), math.sin(pol)
        x = np.array([-cospsi * singha - sinpsi * cosgha * sindec,
                      -cospsi * cosgha + sinpsi * singha * sindec,
                      sinpsi * cosdec])
        dx = self.response.dot(x)
        y = np.array([sinpsi * singha - cospsi * cosgha * sindec,
                      sinpsi * cosgha + cospsi * singha * sindec,
                      cospsi * cosdec])
        dy = self.response.dot(y)
        fplus = (x * dx - y * dy).sum()
        fcross = (x * dy + y * dx).sum()
        return fplus, fcross

    def time_delay_from_earth_center(self, ra, dec, t_gps):
        ra_angle = gmst(t_gps) - ra
        cosd = math.cos(dec)
        ehat = np.array([cosd * math.cos(ra_angle),
                         cosd * -math.sin(ra_angle),
                         math.sin(dec)])
        dx = np.zeros(3) - self.location
        return float(dx.dot(ehat) / LAL_C_SI)


# ---------------------------------------------------------------------------
# [UPSTREAM] pycbc.waveform.utils.apply_fd_time_shift ->
# apply_fseries_time_shift (utils_cpu): phi = -2 pi dt df, phasor recurrence
# with exact cos/sin(phi*k) re-sync every 100 bins.
# ---------------------------------------------------------------------------
def apply_fseries_time_shift(h, dt, delta_f, mode="faithful"):
    n = len(h)
    if mode == "faithful":
        phi = -2.0 * np.pi * dt * delta_f
        cphi, sphi = math.cos(phi), math.sin(phi)
        # vectorised form of the recurrence: for bin k = 100*q + r the
        # reference holds exp(i phi 100 q) * (cphi + i sphi)^r built by r
        # successive complex multiplications.
        nblk = (n + 99) // 100
        q = np.arange(nblk, dtype=np.float64) * 100.0
        base = np.cos(phi * q) + 1j * np.sin(phi * q)
        ph = np.empty((nblk, 100), dtype=np.complex128)
        re = base.real.copy()
        im = base.imag.copy()
        for r in range(100):
            ph[:, r] = re + 1j * im
            re, im = re * cphi - im * sphi, re * sphi + im * cphi
        return h * ph.reshape(-1)[:n]
    ld = np.longdouble
    k = np.arange(n, dtype=ld)
    tau = ld(dt) * ld(delta_f)
    tau = tau - np.rint(tau)
    x = k * tau
    x = x - np.rint(x)
    twopi = 2 * np.arctan2(ld(0), ld(-1))
    ang = -twopi * x
    return h * (np.cos(ang) + 1j * np.sin(ang))


# ---------------------------------------------------------------------------
# [UPSTREAM] pycbc.waveform.generator.FDomainDetFrameGenerator, as constructed
# by the reference at base_data.py:224-230 and test/utils/waveform.py:64-65.
# ---------------------------------------------------------------------------
class CubicSpline(object):
    """gwin/calibration.py:23-78 (``Recalibrate.map_to_adjust``) and :104-173
    (``CubicSpline``): strain * (1 + dA(f)) * (2 + i dphi(f)) / (2 - i dphi(f)) with dA, dphi
    ``scipy.interpolate.UnivariateSpline`` fits (default k = 3 and smoothing s = n_points --
    which for calibration-sized values returns the least-squares cubic polynomial, not an
    interpolant) through ``n_points`` log-spaced frequencies, evaluated at every sample
    frequency of the waveform (extrapolating below/above the spline range)."""
    name = "cubic_spline"

    def __init__(self, minimum_frequency, maximum_frequency, n_points, ifo_name):
        self.ifo_name = ifo_name
        self.params = dict()
        n_points = int(n_points)
        if n_points < 4:
            raise ValueError("Use at least 4 spline points for calibration model")
        self.n_points = n_points
        self.spline_points = np.logspace(np.log10(float(minimum_frequency)),
                                         np.log10(float(maximum_frequency)), n_points)

    def map_to_adjust(self, strain, delta_f, prefix="recalib_", **params):
        self.params.update({key[len(prefix):]: params[key] for key in params
                            if prefix in key and self.ifo_name in key})
        return self.apply_calibration(strain, delta_f)

    def apply_calibration(self, strain, delta_f):
        from scipy.interpolate import UnivariateSpline
        freqs = np.arange(len(strain)) * delta_f          # strain.sample_frequencies
        amp = [self.params["amplitude_{}_{}".format(self.ifo_name, ii)]
               for ii in range(self.n_points)]
        delta_amplitude = UnivariateSpline(self.spline_points, amp)(freqs)
        pha = [self.params["phase_{}_{}".format(self.ifo_name, ii)]
               for ii in range(self.n_points)]
        delta_phase = UnivariateSpline(self.spline_points, pha)(freqs)
        return strain * (1.0 + delta_amplitude) \
            * (2.0 + 1j * delta_phase) / (2.0 - 1j * delta_phase)


class FDomainDetFrameGenerator(object):
    location_args = ("tc", "ra", "dec", "polarization")

    def __init__(self, epoch, detectors, delta_f, f_lower, variable_args=(),
                 mode="faithful", recalib=None, **static):
        # [UPSTREAM] pycbc generator: ``recalib`` = dict detector -> Recalibrate instance,
        # applied to each detector's shifted waveform (base_data.py:225-230 passes it in)
        self.recalib = recalib
        self.epoch = float(epoch)
        self.detectors = OrderedDict((d, Detector(d)) for d in detectors)
        self.detector_names = list(detectors)
        self.delta_f = float(delta_f)
        self.f_lower = float(f_lower)
        self.variable_args = tuple(variable_args)
        self.static = dict(static)
        self.mode = mode
        self.current_params = dict(static)

    def generate(self, **kwargs):
        self.current_params.update(kwargs)
        p = self.current_params
        hp, hc = taylorf2_hp_hc(
            p["mass1"], p["mass2"], p.get("spin1z", 0.0), p.get("spin2z", 0.0),
            p["distance"], p.get("inclination", 0.0), p.get("coa_phase", 0.0),
            self.delta_f, self.f_lower, mode=self.mode,
            lambda1=p.get("lambda1", 0.0), lambda2=p.get("lambda2", 0.0))
        h = OrderedDict()
        for detname, det in self.detectors.items():
            fp, fc = det.antenna_pattern(p["ra"], p["dec"], p["polarization"],
                                         p["tc"])
            thish = fp * hp + fc * hc
            tc = p["tc"] + det.time_delay_from_earth_center(
                p["ra"], p["dec"], p["tc"])
            dt = float(tc - self.epoch)
            h[detname] = apply_fseries_time_shift(thish, dt, self.delta_f,
                                                  mode=self.mode)
            if self.recalib:
                h[detname] = self.recalib[detname].map_to_adjust(
                    h[detname], self.delta_f, **self.current_params)
        return h


# ---------------------------------------------------------------------------
# The gwin models themselves.
# ---------------------------------------------------------------------------
def _inner(a, b, mode):
    """[UPSTREAM] pycbc Array.inner: sum(conj(a) * b) in complex128."""
    if mode == "faithful":
        return complex(np.vdot(a, b))
    t = np.conj(a.astype(np.clongdouble)) * b.astype(np.clongdouble)
    return complex(t.sum())


def log_i0(x):
    """log(I0(x)) without the overflow of numpy.log(special.i0(x)) used at
    marginalized_gaussian_noise.py:459 (identical where that is finite)."""
    from scipy import special
    return float(x + np.log(special.i0e(x)))


class GaussianNoiseOracle(object):
    """Follows gwin/models/gaussian_noise.py:219-350 (and, for
    ``phase_marginalization=True`` / ``distance_marginalization=(grid, weights)``,
    marginalized_gaussian_noise.py:431-511; time marginalisation is not restated).

    data : dict det -> complex128 array (frequency series, bin spacing delta_f)
    psds : dict det -> float64 array or None
    """

    def __init__(self, data, delta_f, epoch, f_lower, psds=None, f_upper=None,
                 norm=None, phase_marginalization=False, mode="faithful",
                 waveform_f_lower=None, distance_marginalization=None,
                 recalibration=None):
        self.mode = mode
        # marginalized_gaussian_noise.py:368-375: (dist_array, deltad * dist_prior)
        self.distance_marginalization = distance_marginalization
        self.dets = list(data.keys())
        # base_data.py:87 -- the model keeps a copy
        self._data = OrderedDict(
            (d, np.array(data[d], dtype=np.complex128)) for d in self.dets)
        dlens = np.array([len(d) for d in self._data.values()])
        if not all(dlens == dlens[0]):
            raise ValueError("all data must be of the same length")
        N = int(dlens[0])
        self.delta_f = float(delta_f)
        # gaussian_noise.py:247-250
        kmin, kmax = get_cutoff_indices(f_lower, f_upper, delta_f, (N - 1) * 2)
        self._kmin, self._kmax = kmin, kmax
        if norm is None:
            norm = 4 * delta_f
        # gaussian_noise.py:253-262
        if psds is None:
            w = np.sqrt(norm) * np.ones(N)
            self._weight = {d: w for d in self.dets}
        else:
            with np.errstate(divide="ignore"):
                self._weight = {d: np.sqrt(norm / np.asarray(psds[d],
                                                             dtype=np.float64))
                                for d in self.dets}
        # gaussian_noise.py:264-265
        for d in self.dets:
            self._data[d][kmin:kmax] *= self._weight[d][kmin:kmax]
        self.generator = FDomainDetFrameGenerator(
            epoch, self.dets, delta_f,
            f_lower if waveform_f_lower is None else waveform_f_lower,
            mode=mode, recalib=recalibration)
        self.phase_marginalization = phase_marginalization
        self._lognl = None

    # gaussian_noise.py:275-291
    def lognl(self):
        if self._lognl is None:
            self._det_lognls = {}
            for det, d in self._data.items():
                s = d[self._kmin:self._kmax]
                self._det_lognls[det] = -0.5 * _inner(s, s, self.mode).real
            self._lognl = sum(self._det_lognls.values())
        return self._lognl

    def det_lognl(self, det):
        self.lognl()
        return self._det_lognls[det]

    def loglr(self, **params):
        """Returns (loglr, {det: cplx <d|h>}, {det: <h|h>}).
        gaussian_noise.py:304-350 / marginalized_gaussian_noise.py:461-511."""
        wfs = self.generator.generate(**params)
        hd = OrderedDict()
        hh = OrderedDict()
        for det, h in wfs.items():
            kmax = min(len(h), self._kmax)
            if self._kmin >= kmax:
                hd[det] = 0j
                hh[det] = 0.
            else:
                slc = slice(self._kmin, kmax)
                hw = h[slc] * self._weight[det][slc]
                hd[det] = _inner(self._data[det][slc], hw, self.mode)
                hh[det] = _inner(hw, hw, self.mode).real
        if self.distance_marginalization is not None:
            # marginalized_gaussian_noise.py:431-448 (_margdistphase_loglr / _margdist_loglr),
            # with mf_snr = abs(sum hd) for both (:508) and scipy's logsumexp as there
            from scipy import special
            dist, b = self.distance_marginalization
            mf = abs(sum(hd.values()))
            x = log_i0(mf) if self.phase_marginalization else mf
            with np.errstate(divide="ignore"):
                lr = special.logsumexp(x / dist - 0.5 * sum(hh.values()) / dist ** 2, b=b)
        elif self.phase_marginalization:
            mf = abs(sum(hd.values()))
            lr = log_i0(mf) - 0.5 * sum(hh.values())
        else:
            lr = 0.
            for det in wfs:
                lr += (hd[det] - 0.5 * hh[det]).real
        return float(lr), hd, hh


PARAM_ORDER = ("mass1", "mass2", "spin1z", "spin2z", "distance", "inclination",
               "coa_phase", "ra", "dec", "polarization", "tc", "lambda1", "lambda2")


def batch_loglr(model, params):
    """Evaluate an oracle model on a [W, 13] array in PARAM_ORDER.
    Returns loglr[W], hd[W, ndet] complex, hh[W, ndet]."""
    params = np.atleast_2d(np.asarray(params, dtype=np.float64))
    W = params.shape[0]
    nd = len(model.dets)
    out = np.empty(W)
    hd = np.empty((W, nd), dtype=np.complex128)
    hh = np.empty((W, nd))
    for i in range(W):
        lr, d, h = model.loglr(**dict(zip(PARAM_ORDER, params[i])))
        out[i] = lr
        hd[i] = [d[k] for k in model.dets]
        hh[i] = [h[k] for k in model.dets]
    return out, hd, hh
