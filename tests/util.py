"""Shared synthetic configurations for the tests.

Data generation uses the ORACLE's waveform for the injection (tests may import
the oracle; the product may not), seeded numpy noise and the analytic ZDHP PSD.
"""
import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import gwin_oracle as O  # noqa: E402

PARAM_ORDER = O.PARAM_ORDER

BBH_INJ = OrderedDict([  # reference fixture values, test/utils/waveform.py:25-35
    ("mass1", 30.0), ("mass2", 30.0), ("spin1z", 0.0), ("spin2z", 0.0),
    ("distance", 300.0), ("inclination", 0.1), ("coa_phase", 1.1),
    ("ra", 0.1), ("dec", 0.1), ("polarization", 0.1), ("tc", 6.1),
    ("lambda1", 0.0), ("lambda2", 0.0)])

BNS_INJ = OrderedDict([
    ("mass1", 1.4), ("mass2", 1.4), ("spin1z", 0.0), ("spin2z", 0.0),
    ("distance", 40.0), ("inclination", 0.5), ("coa_phase", 1.1),
    ("ra", 1.3), ("dec", -0.4), ("polarization", 0.7), ("tc", 250.0),
    ("lambda1", 0.0), ("lambda2", 0.0)])


class Config(object):
    """One synthetic data set: detectors, PSD, data (noise + injection)."""

    def __init__(self, name, seglen, srate, f_lower, dets, inj, epoch=0.,
                 noise=True, seed=20181, psd=True):
        self.name = name
        self.seglen, self.srate, self.f_lower = seglen, srate, f_lower
        self.dets = list(dets)
        self.epoch = float(epoch)
        self.inj = OrderedDict(inj)
        self.inj["tc"] = self.inj["tc"] + self.epoch
        N = int(seglen * srate)
        self.n_f = N // 2 + 1
        self.delta_f = 1.0 / seglen
        self.psd = O.aligo_zdhp_psd(self.n_f, self.delta_f, f_lower) if psd else None
        gen = O.FDomainDetFrameGenerator(self.epoch, self.dets, self.delta_f, f_lower)
        wf = gen.generate(**self.inj)
        rng = np.random.default_rng(seed)
        self.data = OrderedDict()
        for d in self.dets:
            x = np.zeros(self.n_f, dtype=np.complex128)
            if noise:
                ref = self.psd if psd else np.ones(self.n_f)
                x = np.sqrt(ref / (4 * self.delta_f)) * (
                    rng.standard_normal(self.n_f) + 1j * rng.standard_normal(self.n_f))
            n = min(len(wf[d]), self.n_f)
            x[:n] += wf[d][:n]
            self.data[d] = x

    def psds(self):
        return None if self.psd is None else {d: self.psd for d in self.dets}

    def oracle(self, mode="faithful", phase_marginalization=False, **kw):
        return O.GaussianNoiseOracle(
            self.data, self.delta_f, self.epoch, self.f_lower, psds=self.psds(),
            phase_marginalization=phase_marginalization, mode=mode, **kw)

    def cuda_model(self, cls=None, variable_params=PARAM_ORDER, **kw):
        """The product model on the GPU (imports gwin_b200 lazily)."""
        from gwin_b200 import models
        from gwin_b200.types import FrequencySeries
        from gwin_b200.waveform import (FDomainCBCGenerator,
                                        FDomainDetFrameGenerator)
        cls = cls or models.GaussianNoise
        gen = FDomainDetFrameGenerator(
            FDomainCBCGenerator, self.epoch, detectors=self.dets,
            variable_args=variable_params, delta_f=self.delta_f,
            f_lower=self.f_lower, approximant="TaylorF2")
        data = {d: FrequencySeries(self.data[d], self.delta_f, self.epoch)
                for d in self.dets}
        return cls(variable_params, data, gen, self.f_lower,
                   psds=self.psds(), **kw)


def c1(**kw):
    """C1/C2 data: 8 s @ 2048 Hz, f_lower 20, BBH injection."""
    kw.setdefault("dets", ["H1", "L1"])
    return Config("C1", 8, 2048., 20., inj=BBH_INJ, **kw)


def c2(**kw):
    kw.setdefault("dets", ["H1", "L1", "V1"])
    return Config("C2", 8, 2048., 20., inj=BBH_INJ, **kw)


def c3(seglen=256, **kw):
    """C3: 256 s @ 4096 Hz BNS in coloured Gaussian noise, H1+L1+V1."""
    kw.setdefault("dets", ["H1", "L1", "V1"])
    inj = OrderedDict(BNS_INJ)
    inj["tc"] = seglen - 6.0
    return Config("C3", seglen, 4096., 20., inj=inj, **kw)


def draw_bbh(rng, W, cfg, spins=True):
    """Walkers around the test-suite priors (test/test_sampler.py:74-98)."""
    P = np.zeros((W, 13))
    P[:, 0] = rng.uniform(10., 40., W)
    P[:, 1] = rng.uniform(10., 40., W)
    P[:, 2] = rng.uniform(-0.5, 0.5, W) if spins else 0.
    P[:, 3] = rng.uniform(-0.5, 0.5, W) if spins else 0.
    P[:, 4] = rng.uniform(100., 1000., W)
    P[:, 5] = np.arccos(rng.uniform(-1, 1, W))
    P[:, 6] = rng.uniform(0, 2 * np.pi, W)
    P[:, 7] = rng.uniform(0, 2 * np.pi, W)
    P[:, 8] = np.arcsin(rng.uniform(-1, 1, W))
    P[:, 9] = rng.uniform(0, 2 * np.pi, W)
    P[:, 10] = cfg.inj["tc"] + rng.uniform(-0.1, 0.1, W)
    return P


def draw_bns(rng, W, cfg):
    """Box around the BNS injection."""
    P = np.zeros((W, 13))
    P[:, 0] = rng.uniform(1.2, 1.6, W)
    P[:, 1] = rng.uniform(1.2, 1.6, W)
    P[:, 2] = 0.
    P[:, 3] = 0.
    P[:, 4] = rng.uniform(10., 100., W)
    P[:, 5] = np.arccos(rng.uniform(-1, 1, W))
    P[:, 6] = rng.uniform(0, 2 * np.pi, W)
    P[:, 7] = rng.uniform(0, 2 * np.pi, W)
    P[:, 8] = np.arcsin(rng.uniform(-1, 1, W))
    P[:, 9] = rng.uniform(0, 2 * np.pi, W)
    P[:, 10] = cfg.inj["tc"] + rng.uniform(-0.05, 0.05, W)
    return P


def inj_row(cfg):
    return np.array([cfg.inj[k] for k in PARAM_ORDER])


def tol(loglr):
    """north_star tolerance: 1e-6 absolute, 1e-9 relative for loglr >> 1."""
    return np.maximum(1e-6, 1e-9 * np.abs(loglr))


def dist_grid(lo=50., hi=5000., n=10 ** 4):
    """marginalized_gaussian_noise.py:368-375 with Uniform(distance=(lo, hi)): the grid and
    ``deltad * dist_prior`` (0 on the last point: the prior's upper bound is open)."""
    dist = np.linspace(lo, hi, n)
    return dist, (dist[1] - dist[0]) * np.where(dist < hi, 1. / (hi - lo), 0.)


def calib_names(dets, n_points):
    """``recalib_*`` parameter names in kernel order (calibration.py:151-161)."""
    return ["recalib_%s_%s_%d" % (which, d.lower(), i) for d in dets
            for which in ("amplitude", "phase") for i in range(n_points)]


def golden_config(name):
    """The synthetic-data configuration behind tests/golden/<name> (make_golden.py)."""
    return {"c1_gaussian_h1l1.npz": lambda: c1(),
            "c2_margphase_h1l1v1.npz": lambda: c2(),
            "c3_bns256_h1l1v1.npz": lambda: c3(),
            "c2_gps_epoch.npz": lambda: c2(epoch=1126259462.0, seed=7),
            "c3s32_tidal.npz": lambda: c3(seglen=32),
            "c2_margdist.npz": lambda: c2(),
            "c2_margdistphase.npz": lambda: c2(),
            "c2_calibrated.npz": lambda: c2()}[name]()


def autocorrelation_length(chain):
    """Integrated autocorrelation time [iterations] of an ensemble chain [nwalkers, niter]: the
    walker-averaged autocorrelation function summed up to Sokal's window M = 5 tau (what
    pycbc.filter.autocorrelation.calculate_acl, used by gwin/burn_in.py and the samplers' ACL
    thinning, estimates)."""
    x = np.asarray(chain, dtype=float)
    x = x - x.mean(axis=1, keepdims=True)
    n = x.shape[1]
    f = np.fft.rfft(x, 2 * n, axis=1)
    acf = np.fft.irfft(f * np.conj(f), axis=1)[:, :n].mean(axis=0)
    if acf[0] <= 0:
        return 1.0
    rho = acf / acf[0]
    tau = 1.0
    for m in range(1, n):
        tau += 2.0 * rho[m]
        if m >= 5.0 * tau:
            break
    return max(tau, 1.0)
