"""CUDA path vs the CPU oracle, through the C ABI (``-m gpu``).

Tolerance (BASELINE.json north_star): |d loglr| <= 1e-6 absolute, 1e-9 relative
for loglr >> 1.  Both kernels (direct per-bin sincos, block recurrence) are held
to it, and to each other at sizes the oracle cannot reach.
"""
import numpy as np
import pytest

import util
from util import PARAM_ORDER, tol

pytestmark = pytest.mark.gpu

KERNELS = {"direct": 1, "recurrence": 2, "table": 3}


def _oracle_batch(orc, P):
    import gwin_oracle as O
    return O.batch_loglr(orc, P)


def _check(lr, hd, hh, lr0, hd0, hh0, what=""):
    assert np.all(np.isfinite(lr)), what
    err = np.abs(lr - lr0)
    assert np.all(err <= tol(lr0)), "%s: max |dloglr| %.3e at %d (loglr %.6f)" % (
        what, err.max(), err.argmax(), lr0[err.argmax()])
    # the per-detector stats obey the same bound
    assert np.all(np.abs(hh - hh0) <= np.maximum(1e-6, 1e-9 * np.abs(hh0))), what
    assert np.all(np.abs(hd - hd0) <= np.maximum(2e-6, 2e-9 * np.abs(hh0))), what


@pytest.fixture(scope="module")
def cfg1():
    return util.c1()


@pytest.fixture(scope="module")
def cfg2():
    return util.c2()


@pytest.fixture(scope="module")
def cfg3():
    return util.c3()


@pytest.mark.parametrize("kernel", sorted(KERNELS))
def test_c1_gaussian_200_walkers(cfg1, kernel):
    """C1: GaussianNoise loglr, H1+L1, 8 s @ 2048 Hz, 200 emcee walkers."""
    model = cfg1.cuda_model()
    model.plan.configure(kernel=KERNELS[kernel])
    rng = np.random.default_rng(20182)
    P = util.draw_bbh(rng, 200, cfg1)
    P[0] = util.inj_row(cfg1)
    lr, hd, hh = model.plan.loglr_host(P)
    lr0, hd0, hh0 = _oracle_batch(cfg1.oracle(), P)
    _check(lr, hd, hh, lr0, hd0, hh0, "C1 " + kernel)
    assert abs(model.lognl - cfg1.oracle().lognl()) <= 1e-9 * abs(model.lognl)


@pytest.mark.parametrize("kernel", sorted(KERNELS))
def test_c2_marginalized_phase(cfg2, kernel):
    """C2: phase-marginalised model, H1+L1+V1."""
    from gwin_b200 import models
    model = cfg2.cuda_model(cls=models.MarginalizedPhaseGaussianNoise)
    model.plan.configure(kernel=KERNELS[kernel])
    rng = np.random.default_rng(20183)
    P = util.draw_bbh(rng, 256, cfg2)
    P[0] = util.inj_row(cfg2)
    lr, hd, hh = model.plan.loglr_host(P, mode=1)
    lr0, hd0, hh0 = _oracle_batch(cfg2.oracle(phase_marginalization=True), P)
    _check(lr, hd, hh, lr0, hd0, hh0, "C2 " + kernel)


@pytest.mark.parametrize("kernel", sorted(KERNELS))
def test_c3_bns_256s(cfg3, kernel):
    """C3: BNS, 256 s @ 4096 Hz (524 289 bins), coloured noise, 3 detectors."""
    model = cfg3.cuda_model()
    model.plan.configure(kernel=KERNELS[kernel])
    rng = np.random.default_rng(20184)
    P = util.draw_bns(rng, 24, cfg3)
    P[0] = util.inj_row(cfg3)
    lr, hd, hh = model.plan.loglr_host(P)
    lr0, hd0, hh0 = _oracle_batch(cfg3.oracle(), P)
    _check(lr, hd, hh, lr0, hd0, hh0, "C3 " + kernel)
    # and against the extended-precision oracle
    lrT, _, _ = _oracle_batch(cfg3.oracle(mode="truth"), P[:6])
    assert np.all(np.abs(lr[:6] - lrT) <= tol(lrT))


def test_c3_kernels_agree_full_batch(cfg3):
    """At BASELINE size (thousands of walkers x 524 289 bins) the oracle is too
    slow; the two independent kernels must agree with each other instead."""
    model = cfg3.cuda_model()
    rng = np.random.default_rng(20185)
    P = util.draw_bns(rng, 4096, cfg3)
    model.plan.configure(kernel=1)
    a = model.plan.loglr_host(P)
    perm = rng.permutation(len(P))
    for kernel in (2, 3):                           # recurrence; recurrence + moment tables
        model.plan.configure(kernel=kernel)
        b = model.plan.loglr_host(P)
        assert np.all(np.abs(a[0] - b[0]) <= tol(a[0])), kernel
        assert np.all(np.abs(a[1] - b[1]) <= np.maximum(2e-6, 2e-9 * np.abs(a[2]))), kernel
        np.testing.assert_array_equal(a[2], b[2])       # <h|h> is the same closed form
        # determinism: same call, same bits
        c = model.plan.loglr_host(P)
        for x, y in zip(b, c):
            np.testing.assert_array_equal(x, y)
        # order invariance: a permuted batch gives permuted, bit-identical results
        d = model.plan.loglr_host(P[perm])
        np.testing.assert_array_equal(d[0], b[0][perm])
    # the table kernel follows the batch's chirp-mass range: a batch of lighter systems
    # rebuilds the schedule and the tables, and still agrees with the direct kernel
    Q = P[:64].copy()
    Q[:, 0] = rng.uniform(0.6, 0.8, 64)
    Q[:, 1] = rng.uniform(0.6, 0.8, 64)
    model.plan.configure(kernel=1)
    a = model.plan.loglr_host(Q)
    model.plan.configure(kernel=3)
    b = model.plan.loglr_host(Q)
    assert np.all(np.abs(a[0] - b[0]) <= tol(a[0]))


def test_gps_epoch(cfg1):
    """GPS-scale epoch: leap seconds in GMST and the reference's
    ``tc + time_delay`` rounding at ulp(1e9 s) are reproduced."""
    cfg = util.c2(epoch=1126259462.0, seed=7)
    model = cfg.cuda_model()
    rng = np.random.default_rng(5)
    P = util.draw_bbh(rng, 64, cfg)
    for kernel in (1, 2, 3, 0):
        model.plan.configure(kernel=kernel)
        lr, hd, hh = model.plan.loglr_host(P)
        lr0, hd0, hh0 = _oracle_batch(cfg.oracle(), P)
        _check(lr, hd, hh, lr0, hd0, hh0, "gps epoch k%d" % kernel)


def test_zero_noise_identity():
    """loglr(theta_inj) = 1/2 sum <h|h> in zero noise (cf. the reference
    docstring gaussian_noise.py:158-164)."""
    cfg = util.c2(noise=False)
    model = cfg.cuda_model()
    lr, hd, hh = model.plan.loglr_host(util.inj_row(cfg)[None, :])
    assert abs(lr[0] - 0.5 * hh[0].sum()) <= 1e-9 * lr[0]
    assert np.all(np.abs(hd[0].imag) <= 1e-6 * hh[0])


def test_reference_tc_grid_argmax():
    """The reference's only numeric pin (test/test_models.py:105-117): zero-noise
    data = injected waveform, 4 s @ 2048 Hz, fmin 30, H1/L1/V1; logplr over the
    grid tc = 1.1 + arange(8192)/2048 peaks at tc = 3.1."""
    import gwin_oracle as O
    from gwin_b200 import models
    inj = dict(util.BBH_INJ)
    inj["tc"] = 3.1
    cfg = util.Config("ref", 4, 2048., 30., ["H1", "L1", "V1"], inj, noise=False)
    model = cfg.cuda_model(variable_params=("tc",),
                           static_params={k: v for k, v in inj.items() if k != "tc"})
    times = 1.1 + np.arange(8192) / 2048.
    call = models.CallModel(model, "logplr", return_all_stats=False)
    vals = call.batch(times[:, None]).values
    assert np.isclose(times[np.argmax(vals)], 3.1)
    call.callstat = "logposterior"
    vals = call.batch(times[:, None]).values
    assert np.isclose(times[np.argmax(vals)], 3.1)
    # scalar protocol, as the reference test calls it
    assert np.isclose(call([3.1]), vals[np.argmin(np.abs(times - 3.1))])


def test_edge_cases(cfg2):
    model = cfg2.cuda_model()
    orc = cfg2.oracle()
    base = util.inj_row(cfg2)
    P = np.tile(base, (12, 1))
    P[1, 0] = P[1, 1] = 150.    # f_ISCO = 14.7 Hz < f_lower: NoWaveformError -> -inf
    P[2, 0] = P[2, 1] = 100.    # f_ISCO ~ 22 Hz: a handful of bins
    P[3, 0] = np.nan
    P[4, 4] = -1.               # negative distance
    P[5, 1] = 0.
    P[6, 10] = np.inf
    P[7, 0], P[7, 1] = 1.0, 1.0  # BNS-like in an 8 s segment: clipped at kmax
    skip = np.zeros(12, dtype=np.uint8)
    skip[8] = 1
    for kernel in (1, 2, 3, 0):
        model.plan.configure(kernel=kernel)
        lr, hd, hh = model.plan.loglr_host(P, skip=skip)
        assert not np.any(np.isnan(lr))
        for i in (1, 3, 4, 5, 6, 8):
            assert lr[i] == -np.inf and np.all(hh[i] == 0) and np.all(hd[i] == 0)
        for i in (0, 2, 7, 9, 10, 11):
            l0, d0, h0 = orc.loglr(**dict(zip(PARAM_ORDER, P[i])))
            assert abs(lr[i] - l0) <= tol(l0), (kernel, i, lr[i], l0)
            assert np.allclose(hh[i], list(h0.values()), rtol=1e-9, atol=1e-9)
    # W = 0 and W = 1
    lr, hd, hh = model.plan.loglr_host(np.empty((0, 13)))
    assert lr.shape == (0,)
    lr1, _, _ = model.plan.loglr_host(P[:1])
    assert abs(lr1[0] - orc.loglr(**dict(zip(PARAM_ORDER, P[0])))[0]) <= tol(lr1[0])


def test_waveform_below_kmin():
    """Model f_lower above the end of the waveform: every detector contributes
    0 and loglr = 0 (gaussian_noise.py:328-333)."""
    import gwin_oracle as O
    cfg = util.c1()
    from gwin_b200 import models
    from gwin_b200.types import FrequencySeries
    from gwin_b200.waveform import FDomainCBCGenerator, FDomainDetFrameGenerator
    gen = FDomainDetFrameGenerator(FDomainCBCGenerator, cfg.epoch, detectors=cfg.dets,
                                   variable_args=PARAM_ORDER, delta_f=cfg.delta_f,
                                   f_lower=20., approximant="TaylorF2")
    data = {d: FrequencySeries(cfg.data[d], cfg.delta_f, cfg.epoch) for d in cfg.dets}
    model = models.GaussianNoise(PARAM_ORDER, data, gen, 100., psds=cfg.psds())
    row = util.inj_row(cfg)        # 30+30: f_ISCO = 73 Hz < 100 Hz
    lr, hd, hh = model.plan.loglr_host(row[None, :])
    assert lr[0] == 0.0 and np.all(hh == 0) and np.all(hd == 0)
    orc = O.GaussianNoiseOracle(cfg.data, cfg.delta_f, cfg.epoch, 100., psds=cfg.psds(),
                                waveform_f_lower=20.)
    assert orc.loglr(**dict(zip(PARAM_ORDER, row)))[0] == 0.0


@pytest.mark.parametrize("dets", [["H1"], ["H1", "L1"], ["L1", "V1", "H1"]])
def test_detector_counts_unit_psd_fupper(dets):
    """1/2/3 detectors, psds=None (unit weight, gaussian_noise.py:254-256) and a
    finite f_upper."""
    import gwin_oracle as O
    from gwin_b200 import models
    from gwin_b200.types import FrequencySeries
    from gwin_b200.waveform import FDomainCBCGenerator, FDomainDetFrameGenerator
    cfg = util.Config("u", 8, 2048., 20., dets, util.BBH_INJ, psd=False, seed=3)
    # scale the white noise so that loglr is O(100)
    scale = 1e-22
    for d in dets:
        cfg.data[d] *= scale
    gen = FDomainDetFrameGenerator(FDomainCBCGenerator, cfg.epoch, detectors=dets,
                                   variable_args=PARAM_ORDER, delta_f=cfg.delta_f,
                                   f_lower=20., approximant="TaylorF2")
    data = {d: FrequencySeries(cfg.data[d], cfg.delta_f, cfg.epoch) for d in dets}
    norm = 4 * cfg.delta_f / scale ** 2
    model = models.GaussianNoise(PARAM_ORDER, data, gen, 20., f_upper=60., norm=norm)
    orc = O.GaussianNoiseOracle(cfg.data, cfg.delta_f, cfg.epoch, 20., f_upper=60., norm=norm)
    assert (model._kmin, model._kmax) == (orc._kmin, orc._kmax)
    rng = np.random.default_rng(11)
    P = util.draw_bbh(rng, 40, cfg)
    for kernel in (1, 2, 3, 0):
        model.plan.configure(kernel=kernel)
        lr, hd, hh = model.plan.loglr_host(P)
        lr0, hd0, hh0 = O.batch_loglr(orc, P)
        _check(lr, hd, hh, lr0, hd0, hh0, "unit psd %s k%d" % (dets, kernel))
    assert abs(model.lognl - orc.lognl()) <= 1e-9 * abs(orc.lognl())


@pytest.mark.parametrize("seglen", [4, 16, 64, 128])
def test_segment_lengths_bns(seglen):
    """C5 regimes: short segments put the whole band (4 s) or its low-frequency part (16, 64 s)
    into directly-evaluated / short quartic blocks; every schedule regime must hold the
    tolerance."""
    import gwin_oracle as O
    cfg = util.c3(seglen=seglen)
    model = cfg.cuda_model()
    rng = np.random.default_rng(100 + seglen)
    P = util.draw_bns(rng, 12, cfg)
    P[0] = util.inj_row(cfg)
    lr0, hd0, hh0 = O.batch_loglr(cfg.oracle(), P)
    for kernel in (1, 2, 3):
        model.plan.configure(kernel=kernel)
        lr, hd, hh = model.plan.loglr_host(P)
        _check(lr, hd, hh, lr0, hd0, hh0, "T=%d k%d" % (seglen, kernel))


def test_large_segment_two_million_bins():
    """Size limit: 512 s @ 8192 Hz -> 2 097 153 bins per detector (bin index needs 22 bits: the
    exact k*tau_hi reduction, the 23-bit sort key and the tile stream all run at their limits)."""
    import gwin_oracle as O
    inj = dict(util.BNS_INJ)
    inj["tc"] = 500.0
    cfg = util.Config("big", 512, 8192., 20., ["H1", "L1"], inj, seed=11)
    assert cfg.n_f == 2097153
    model = cfg.cuda_model()
    rng = np.random.default_rng(12)
    P = util.draw_bns(rng, 3, cfg)
    P[0] = util.inj_row(cfg)
    P[1, :2] = 1.0, 1.0            # waveform reaches 2199 Hz = bin 1.1 million
    lr0, hd0, hh0 = O.batch_loglr(cfg.oracle(), P)
    for kernel in (1, 2, 3):
        model.plan.configure(kernel=kernel)
        lr, hd, hh = model.plan.loglr_host(P)
        _check(lr, hd, hh, lr0, hd0, hh0, "2M bins k%d" % kernel)


def test_tidal_deformabilities(cfg3):
    """lambda1, lambda2 (columns 11, 12 of the parameter row): both kernels against the oracle;
    lambda = 0 reproduces the point-particle result bit for bit."""
    model = cfg3.cuda_model()
    rng = np.random.default_rng(31)
    P = util.draw_bns(rng, 16, cfg3)
    base = model.plan.loglr_host(P)
    P2 = P.copy()
    P2[:, 11] = rng.uniform(0., 3000., len(P))
    P2[:, 12] = rng.uniform(0., 3000., len(P))
    lr0, hd0, hh0 = _oracle_batch(cfg3.oracle(), P2)
    for kernel in (1, 2, 3, 0):
        model.plan.configure(kernel=kernel)
        lr, hd, hh = model.plan.loglr_host(P2)
        _check(lr, hd, hh, lr0, hd0, hh0, "tidal k%d" % kernel)
    assert np.max(np.abs(lr0 - base[0])) > 1.0
    # lambda = 0 is the point-particle likelihood (the tables have meanwhile been rebuilt for the tidal
    # batch's coverage: same numbers to rounding, not bit for bit)
    model.plan.configure(kernel=0)
    again = model.plan.loglr_host(P)[0]
    assert np.all(np.abs(again - base[0]) <= 1e-10 * np.maximum(1.0, np.abs(base[0])))
    model.plan.configure(kernel=2)
    b2 = model.plan.loglr_host(P)[0]
    Pz = P.copy()
    Pz[:, 11:13] = 0.0
    np.testing.assert_array_equal(model.plan.loglr_host(Pz)[0], b2)
    # through the model API with lambda as a sampled parameter
    from gwin_b200 import models
    static = {k: v for k, v in cfg3.inj.items() if k not in ("lambda1", "lambda2")}
    m = cfg3.cuda_model(variable_params=("lambda1", "lambda2"), static_params=static)
    vals = models.CallModel(m, "loglr", return_all_stats=False).batch(np.array([[0., 0.], [400., 900.]])).values
    ref = [cfg3.oracle().loglr(**dict(cfg3.inj, lambda1=a, lambda2=b))[0] for a, b in ((0., 0.), (400., 900.))]
    assert np.all(np.abs(vals - ref) <= tol(np.array(ref)))


import glob
import os


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                               "golden", "*.npz"))))
def test_golden_fixtures(path):
    """The CUDA path against the COMMITTED golden vectors (tests/golden/make_golden.py): no
    oracle computation at run time.  Held to the float64-faithful numbers and to the
    extended-precision 'truth' ones."""
    g = np.load(path)
    cfg = util.golden_config(os.path.basename(path))
    model = cfg.cuda_model()
    marg = bool(g["marg"])
    mode = int(marg)
    calib = None
    if "dist_marg" in g:
        model.plan.set_distance_marg(*util.dist_grid(*g["dist_marg"]))
        mode = 3 if marg else 2
    if "calib_spec" in g:
        fmin, fmax, n = g["calib_spec"]
        sf = np.logspace(np.log10(fmin), np.log10(fmax), int(n))
        model.plan.set_calibration(np.tile(sf, (len(cfg.dets), 1)))
        calib = g["calib_values"]
    assert (model._kmin, model._kmax) == (int(g["kmin"]), int(g["kmax"]))
    assert abs(model.lognl - float(g["lognl"])) <= 1e-9 * abs(float(g["lognl"]))
    for kernel in (1, 2, 3, 0):
        model.plan.configure(kernel=kernel)
        lr, hd, hh = model.plan.loglr_host(g["params"], mode=mode, calib=calib)
        _check(lr, hd, hh, g["loglr"], g["hd"], g["hh"], "%s k%d" % (os.path.basename(path), kernel))
        assert np.all(np.abs(lr - g["loglr_truth"]) <= tol(g["loglr_truth"]))


def _dist_grid(lo=50., hi=5000., n=10 ** 4):
    """marginalized_gaussian_noise.py:368-375 with Uniform(distance=(lo, hi)): the prior is 0
    at the closed-open upper bound, i.e. on the last grid point."""
    dist = np.linspace(lo, hi, n)
    b = (dist[1] - dist[0]) * np.where(dist < hi, 1. / (hi - lo), 0.)
    return dist, b


@pytest.mark.parametrize("phase", [False, True])
def test_distance_marginalization(cfg2, phase):
    """GWIN_MODE_MARG_DIST / _DIST_PHASE against the oracle's scipy logsumexp, for templates
    at 1 Mpc (what the reference's x / D_j scaling means) and at their drawn distances (what
    the reference's own test does, test/test_models.py:122-145)."""
    model = cfg2.cuda_model()
    dist, b = _dist_grid()
    model.plan.set_distance_marg(dist, b)
    orc = cfg2.oracle(phase_marginalization=phase, distance_marginalization=(dist, b))
    rng = np.random.default_rng(77)
    P = util.draw_bbh(rng, 12, cfg2)
    P1 = P.copy()
    P1[:, 4] = 1.0
    P1[:3, 4] = 0.02                    # absurdly loud: the sum is dominated by the far end
    mode = 3 if phase else 2
    for what, Q in (("drawn", P), ("1 Mpc", P1)):
        lr0, hd0, hh0 = _oracle_batch(orc, Q)
        for kernel in (1, 2, 3, 0):
            model.plan.configure(kernel=kernel)
            lr, hd, hh = model.plan.loglr_host(Q, mode=mode)
            _check(lr, hd, hh, lr0, hd0, hh0, "margdist %s k%d" % (what, kernel))
    model.plan.configure(kernel=0)
    assert np.ptp(lr0) > 10.
    # skipped / invalid walkers stay -inf, no NaN
    Q = P1.copy()
    Q[1, 0] = np.nan
    skip = np.zeros(len(Q), np.uint8)
    skip[2] = 1
    lr = model.plan.loglr_host(Q, mode=mode, skip=skip)[0]
    assert lr[1] == -np.inf and lr[2] == -np.inf and np.all(np.isfinite(np.delete(lr, [1, 2])))
    # the grid is required, and clearing it brings the error back
    model.plan.set_distance_marg(np.empty(0), np.empty(0))
    with pytest.raises(ValueError):
        model.plan.loglr_host(P, mode=mode)


def test_reference_tc_grid_argmax_distance_marginalized():
    """The reference's second numeric pin (test/test_models.py:122-145 inheriting :105-117):
    MarginalizedGaussianNoise(distance_marginalization=True, Uniform(distance=(50, 5000)))
    peaks at tc = 3.1 on the same grid."""
    from gwin_b200 import distributions as D
    from gwin_b200 import models
    inj = dict(util.BBH_INJ)
    inj["tc"] = 3.1
    cfg = util.Config("ref", 4, 2048., 30., ["H1", "L1", "V1"], inj, noise=False)
    model = cfg.cuda_model(cls=models.MarginalizedGaussianNoise, variable_params=("tc",),
                           static_params={k: v for k, v in inj.items() if k != "tc"},
                           distance_marginalization=True,
                           marg_prior=[D.Uniform(distance=(50, 5000))])
    times = 1.1 + np.arange(8192) / 2048.
    call = models.CallModel(model, "logplr", return_all_stats=False)
    for callstat in ("logplr", "logposterior"):
        call.callstat = callstat
        vals = call.batch(times[:, None]).values
        assert np.isclose(times[np.argmax(vals)], 3.1)
    dist, b = _dist_grid()
    assert np.array_equal(model._dist_array, dist) and np.allclose(model._deltad * model.dist_prior, b, rtol=1e-15)
    orc = cfg.oracle(distance_marginalization=(dist, b))
    call.callstat = "loglr"
    for t in (3.1, 2.4):
        lr0 = orc.loglr(**dict(inj, tc=t))[0]
        assert abs(call([t]) - lr0) <= tol(lr0)


def _calib_setup(cfg, n_points, fmax, rng, W, scale=0.05):
    import gwin_oracle as O
    rec = {d: O.CubicSpline(cfg.f_lower, fmax, n_points, d.lower()) for d in cfg.dets}
    vals = rng.normal(0., scale, (W, len(cfg.dets) * 2 * n_points))
    names = []
    for d in cfg.dets:
        for which in ("amplitude", "phase"):
            names += ["recalib_%s_%s_%d" % (which, d.lower(), i) for i in range(n_points)]
    return rec, vals, names


def _calib_oracle_batch(orc, P, vals, names):
    nd = len(orc.dets)
    lr = np.empty(len(P))
    hd = np.empty((len(P), nd), complex)
    hh = np.empty((len(P), nd))
    for i, row in enumerate(P):
        kw = dict(zip(PARAM_ORDER, row))
        kw.update(zip(names, vals[i]))
        l, d, h = orc.loglr(**kw)
        lr[i], hd[i], hh[i] = l, [d[k] for k in orc.dets], [h[k] for k in orc.dets]
    return lr, hd, hh


@pytest.mark.parametrize("n_points,fmax", [(4, 1024.), (6, 512.), (10, 1024.)])
def test_calibration_cubic_spline(cfg2, n_points, fmax):
    """gwin/calibration.py:142-173 applied inside the kernels against the oracle's scipy
    UnivariateSpline; spline ranges narrower than the band exercise the extrapolation."""
    rng = np.random.default_rng(100 + n_points)
    W = 10
    rec, vals, names = _calib_setup(cfg2, n_points, fmax, rng, W)
    vals[0] = 0.0                               # zero calibration == the plain likelihood
    vals[1] *= 6.0                              # large (but still < 1) excursions
    vals[2, :n_points] = -0.9                   # 1 + dA = 0.1: the recurrence kernel falls back to
    vals[4, :n_points] = np.linspace(-0.99, 0.5, n_points)   # bin-by-bin evaluation where the log
    vals[5, n_points:2 * n_points] = np.linspace(-0.99, 0.99, n_points)   # is not smooth; big phases
    P = util.draw_bbh(rng, W, cfg2)
    model = cfg2.cuda_model()
    plain = model.plan.loglr_host(P)
    model.plan.set_calibration([rec[d].spline_points for d in cfg2.dets])
    for marg in (False, True):
        orc = cfg2.oracle(phase_marginalization=marg, recalibration=rec)
        lr0, hd0, hh0 = _calib_oracle_batch(orc, P, vals, names)
        for kernel in (1, 2, 3, 0):
            model.plan.configure(kernel=kernel)
            lr, hd, hh = model.plan.loglr_host(P, mode=int(marg), calib=vals)
            _check(lr, hd, hh, lr0, hd0, hh0, "calib n=%d marg=%d k%d" % (n_points, marg, kernel))
            # parameter-major layout gives the same numbers
            lr2 = model.plan.loglr_host(np.ascontiguousarray(P.T), mode=int(marg), param_major=True,
                                        calib=np.ascontiguousarray(vals.T))[0]
            np.testing.assert_array_equal(lr, lr2)
    model.plan.configure(kernel=0)
    import torch
    lr3 = model.plan.loglr_device(torch.tensor(P, device="cuda"), calib=torch.tensor(vals, device="cuda"))[0]
    np.testing.assert_array_equal(lr3.cpu().numpy(), model.plan.loglr_host(P, calib=vals)[0])
    assert abs(model.plan.loglr_host(P, calib=vals)[0][0] - plain[0][0]) <= tol(plain[0][0])
    assert np.max(np.abs(lr0[1:] - plain[0][1:])) > 0.1
    # non-finite value -> -inf; uncalibrated call on a calibrated plan is untouched
    bad = vals.copy()
    bad[3, 5] = np.nan
    assert model.plan.loglr_host(P, calib=bad)[0][3] == -np.inf
    np.testing.assert_array_equal(model.plan.loglr_host(P)[0], plain[0])
    model.plan.set_calibration(None)
    with pytest.raises(ValueError):
        model.plan.loglr_host(P, calib=vals)


def test_calibration_bns_long_segment():
    """Calibrated C3-like case (32 s so the oracle stays fast): both kernels."""
    cfg = util.c3(seglen=32)
    rng = np.random.default_rng(9)
    W = 6
    rec, vals, names = _calib_setup(cfg, 8, 2048., rng, W)
    P = util.draw_bns(rng, W, cfg)
    model = cfg.cuda_model()
    model.plan.set_calibration([rec[d].spline_points for d in cfg.dets])
    lr0, hd0, hh0 = _calib_oracle_batch(cfg.oracle(recalibration=rec), P, vals, names)
    for kernel in (1, 2, 3, 0):
        model.plan.configure(kernel=kernel)
        lr, hd, hh = model.plan.loglr_host(P, calib=vals)
        _check(lr, hd, hh, lr0, hd0, hh0, "calib bns k%d" % kernel)


def test_four_detectors_and_large_batch():
    """n_det = 4 (the narrower TMA tile variant of the kernels) and a 100k-walker batch that
    grows the workspace: both kernels against the oracle on a sample and against each other on
    everything."""
    import gwin_oracle as O
    cfg = util.Config("four", 8, 2048., 20., ["H1", "L1", "V1", "K1"], util.BBH_INJ, seed=5)
    model = cfg.cuda_model()
    rng = np.random.default_rng(6)
    P = util.draw_bbh(rng, 100003, cfg)
    model.plan.configure(kernel=1)
    a = model.plan.loglr_host(P)
    model.plan.configure(kernel=2)
    b = model.plan.loglr_host(P)
    assert a[1].shape == (100003, 4)
    assert np.all(np.abs(a[0] - b[0]) <= tol(a[0]))
    for kernel in (3, 0):
        model.plan.configure(kernel=kernel)
        c = model.plan.loglr_host(P)
        assert np.all(np.abs(a[0] - c[0]) <= tol(a[0])), kernel
    model.plan.configure(kernel=2)
    sel = rng.choice(len(P), 48, replace=False)
    lr0, hd0, hh0 = O.batch_loglr(cfg.oracle(), P[sel])
    _check(b[0][sel], b[1][sel], b[2][sel], lr0, hd0, hh0, "4 detectors")
    # marginalised phase on the same batch
    m = model.plan.loglr_host(P[sel], mode=1)
    lrm = O.batch_loglr(cfg.oracle(phase_marginalization=True), P[sel])[0]
    assert np.all(np.abs(m[0] - lrm) <= tol(lrm))
    # a smaller batch after a larger one reuses the workspace; a different batch size means a
    # different chunking (and block schedule) of the frequency axis, so results agree to
    # rounding, not bit for bit
    c = model.plan.loglr_host(P[:77])
    assert np.all(np.abs(c[0] - b[0][:77]) <= 1e-9 * np.maximum(1.0, np.abs(c[0])))


@pytest.mark.parametrize("dets", [["H1"], ["H1", "L1", "V1", "K1"]])
def test_table_kernel_detector_counts(dets):
    """The moment-table kernel with 1 and 4 detectors (4 needs > 48 KB of dynamic shared memory
    per CTA for the staged rows) on a 64 s segment, against the oracle."""
    import gwin_oracle as O
    inj = dict(util.BNS_INJ)
    inj["tc"] = 60.0
    cfg = util.Config("tabdet", 64, 4096., 20., dets, inj, seed=21)
    model = cfg.cuda_model()
    rng = np.random.default_rng(22)
    P = util.draw_bns(rng, 40, cfg)
    lr0, hd0, hh0 = O.batch_loglr(cfg.oracle(), P)
    for kernel in (2, 3, 0):
        model.plan.configure(kernel=kernel)
        lr, hd, hh = model.plan.loglr_host(P)
        _check(lr, hd, hh, lr0, hd0, hh0, "%d detectors k%d" % (len(dets), kernel))
    assert model.plan.last_launches >= 6            # AUTO took the table path here


def test_table_kernel_coalescence_times_anywhere_in_the_segment():
    """Walkers whose coalescence times are spread over the whole segment -- including across its
    wrap-around -- sit far apart in the moment tables: the warp cannot stage a common set of rows
    and every lane reads its own (and taps wrap around the table row).  Same answers."""
    cfg = util.c3(seglen=64)
    model = cfg.cuda_model()
    rng = np.random.default_rng(23)
    P = util.draw_bns(rng, 96, cfg)
    P[:, 10] = cfg.epoch + rng.uniform(0.0, 64.0, len(P))
    P[:8, 10] = cfg.epoch + np.array([0.0, 1e-3, 63.999, 64.0, 0.01, 63.99, 32.0, 0.5])
    out = {}
    for kernel in (1, 2, 3):
        model.plan.configure(kernel=kernel)
        out[kernel] = model.plan.loglr_host(P)
    for kernel in (2, 3):
        assert np.all(np.abs(out[kernel][0] - out[1][0]) <= tol(out[1][0])), kernel
        assert np.all(np.abs(out[kernel][1] - out[1][1]) <= np.maximum(2e-6, 2e-9 * np.abs(out[1][2]))), kernel
    import gwin_oracle as O
    lr0 = O.batch_loglr(cfg.oracle(), P[:12])[0]
    assert np.all(np.abs(out[3][0][:12] - lr0) <= tol(lr0))


def test_table_coverage_declared_and_automatic():
    """Coverage of the moment tables.  Declared from prior draws (``plan.cover``): a batch inside
    it runs wholly on the table path; a batch outside it is still right -- through the slow path,
    counted.  Automatic: the host entry point measures the first batch and, after a batch that
    sent many walkers down the slow path, measures again."""
    cfg = util.c3(seglen=64)
    model = cfg.cuda_model()
    plan = model.plan
    rng = np.random.default_rng(41)
    ref = util.inj_row(cfg)
    P = util.draw_bns(rng, 400, cfg)
    f = 1.0 + 0.004 * rng.uniform(-1, 1, len(P))          # chirp mass within +-0.4 %
    P[:, 0], P[:, 1] = ref[0] * f, ref[1] * f * (1.0 + 0.01 * rng.uniform(-1, 1, len(P)))
    a = _direct(plan, P)
    plan.configure(kernel=3)
    plan.cover(P[:200])
    b = plan.loglr_host(P)                              # the tables are built on first use
    narrow = plan.table_info()
    assert narrow["subblocks"] > 0 and narrow["references"] >= narrow["subblocks"]
    assert plan.slow_path_walkers() == 0
    _agree(b, a, "declared narrow coverage")
    lr0 = _oracle_batch(cfg.oracle(), P[:6])[0]
    assert np.all(np.abs(b[0][:6] - lr0) <= tol(lr0))
    # the whole mass box is far outside the declared coverage: slow path, same answers
    Pw = util.draw_bns(rng, 96, cfg)
    w = plan.loglr_host(Pw)
    n_slow = plan.slow_path_walkers()
    assert n_slow > 48 and plan.slow_path_walkers() == 0   # reading resets the count
    d = _direct(plan, Pw)
    plan.configure(kernel=3)
    _agree(w, d, "outside the declared coverage")
    # a declared coverage is not re-planned behind the caller's back ...
    assert plan.table_info() == narrow
    # ... adding the box to it is the caller's move: more references, nobody on the slow path
    plan.cover(Pw)
    w2 = plan.loglr_host(Pw)
    wide = plan.table_info()
    assert plan.slow_path_walkers() == 0 and wide["references"] > narrow["references"]
    _agree(w2, d, "coverage extended to the box")
    _agree(plan.loglr_host(P), a, "narrow batch on the wide tables")
    # automatic coverage: first batch measured; a batch that leaves it is right (slow path) and
    # triggers a new measurement, so the batch after it is fast again
    plan.cover(None)
    _agree(plan.loglr_host(P), a, "automatic, narrow batch")
    assert plan.slow_path_walkers() == 0
    auto_narrow = plan.table_info()
    _agree(plan.loglr_host(Pw), d, "automatic, first wide batch")
    assert plan.slow_path_walkers() > 48
    _agree(plan.loglr_host(Pw), d, "automatic, second wide batch")
    assert plan.slow_path_walkers() == 0 and plan.table_info()["references"] > auto_narrow["references"]


def test_linearity_in_data(cfg2):
    """Size-independent property: <d|h> is linear in the data, <h|h> does not
    depend on it."""
    rng = np.random.default_rng(2)
    P = util.draw_bbh(rng, 64, cfg2)
    a = cfg2.cuda_model().plan.loglr_host(P)
    cfgb = util.c2(seed=99)
    b = cfgb.cuda_model().plan.loglr_host(P)
    cfgs = util.c2(seed=99)
    for d in cfgs.dets:
        cfgs.data[d] = cfg2.data[d] + cfgb.data[d]
    s = cfgs.cuda_model().plan.loglr_host(P)
    scale = np.abs(a[1]) + np.abs(b[1]) + 1.0
    assert np.all(np.abs(s[1] - (a[1] + b[1])) <= 1e-12 * scale)
    np.testing.assert_allclose(s[2], a[2], rtol=1e-14)


# ---------------------------------------------------------------------------
# round 2: the configuration users and bench.py actually run (kernel AUTO at C3), wide
# parameter ranges through the moment tables, and the slow-path accounting
# ---------------------------------------------------------------------------
def _direct(plan, P, **kw):
    plan.configure(kernel=1)
    out = plan.loglr_host(P, **kw)
    plan.configure(kernel=0)
    return out


def _agree(a, b, what=""):
    err = np.abs(a[0] - b[0])
    assert np.all(err <= tol(b[0])), "%s: max |dloglr| %.3e at %d" % (what, err.max(), err.argmax())
    assert np.all(np.abs(a[1] - b[1]) <= np.maximum(2e-6, 2e-9 * np.abs(b[2]))), what
    assert np.all(np.abs(a[2] - b[2]) <= np.maximum(1e-6, 1e-9 * np.abs(b[2]))), what


def test_bench_configuration_c3_auto(cfg3):
    """The headline configuration exactly as bench.py builds it: C3, plain model constructor,
    kernel AUTO, 4096 walkers from the bench's mass box -- held to the direct kernel (all), the
    oracle (24) and the extended-precision oracle (6); no walker may have left the fast path."""
    import bench
    model = cfg3.cuda_model()
    plan = model.plan
    rng = np.random.default_rng(20182)
    P = bench.draw_walkers(rng, 4096)
    P[:, 10] += cfg3.inj["tc"] - bench.INJ["tc"]
    auto = plan.loglr_host(P)
    assert plan.slow_path_walkers() == 0
    assert plan.last_launches >= 5                  # AUTO took the table path
    _agree(auto, _direct(plan, P), "bench config, AUTO vs direct")
    sel = np.argsort(P[:, 0] + P[:, 1])[np.linspace(0, len(P) - 1, 24).astype(int)]   # spans the sorted ke range
    lr0, hd0, hh0 = _oracle_batch(cfg3.oracle(), P[sel])
    _check(auto[0][sel], auto[1][sel], auto[2][sel], lr0, hd0, hh0, "bench config vs oracle")
    lrT = _oracle_batch(cfg3.oracle(mode="truth"), P[sel[:6]])[0]
    assert np.all(np.abs(auto[0][sel[:6]] - lrT) <= tol(lrT))
    # the device entry point gives the same bits as the host entry point
    import torch
    dev = plan.loglr_device(torch.tensor(P, device="cuda"))
    np.testing.assert_array_equal(dev[0].cpu().numpy(), auto[0])
    # a walker's result does not depend on the batch it is evaluated in (its neighbours in a warp set the
    # window of table rows the tensor-core product runs over, not the numbers it adds up for the walker)
    for pick in (slice(0, None, 3), rng.permutation(len(P))[:1500], slice(7, 8)):
        sub = plan.loglr_host(np.ascontiguousarray(P[pick]))
        for x, y in zip(sub, auto):
            np.testing.assert_array_equal(x, y[pick])


def test_calibrated_batches_take_the_table_path_c3(cfg3):
    """Calibrated likelihood (gwin/calibration.py:142-173) through the moment tables at the headline
    size: the log-calibration's Taylor coefficients join the per-detector series, the first moment comes
    from the derivative of the interpolation weights.  2048 walkers with 8-point splines on three
    detectors -- values of the size the reference's priors give (a few per cent) and ten times that --
    against the calibrated direct kernel; a sample against the oracle's scipy splines."""
    import bench
    rng = np.random.default_rng(77)
    W = 2048
    P = bench.draw_walkers(rng, W)
    P[:, 10] += cfg3.inj["tc"] - bench.INJ["tc"]
    model = cfg3.cuda_model()
    plan = model.plan
    for scale in (0.03, 0.3):
        rec, vals, names = _calib_setup(cfg3, 8, 2048., rng, W, scale=scale)
        plan.set_calibration([rec[d].spline_points for d in cfg3.dets])
        plan.configure(kernel=0)
        auto = plan.loglr_host(P, calib=vals)
        n_slow = plan.slow_path_walkers()
        assert plan.last_launches >= 5                  # AUTO took the table path
        assert n_slow <= (0 if scale < 0.1 else W // 4), n_slow      # (large values: some walkers fall back to the direct kernel)
        _agree(auto, _direct(plan, P, calib=vals), "calibrated, AUTO vs direct, scale %g" % scale)
        plan.configure(kernel=3)
        _agree(plan.loglr_host(P, calib=vals), auto, "calibrated, TABLE vs AUTO")
        plan.configure(kernel=0)
        if scale < 0.1:
            sel = np.argsort(P[:, 0] + P[:, 1])[np.linspace(0, W - 1, 3).astype(int)]
            lr0, hd0, hh0 = _calib_oracle_batch(cfg3.oracle(recalibration=rec), P[sel], vals[sel], names)
            _check(auto[0][sel], auto[1][sel], auto[2][sel], lr0, hd0, hh0, "calibrated table path vs oracle")
            # and the calibration matters at this size
            plain = plan.loglr_host(P)
            assert np.max(np.abs(plain[0] - auto[0])) > 1.0
    plan.set_calibration(None)


def _draw_wide(rng, W, cfg, spin=0.9, qmax=3.0, lam=5000.0, mc=(1.0, 1.6)):
    """Spinning, unequal-mass, tidally deformed systems: chirp mass in ``mc``, mass ratio up to
    ``qmax``, aligned spins up to ``spin``, tidal deformabilities up to ``lam``."""
    P = util.draw_bns(rng, W, cfg)
    mchirp = rng.uniform(mc[0], mc[1], W)
    q = rng.uniform(1.0, qmax, W)
    eta = q / (1.0 + q) ** 2
    mtot = mchirp * eta ** -0.6
    P[:, 0] = mtot * q / (1.0 + q)
    P[:, 1] = mtot / (1.0 + q)
    P[:, 2] = rng.uniform(-spin, spin, W)
    P[:, 3] = rng.uniform(-spin, spin, W)
    P[:, 11] = rng.uniform(0.0, lam, W)
    P[:, 12] = rng.uniform(0.0, lam, W)
    return P


@pytest.mark.parametrize("seglen", [64, 256])
def test_table_kernel_spins_mass_ratio_tides(seglen):
    """Long segments, kernel AUTO and TABLE over spins +-0.9, q <= 3, lambda <= 5000 (the tidal
    phase at 1 kHz is comparable to the point-particle phase): every walker within tolerance
    of the direct kernel, a sample within tolerance of the oracle.  Walkers the tables cannot
    take go through the slow path -- counted, never wrong."""
    cfg = util.c3(seglen=seglen)
    model = cfg.cuda_model()
    plan = model.plan
    rng = np.random.default_rng(900 + seglen)
    W = 1024 if seglen == 64 else 512
    for what, P in (("wide", _draw_wide(rng, W, cfg)),
                    ("extreme corners", _draw_wide(rng, 64, cfg, mc=(1.55, 1.6), qmax=1.02))):
        if what == "extreme corners":               # all spins / tides at their bounds, both signs
            P[:, 2] = np.where(np.arange(len(P)) & 1, 0.9, -0.9)
            P[:, 3] = np.where(np.arange(len(P)) & 2, 0.9, -0.9)
            P[:, 11] = np.where(np.arange(len(P)) & 4, 5000.0, 0.0)
            P[:, 12] = np.where(np.arange(len(P)) & 8, 5000.0, 0.0)
        ref = _direct(plan, P)
        for kernel in (0, 3, 2):
            plan.configure(kernel=kernel)
            _agree(plan.loglr_host(P), ref, "%s T=%d k%d" % (what, seglen, kernel))
        plan.configure(kernel=0)
        n = 12 if seglen == 64 else 6
        lr0, hd0, hh0 = _oracle_batch(cfg.oracle(), P[:n])
        out = plan.loglr_host(P[:n])
        _check(out[0], out[1], out[2], lr0, hd0, hh0, "%s T=%d vs oracle" % (what, seglen))


def test_device_entry_walkers_below_schedule_mchirp_min(cfg3):
    """The device entry points do not see the batch before launching: a walker lighter than the
    schedule's smallest chirp mass (or outside the range the moment tables cover) must come out
    right anyway -- through the slow path -- and be counted."""
    import torch
    model = cfg3.cuda_model()
    plan = model.plan
    rng = np.random.default_rng(77)
    P = util.draw_bns(rng, 256, cfg3)
    plan.loglr_host(P)                              # schedule + tables now follow this batch
    plan.slow_path_walkers()
    Q = P.copy()
    Q[::16, 0] = 0.5                                # 16 much lighter systems
    Q[::16, 1] = 0.45
    dev = plan.loglr_device(torch.tensor(Q, device="cuda"))
    torch.cuda.synchronize()
    assert plan.slow_path_walkers() >= 16
    ref = _direct(plan, Q)
    got = (dev[0].cpu().numpy(), dev[1].cpu().numpy().view(np.complex128)[..., 0], dev[2].cpu().numpy())
    _agree(got, ref, "device entry, light walkers")
    # the host entry point sees how many walkers took the slow path and measures the coverage again
    # before the next batch: right both times, fast the second time
    _agree(plan.loglr_host(Q), ref, "host entry, light walkers")
    assert plan.slow_path_walkers() >= 16
    _agree(plan.loglr_host(Q), ref, "host entry, light walkers, re-planned")
    assert plan.slow_path_walkers() == 0


def test_interleaved_host_and_device_calls_on_different_streams(cfg2):
    """Calls on one plan from different streams share its workspace; the library orders them."""
    import torch
    model = cfg2.cuda_model()
    plan = model.plan
    rng = np.random.default_rng(5)
    P = util.draw_bbh(rng, 4096, cfg2)
    ref = plan.loglr_host(P)
    Pd = torch.tensor(P, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for i in range(6):
        st = (s1, s2)[i & 1]
        st.wait_stream(torch.cuda.current_stream())
        outs.append(plan.loglr_device(Pd, stream=st.cuda_stream))
        if i == 3:
            np.testing.assert_array_equal(plan.loglr_host(P)[0], ref[0])
    torch.cuda.synchronize()
    for o in outs:
        np.testing.assert_array_equal(o[0].cpu().numpy(), ref[0])


def _upstream_files():
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(glob.glob(os.path.join(here, "*_upstream.npz")))


@pytest.mark.skipif(not _upstream_files(), reason="no tests/golden/*_upstream.npz (see make_upstream_golden.py)")
@pytest.mark.parametrize("path", _upstream_files())
def test_upstream_golden_vectors_cuda(path):
    """The CUDA path against vectors computed by the reference's own classes over pycbc /
    lalsimulation, every kernel."""
    u = np.load(path)
    cfg = util.golden_config(os.path.basename(path).replace("_upstream", ""))
    model = cfg.cuda_model()
    for kernel in (1, 2, 3, 0):
        model.plan.configure(kernel=kernel)
        lr = model.plan.loglr_host(u["params"], mode=int(bool(u["marg"])))[0]
        assert np.all(np.abs(lr - u["loglr"]) <= tol(u["loglr"])), kernel
