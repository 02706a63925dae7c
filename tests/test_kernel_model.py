"""The CUDA kernels' arithmetic, modelled in numpy (tests/kernel_model.py), against
the oracle: coefficient folding, closed-form <h|h>, and the block recurrence with
its schedule.  Runs without a GPU; the GPU tests check the real kernels."""
import numpy as np
import pytest

import gwin_oracle as O
import util
from kernel_model import PlanModel


def _plan(cfg, **kw):
    D = [O.Detector(d) for d in cfg.dets]
    return PlanModel(cfg.delta_f, cfg.epoch, cfg.f_lower, None, cfg.f_lower,
                     [d.response for d in D], [d.location for d in D],
                     np.stack([cfg.data[d] for d in cfg.dets]),
                     np.stack([cfg.psd] * len(cfg.dets)), **kw)


@pytest.mark.parametrize("epoch", [0.0, 1126259462.0])
def test_direct_and_recurrence_match_oracle_bbh(epoch):
    cfg = util.c2(epoch=epoch)
    pm = _plan(cfg, mchirp_min=8.0)
    orc = cfg.oracle()
    rng = np.random.default_rng(1)
    P = util.draw_bbh(rng, 5, cfg)
    P[0] = util.inj_row(cfg)
    for row in P:
        lr0 = orc.loglr(**dict(zip(util.PARAM_ORDER, row)))[0]
        assert abs(pm.loglr_direct(row)[0] - lr0) <= util.tol(lr0)
        lr, hd, hh, perr = pm.loglr_recur(row, return_phase_err=True)
        assert abs(lr - lr0) <= util.tol(lr0)
        assert perr <= pm.phase_tol          # the schedule's bound holds


def test_recurrence_bns_256s_schedule_bound():
    """C3: quartic blocks sized for mchirp >= 1.0 keep the block-polynomial phase
    error under phase_tol for walkers of the C3 mass box, and loglr within 1e-6."""
    cfg = util.c3()
    pm = _plan(cfg, mchirp_min=1.0)
    bs = np.diff(pm.blk_start)
    assert bs.sum() == pm.kmax - pm.kmin and bs.min() >= 4 and bs.max() <= 263
    assert bs.mean() > 150                  # long blocks dominate: setup is amortised
    orc = cfg.oracle()
    rng = np.random.default_rng(2)
    P = util.draw_bns(rng, 2, cfg)
    P[0, :2] = 1.2, 1.2                     # lowest chirp mass of the box
    for row in P:
        lr0 = orc.loglr(**dict(zip(util.PARAM_ORDER, row)))[0]
        lr, hd, hh, perr = pm.loglr_recur(row, return_phase_err=True)
        assert perr <= pm.phase_tol
        assert abs(lr - lr0) <= util.tol(lr0)


def test_marginalized_phase_model():
    cfg = util.c2()
    pm = _plan(cfg, mchirp_min=8.0)
    orc = cfg.oracle(phase_marginalization=True)
    row = util.inj_row(cfg)
    lr0 = orc.loglr(**dict(zip(util.PARAM_ORDER, row)))[0]
    assert abs(pm.loglr_recur(row, marg=True)[0] - lr0) <= util.tol(lr0)


def test_tidal_terms_in_the_kernel_arithmetic():
    cfg = util.c3(seglen=32)
    pm = _plan(cfg, mchirp_min=1.0)
    orc = cfg.oracle()
    rng = np.random.default_rng(3)
    P = util.draw_bns(rng, 2, cfg)
    P[:, 11] = 800., 2500.
    P[:, 12] = 1500., 0.
    for row in P:
        lr0 = orc.loglr(**dict(zip(util.PARAM_ORDER, row)))[0]
        assert abs(pm.loglr_direct(row)[0] - lr0) <= util.tol(lr0)
        lr, hd, hh, perr = pm.loglr_recur(row, return_phase_err=True)
        assert abs(lr - lr0) <= util.tol(lr0) and perr <= pm.phase_tol


def test_calibration_arithmetic():
    """The kernels' calibration maths against the oracle's scipy UnivariateSpline
    (gwin/calibration.py:142-173): least-squares cubic, closed-form <h|h> with the sextic
    amplitude factor, and the log-calibration joining the block polynomial."""
    cfg = util.c1()
    rng = np.random.default_rng(1)
    n = 6
    rec = {d: O.CubicSpline(20., 600., n, d.lower()) for d in cfg.dets}
    orc = cfg.oracle(recalibration=rec)
    pm = _plan(cfg, mchirp_min=5.0)
    pm.set_calibration([rec[d].spline_points for d in cfg.dets])
    for row in util.draw_bbh(rng, 3, cfg):
        vals = rng.normal(0, 0.08, (len(cfg.dets), 2, n))
        kw = dict(zip(util.PARAM_ORDER, row))
        for i, d in enumerate(cfg.dets):
            for j in range(n):
                kw['recalib_amplitude_%s_%d' % (d.lower(), j)] = vals[i, 0, j]
                kw['recalib_phase_%s_%d' % (d.lower(), j)] = vals[i, 1, j]
        lr0, hd0, hh0 = orc.loglr(**kw)
        for fn in (pm.loglr_direct, pm.loglr_recur):
            lr, hd, hh = fn(row, calib=vals)
            assert abs(lr - lr0) <= util.tol(lr0)
            assert np.allclose(hh, list(hh0.values()), rtol=1e-10)
        assert abs(pm.loglr_recur(row)[0] - lr0) > 1e-3      # the calibration matters


def test_table_kernel_arithmetic():
    """The moment-table scheme (analytic Taylor coefficients per sub-block, reference chirps,
    deconvolved DFT tables, 13-tap exponential-of-semicircle interpolation) against the direct
    sum and the oracle, for a coverage measured on the walkers themselves."""
    cfg = util.c3(seglen=32)
    pm = _plan(cfg, mchirp_min=1.0)
    rng = np.random.default_rng(3)
    rows = util.draw_bns(rng, 3, cfg)
    rows[2, 2:4] = 0.04, -0.03                      # small spins
    rows[2, 11:13] = 300.0, 500.0                   # and tides
    pm.build_tables(pm.cover_stats(rows))
    ms = [B['m'] for B in pm.tblk]
    assert pm.k_split < pm.kmax and min(ms) >= pm.TAB_MIN and max(ms) <= pm.TAB_MAX
    assert set(ms) <= set(pm.table_sizes()) and max(B['K'] for B in pm.tblk) > 1
    orc = cfg.oracle()
    for row in rows:
        a = pm.loglr_direct(row)
        b = pm.loglr_table(row)
        assert not pm.slow and pm.worst_est <= pm.TAB_EST_MAX
        assert abs(a[0] - b[0]) < 1e-8 and np.max(np.abs(a[1] - b[1])) < 1e-8
        lr0 = orc.loglr(**dict(zip(util.PARAM_ORDER, row)))[0]
        assert abs(b[0] - lr0) <= util.tol(lr0)


def test_calibrated_table_kernel_arithmetic():
    """loglr_table_kernel<CALIB>: the log-calibration's Taylor coefficients about a sub-block's centre
    (differential equations of log(1 + r) and atan: no transcendental function), the per-detector
    series with a first-order term, and the first moment from the derivative of the interpolation
    weights (csrc/tab_poly.inc: TAB_DWEIGHTS) -- against the calibrated direct sum and the oracle's
    scipy splines (gwin/calibration.py:142-173)."""
    cfg = util.c3(seglen=32)
    n = 8
    rec = {d: O.CubicSpline(20., 2048., n, d.lower()) for d in cfg.dets}
    pm = _plan(cfg, mchirp_min=1.0)
    pm.set_calibration([rec[d].spline_points for d in cfg.dets])
    rng = np.random.default_rng(4)
    rows = util.draw_bns(rng, 2, cfg)
    pm.build_tables(pm.cover_stats(rows))
    # the generated weight polynomials and their derivative against the kernel function itself
    W, beta = pm.TAB_W, 2.30 * pm.TAB_W
    for u0 in (-0.49, -0.2, 0.0, 0.31, 0.45):
        xs = u0 + (W - 1) / 2.0 - np.arange(W)
        assert np.max(np.abs(pm.tab_weights(2 * u0) - pm.es_kernel(xs, W, beta))) < 1e-12
        fd = (pm.es_kernel(xs + 1e-6, W, beta) - pm.es_kernel(xs - 1e-6, W, beta)) / 2e-6
        assert np.max(np.abs(pm.tab_weights(2 * u0, derivative=True) - fd)) < 1e-8
    # the series of log c_d against numpy's log1p / arctan along a sub-block
    vals = rng.normal(0, 0.05, (len(cfg.dets), 2, n))
    coef = pm.calib_coef(vals)
    B = pm.tblk[len(pm.tblk) // 2]
    kc, h = B['start'] + B['m'] // 2, B['m'] / 2.0
    Ls, c0, amp = pm.calib_series(coef, 1, kc, h)
    k = B['start'] + np.arange(B['m'])
    s = (k - kc) / h
    exact = pm.calib_log(coef, 1, k) - pm.calib_log(coef, 1, np.array([kc]))[0]
    assert np.max(np.abs(sum(Ls[m] * s ** m for m in range(1, 7)) - exact)) < 1e-13
    assert abs(c0 - pm.calib_factor(coef, 1, np.array([kc]))[0]) < 1e-15 and amp > 0.5
    orc = cfg.oracle(recalibration=rec)
    for row in rows:
        vals = rng.normal(0, 0.05, (len(cfg.dets), 2, n))
        a = pm.loglr_direct(row, calib=vals)
        b = pm.loglr_table(row, calib=vals)
        assert not pm.slow and pm.worst_est <= pm.TAB_EST_MAX
        assert abs(a[0] - b[0]) < 2e-8 and np.max(np.abs(a[1] - b[1])) < 2e-8
        kw = dict(zip(util.PARAM_ORDER, row))
        for i, d in enumerate(cfg.dets):
            for j in range(n):
                kw['recalib_amplitude_%s_%d' % (d.lower(), j)] = vals[i, 0, j]
                kw['recalib_phase_%s_%d' % (d.lower(), j)] = vals[i, 1, j]
        lr0 = orc.loglr(**kw)[0]
        assert abs(b[0] - lr0) <= util.tol(lr0)
        assert abs(pm.loglr_table(row)[0] - b[0]) > 1e-2        # the calibration matters
    # a detector whose amplitude factor comes near zero leaves the table path
    bad = np.zeros((len(cfg.dets), 2, n))
    bad[0, 0, :] = -0.9
    pm.loglr_table(rows[0], calib=bad)
    assert pm.slow


def test_table_truncation_estimate_flags_uncovered_walkers():
    """A walker outside what the tables were built for -- far in chirp mass, or with spins the
    coverage did not see -- is recognised from its own truncation estimate (the kernel then takes
    the slow path); the majorant the schedule is built with bounds the estimate of covered ones."""
    cfg = util.c3(seglen=32)
    ref = util.inj_row(cfg)
    pm = _plan(cfg, mchirp_min=1.0)
    rng = np.random.default_rng(9)
    rows = np.tile(ref, (4, 1))
    rows[:, :2] *= 1.0 + 0.004 * rng.uniform(-1, 1, (4, 1))
    pm.build_tables(pm.cover_stats(rows))
    for row in rows:
        a, b = pm.loglr_direct(row), pm.loglr_table(row)
        assert not pm.slow and pm.worst_est <= 4 * pm.TAB_EST_DESIGN
        assert abs(a[0] - b[0]) < 1e-8
    far = ref.copy()
    far[:2] *= 0.7
    pm.loglr_table(far)
    assert pm.slow
    spinning = ref.copy()
    spinning[2:4] = 0.8, -0.6
    pm.loglr_table(spinning)
    assert pm.slow and pm.worst_est > pm.TAB_EST_MAX


def test_fine_decomposition_of_the_tails():
    """The closed-form choice of fine sub-blocks tiles the stretch between the last whole sub-block
    and the walker's last bin exactly, left to right, with at most one sub-block per level."""
    from kernel_model import PlanModel
    for pm, fmin in ((8192, 128), (4096, 128), (3072, 192), (6144, 192), (256, 128)):
        for length in list(range(0, min(pm, 700))) + [pm - 1, pm - fmin, pm // 2, pm // 2 - 1, pm // 2 + 1]:
            taken, left = PlanModel.fine_decomposition(length, pm, fmin)
            pos = 0
            for off, size in taken:
                assert off == pos and off % size == 0 and size >= fmin and off + size <= pm
                pos += size
            assert pos + left == length and 0 <= left < fmin
            assert len(set(size for _, size in taken)) == len(taken)
