"""Generates tests/golden/*.npz from the oracle (faithful float64 + extended
precision "truth" modes).

The reference itself cannot run here (Python-2-only, pycbc/lalsuite absent), so
these are ORACLE outputs, committed so that (a) the oracle cannot drift
silently, (b) the C restatement and the CUDA path are checked against fixed
numbers, not only against a live re-computation.  Data are regenerated from the
seeds in tests/util.py; a SHA-256 of the data bytes is stored to detect a change
of the generator.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import util  # noqa: E402
import gwin_oracle as O  # noqa: E402


def digest(cfg):
    h = hashlib.sha256()
    for d in cfg.dets:
        h.update(np.ascontiguousarray(cfg.data[d]).tobytes())
    return h.hexdigest()


def make(name, cfg, P, marg, dist_marg=None, calib=None):
    """dist_marg = (lo, hi): Uniform(distance=(lo, hi)) distance marginalisation on the
    reference's 10**4-point grid; calib = (fmin, fmax, n_points, values[W, n_det*2*n_points]):
    cubic-spline recalibration of every detector."""
    kw, extra = {}, {}
    if dist_marg is not None:
        kw["distance_marginalization"] = util.dist_grid(*dist_marg)
        extra["dist_marg"] = np.array(dist_marg, dtype=float)
    names = []
    if calib is not None:
        fmin, fmax, n, vals = calib
        kw["recalibration"] = {d: O.CubicSpline(fmin, fmax, n, d.lower()) for d in cfg.dets}
        names = util.calib_names(cfg.dets, n)
        extra.update(calib_spec=np.array([fmin, fmax, n], dtype=float), calib_values=vals)

    def batch(orc):
        if calib is None:
            return O.batch_loglr(orc, P)
        out = [orc.loglr(**dict(list(zip(util.PARAM_ORDER, row)) + list(zip(names, v))))
               for row, v in zip(P, calib[3])]
        return (np.array([o[0] for o in out]),
                np.array([[o[1][d] for d in cfg.dets] for o in out]),
                np.array([[o[2][d] for d in cfg.dets] for o in out]))
    orc = cfg.oracle(phase_marginalization=marg, **kw)
    orcT = cfg.oracle(phase_marginalization=marg, mode="truth", **kw)
    lr, hd, hh = batch(orc)
    lrT, hdT, hhT = batch(orcT)
    np.savez(os.path.join(HERE, name + ".npz"), params=P, loglr=lr, hd=hd, hh=hh,
             loglr_truth=lrT, hd_truth=hdT, hh_truth=hhT, lognl=orc.lognl(),
             kmin=orc._kmin, kmax=orc._kmax, marg=marg, dets=np.array(cfg.dets),
             data_sha256=digest(cfg), **extra)
    print(name, "W=%d" % len(P), "max|faithful-truth| = %.2e" % np.abs(lr - lrT).max())


def main():
    rng = np.random.default_rng(4242)
    cfg = util.c1()
    P = util.draw_bbh(rng, 16, cfg)
    P[0] = util.inj_row(cfg)
    make("c1_gaussian_h1l1", cfg, P, False)
    cfg = util.c2()
    P = util.draw_bbh(rng, 16, cfg)
    P[0] = util.inj_row(cfg)
    make("c2_margphase_h1l1v1", cfg, P, True)
    cfg = util.c3()
    P = util.draw_bns(rng, 4, cfg)
    P[0] = util.inj_row(cfg)
    make("c3_bns256_h1l1v1", cfg, P, False)
    cfg = util.c2(epoch=1126259462.0, seed=7)
    P = util.draw_bbh(rng, 8, cfg)
    make("c2_gps_epoch", cfg, P, False)
    # round-1 extensions: tidal terms, distance marginalisation, calibration
    cfg = util.c3(seglen=32)
    P = util.draw_bns(rng, 6, cfg)
    P[:, 11] = rng.uniform(0., 3000., len(P))
    P[:, 12] = rng.uniform(0., 3000., len(P))
    make("c3s32_tidal", cfg, P, False)
    cfg = util.c2()
    P = util.draw_bbh(rng, 8, cfg)
    P[:, 4] = 1.0                                  # templates at 1 Mpc: what x / D_j assumes
    make("c2_margdist", cfg, P, False, dist_marg=(50., 5000.))
    make("c2_margdistphase", cfg, P, True, dist_marg=(50., 5000.))
    P = util.draw_bbh(rng, 8, cfg)
    vals = rng.normal(0., 0.05, (len(P), len(cfg.dets) * 2 * 6))
    vals[1] *= 5.0
    make("c2_calibrated", cfg, P, False, calib=(20., 1024., 6, vals))


if __name__ == "__main__":
    main()
