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


def make(name, cfg, P, marg):
    orc = cfg.oracle(phase_marginalization=marg)
    orcT = cfg.oracle(phase_marginalization=marg, mode="truth")
    lr, hd, hh = O.batch_loglr(orc, P)
    lrT, hdT, hhT = O.batch_loglr(orcT, P)
    np.savez(os.path.join(HERE, name + ".npz"), params=P, loglr=lr, hd=hd, hh=hh,
             loglr_truth=lrT, hd_truth=hdT, hh_truth=hhT, lognl=orc.lognl(),
             kmin=orc._kmin, kmax=orc._kmax, marg=marg, dets=np.array(cfg.dets),
             data_sha256=digest(cfg))
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


if __name__ == "__main__":
    main()
