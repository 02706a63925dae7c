"""Pins the committed golden vectors against the REAL stack (run where pycbc + lalsuite exist).

The reference cannot run in the build container (Python-2-only gwin, pycbc / lalsuite absent), so
tests/golden/*.npz are outputs of this repository's own oracle: numeric parity with upstream is
unpinned.  This script closes that gap wherever the upstream packages are
installed: for every parameter row of every plain / phase-marginalised golden file it evaluates the
log likelihood ratio with the reference's own model classes -- ``gwin.models`` (gwastro/gwin, Python 2)
or their later home ``pycbc.inference.models`` -- on the SAME data (regenerated from the seeds in
tests/util.py; the SHA-256 stored in the golden file is checked) and writes
``tests/golden/<name>_upstream.npz``.  ``tests/test_oracle.py::test_upstream_golden_vectors`` and
``tests/test_gpu_parity.py::test_upstream_golden_vectors_cuda`` pick such files up when present and
hold the oracle and the CUDA path to them at the north-star tolerance.

    python tests/golden/make_upstream_golden.py            # needs: pycbc, lalsuite (and gwin or pycbc.inference)

Reference call sites reproduced: test/utils/base.py:22-57 (model fixtures), gwin/models/gaussian_noise.py:219-350,
gwin/models/marginalized_gaussian_noise.py:232-272, 456-511.
"""
from __future__ import print_function

import glob
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))

# only files whose model the reference's classes cover directly
PLAIN = ("c1_gaussian_h1l1.npz", "c2_margphase_h1l1v1.npz", "c3_bns256_h1l1v1.npz", "c2_gps_epoch.npz",
         "c3s32_tidal.npz")


def upstream_models():
    try:
        from gwin import models                      # gwastro/gwin (Python 2)
        return models, "gwin"
    except ImportError:
        from pycbc.inference import models           # the same classes, later home
        return models, "pycbc.inference"


def main():
    import util
    from pycbc.types import FrequencySeries
    from pycbc.waveform.generator import FDomainCBCGenerator, FDomainDetFrameGenerator
    import pycbc
    models, origin = upstream_models()
    for path in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        name = os.path.basename(path)
        if name not in PLAIN:
            continue
        g = np.load(path)
        cfg = util.golden_config(name)
        h = hashlib.sha256()
        for d in cfg.dets:
            h.update(np.ascontiguousarray(cfg.data[d]).tobytes())
        if "data_sha256" in g and str(g["data_sha256"]) != h.hexdigest():
            raise SystemExit("%s: regenerated data differ from the data the golden file was made with" % name)
        marg = bool(g["marg"])
        variable = list(util.PARAM_ORDER)
        gen = FDomainDetFrameGenerator(FDomainCBCGenerator, cfg.epoch, detectors=cfg.dets, variable_args=variable,
                                       delta_f=cfg.delta_f, f_lower=cfg.f_lower, approximant="TaylorF2")
        data = dict((d, FrequencySeries(cfg.data[d], delta_f=cfg.delta_f, epoch=cfg.epoch)) for d in cfg.dets)
        psds = dict((d, FrequencySeries(cfg.psd, delta_f=cfg.delta_f)) for d in cfg.dets)
        if marg:
            try:
                model = models.MarginalizedPhaseGaussianNoise(variable, data, gen, cfg.f_lower, psds=psds)
            except AttributeError:
                model = models.MarginalizedGaussianNoise(variable, data, gen, cfg.f_lower, psds=psds,
                                                         phase_marginalization=True, marg_prior=[])
        else:
            model = models.GaussianNoise(variable, data, gen, cfg.f_lower, psds=psds)
        P = g["params"]
        lr = np.empty(len(P))
        hh = np.full((len(P), len(cfg.dets)), np.nan)
        for i, row in enumerate(P):
            model.update(**dict(zip(variable, row)))
            lr[i] = model.loglr
            for j, d in enumerate(cfg.dets):
                try:
                    hh[i, j] = model.det_optimal_snrsq(d)
                except Exception:          # older classes keep the per-detector numbers in current_stats
                    hh[i, j] = getattr(model.current_stats, "%s_optimal_snrsq" % d, np.nan)
        out = os.path.join(HERE, name[:-4] + "_upstream.npz")
        np.savez(out, params=P, loglr=lr, hh=hh, lognl=float(model.lognl), marg=marg,
                 origin="%s / pycbc %s" % (origin, pycbc.__version__))
        print("%s: max |loglr_upstream - loglr_oracle| = %.3e" % (name, np.max(np.abs(lr - g["loglr"]))))


if __name__ == "__main__":
    main()
