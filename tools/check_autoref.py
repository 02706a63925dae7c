"""GWIN_TABLE_AUTOREF=1 python tools/check_autoref.py : the host entry point chooses the reference
waveform of the moment tables from the batches.  A sequence of batches (wide, narrowing, drifting
out of the covered range, wide again) is evaluated with kernel AUTO and compared with the direct
kernel; prints timing per batch and the worst deviation; exits non-zero on a tolerance failure."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import util

seglen = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
cfg = util.c3(seglen=seglen)
model = cfg.cuda_model()
plan = model.plan
rng = np.random.default_rng(5)
ref = util.inj_row(cfg)
worst = 0.0
ok = True
for label, scale, centre in (("box", None, 1.0), ("3 %", 0.03, 1.0), ("0.5 %", 0.005, 1.0), ("0.5 %", 0.005, 1.0),
                             ("0.5 % shifted", 0.005, 1.02), ("0.1 %", 0.001, 1.02), ("box", None, 1.0)):
    P = util.draw_bns(rng, W, cfg)
    if scale is not None:
        f = centre * (1.0 + scale * rng.uniform(-1, 1, W))
        P[:, 0], P[:, 1] = ref[0] * f, ref[1] * f
    plan.configure(kernel=0)
    plan.loglr_host(P)                      # (re)builds tables if the policy asks for it
    t0 = time.perf_counter()
    a = plan.loglr_host(P)
    dt = time.perf_counter() - t0
    launches = plan.last_launches
    plan.configure(kernel=1)
    d = plan.loglr_host(P)
    err = np.abs(a[0] - d[0])
    lim = util.tol(d[0])
    worst = max(worst, float(err.max()))
    good = bool(np.all(err <= lim))
    ok = ok and good
    print("%-14s W=%d  %.3f ms  launches %d  max |auto - direct| %.2e  %s" % (label, W, dt * 1e3, launches, err.max(), "ok" if good else "FAIL"))
print("worst %.2e; %s" % (worst, "OK" if ok else "FAILED"))
sys.exit(0 if ok else 1)
