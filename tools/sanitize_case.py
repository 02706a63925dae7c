"""Small case for compute-sanitizer: both kernels, both modes, ragged batch, edge rows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import util
cfg = util.c2()
model = cfg.cuda_model()
rng = np.random.default_rng(0)
P = util.draw_bbh(rng, 333, cfg)
P[5, 0] = np.nan; P[6, 0] = P[6, 1] = 200.; P[7, 0] = P[7, 1] = 1.0
for kernel in (1, 2):
    model.plan.configure(kernel=kernel)
    for mode in (0, 1):
        lr, hd, hh = model.plan.loglr_host(P, mode=mode)
        print(kernel, mode, np.isfinite(lr).sum(), float(np.nanmax(lr[np.isfinite(lr)])))
cfg3 = util.c3(seglen=16)
m3 = cfg3.cuda_model()
P3 = util.draw_bns(rng, 200, cfg3)
print(m3.plan.loglr_host(P3)[0][:3])
h, n = m3.plan.generate(P3[0])
print(n, abs(h).max())
