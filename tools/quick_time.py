"""Development timing probe (not the bench): C3-style batch, CUDA events.

  python tools/quick_time.py [W] [seglen] [kernels]     kernels: direct,recurrence,table,auto
  MCW=<rel>   narrow batch: chirp masses within +-rel of the injection's
  SPIN=<x>    aligned spins uniform in +-x;  LAM=<x> tidal deformabilities up to x
  CALIB=<n>   calibrated kernels with n spline points per detector
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import util
from gwin_b200 import _lib

W = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
seglen = int(sys.argv[2]) if len(sys.argv) > 2 else 256
kernels = sys.argv[3].split(",") if len(sys.argv) > 3 else ["direct", "recurrence"]
t0 = time.time()
cfg = util.c3(seglen=seglen)
model = cfg.cuda_model()
print("lib %s setup %.1fs  kmin %d kmax %d" % (os.path.basename(_lib.LIB_PATH), time.time() - t0, model._kmin, model._kmax))
rng = np.random.default_rng(20182)
P = util.draw_bns(rng, W, cfg)
mcw = float(os.environ.get("MCW", "0"))
if mcw:
    ref = util.inj_row(cfg)
    f = 1.0 + mcw * rng.uniform(-1, 1, W)
    P[:, 0], P[:, 1] = ref[0] * f, ref[1] * f * (1.0 + 0.01 * rng.uniform(-1, 1, W))
spin = float(os.environ.get("SPIN", "0"))
if spin:
    P[:, 2:4] = rng.uniform(-spin, spin, (W, 2))
lam = float(os.environ.get("LAM", "0"))
if lam:
    P[:, 11:13] = rng.uniform(0, lam, (W, 2))
n = np.array([min(int(1.0 / (6 ** 1.5 * np.pi * (a + b) * 4.925491025543576e-06) / cfg.delta_f + 1), model._kmax) - model._kmin
              for a, b in P[:, :2]])
bins = n.mean()
Pd = torch.tensor(P, device="cuda")
ncal = int(os.environ.get("CALIB", "0"))
Cd = None
if ncal:
    sf = np.tile(np.logspace(np.log10(cfg.f_lower), np.log10(2048.), ncal), (len(cfg.dets), 1))
    model.plan.set_calibration(sf)
    Cd = torch.tensor(rng.normal(0, 0.05, (W, len(cfg.dets) * 2 * ncal)), device="cuda")
res = {}
for name in kernels:
    kid = {"direct": 1, "recurrence": 2, "table": 3, "auto": 0}[name]
    model.plan.configure(kernel=kid, phase_tol=float(os.environ.get("PHASE_TOL", "0")))
    t0 = time.time()
    out = model.plan.loglr_device(Pd, calib=Cd)
    torch.cuda.synchronize()
    print("%s: first call %.2f s, launches %d, tables %s, slow-path walkers %d"
          % (name, time.time() - t0, model.plan.last_launches, model.plan.table_info(), model.plan.slow_path_walkers()))
    res[name] = out[0].cpu().numpy()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        model.plan.loglr_device(Pd, out=out, calib=Cd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # cold-L2 variant (what bench.py times): flush between reps
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cold = 0.0
    model.plan.kernel_timing(True)
    for _ in range(reps):
        flush.fill_(1)
        e0.record()
        model.plan.loglr_device(Pd, out=out, calib=Cd)
        e1.record()
        torch.cuda.synchronize()
        cold += e0.elapsed_time(e1) / reps
    kms, kn = model.plan.kernel_timing(False)
    print("   cold-L2 %.3f ms/batch (loglr launches %.3f ms)" % (cold, kms))
    evs = W / (ms * 1e-3)
    print("%-10s W=%d T=%d  %.3f ms/batch  %.3e evals/s  %.3e bin-evals/s  slow-path walkers %d"
          % (name, W, seglen, ms, evs, evs * bins, model.plan.slow_path_walkers()))
names = list(res)
for a in range(len(names)):
    for b in range(a + 1, len(names)):
        print("max |%s - %s| = %.3e" % (names[a], names[b], np.abs(res[names[a]] - res[names[b]]).max()))
