"""C2 (BASELINE.json configs[1]): MarginalizedPhaseGaussianNoise, BBH injection, H1+L1+V1, 8 s @ 2048 Hz,
4096 walkers per batch on one GPU: device-resident and host-buffer timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import util
from gwin_b200 import models
cfg = util.c2()
model = cfg.cuda_model(cls=models.MarginalizedPhaseGaussianNoise)
for W in (200, 4096, 65536):
    rng = np.random.default_rng(1)
    P = util.draw_bbh(rng, W, cfg)
    mtot = P[:, 0] + P[:, 1]
    bins = np.clip(np.minimum((1 / (6 ** 1.5 * np.pi * mtot * 4.925491025543576e-06) / cfg.delta_f + 1).astype(int), model._kmax) - model._kmin, 0, None).mean()
    Pd = torch.tensor(np.ascontiguousarray(P), device="cuda")
    model.plan.configure(kernel=0, mchirp_min=8.0)
    out = model.plan.loglr_device(Pd, mode=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        model.plan.loglr_device(Pd, mode=1, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        model.plan.loglr_host(P, mode=1)
    hs = (time.perf_counter() - t0) / reps * 1e3
    print("C2 W=%6d mean bins %.0f: device-resident %.3f ms/batch (%.3e evals/s); host buffers %.3f ms/batch (%.3e evals/s); launches %d"
          % (W, bins, ms, W / ms * 1e3, hs, W / hs * 1e3, model.plan.last_launches))
