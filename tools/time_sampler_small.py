"""C1-size device-resident ensemble sampler: iterations/s launched kernel by kernel vs replayed
as one CUDA graph per iteration (development probe)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import util
from gwin_b200 import distributions as D
from gwin_b200.sampler import CudaEnsembleSampler

nwalkers = int(sys.argv[1]) if len(sys.argv) > 1 else 200
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
cfg = util.c1()
inj = cfg.inj
names = ['tc', 'distance', 'inclination', 'coa_phase', 'ra', 'dec', 'polarization']
prior = D.JointDistribution(
    names, D.Uniform(tc=(inj['tc'] - 0.05, inj['tc'] + 0.05)), D.UniformRadius(distance=(50., 2000.)),
    D.SinAngle(inclination=None), D.UniformAngle(coa_phase=None), D.UniformSky(),
    D.UniformAngle(polarization=None))
static = {k: v for k, v in inj.items() if k not in names}
model = cfg.cuda_model(variable_params=tuple(names), static_params=static, prior=prior)
for graph in (False, True, False, True):
    np.random.seed(1)
    s = CudaEnsembleSampler(model, nwalkers, seed=3)
    s.set_p0()
    s.run(8, use_graph=graph)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run(niter, use_graph=graph)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("graph=%d  %d walkers x %d iterations: %.3f s  %.0f iterations/s  %.3e evals/s  chk %.6f"
          % (graph, nwalkers, niter, dt, niter / dt, nwalkers * niter / dt, float(s.chain[:, -1, :].sum())))
