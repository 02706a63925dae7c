import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import util
from gwin_b200 import _lib, distributions as D
from gwin_b200.sampler import CudaPTSampler
cfg = util.c3(); inj = cfg.inj
names = ['mass1', 'mass2', 'distance', 'inclination', 'coa_phase', 'ra', 'dec', 'polarization', 'tc']
prior = D.JointDistribution(names, D.Uniform(mass1=(1.2, 1.6), mass2=(1.2, 1.6)), D.UniformRadius(distance=(10., 100.)),
    D.SinAngle(inclination=None), D.UniformAngle(coa_phase=None), D.UniformSky(), D.UniformAngle(polarization=None),
    D.Uniform(tc=(inj['tc'] - 0.05, inj['tc'] + 0.05)))
model = cfg.cuda_model(variable_params=tuple(names), static_params={'spin1z': 0., 'spin2z': 0.}, prior=prior)
np.random.seed(1)
pt = CudaPTSampler(model, 8, 2048, seed=1)
pt.set_p0(); pt.run(1); torch.cuda.synchronize()
lib = pt._lib
import collections
acc = collections.defaultdict(float)
orig = {}
def wrap(name):
    f = getattr(lib, name)
    def g(*a):
        t = time.perf_counter(); r = f(*a); acc[name] += time.perf_counter() - t; return r
    return g
class L: pass
proxy = L()
for n in ('gwin_ens_propose', 'gwin_ens_prior', 'gwin_ens_accept', 'gwin_ens_swap', 'gwin_loglr_batch_strided', 'gwin_last_error'):
    setattr(proxy, n, wrap(n))
pt._lib = proxy
t0 = time.perf_counter(); pt.run(10); t1 = time.perf_counter() - t0
torch.cuda.synchronize(); t2 = time.perf_counter() - t0
print("run(10): returned after %.1f ms, all done %.1f ms" % (t1 * 1e3, t2 * 1e3))
for k, v in acc.items(): print("  %-28s %.2f ms total" % (k, v * 1e3))
pt._lib = lib
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    pt.run(3)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
print("cuda events:", len(evs))
prev_end = None
rows = []
for e in evs:
    s, en = e.time_range.start, e.time_range.end
    gap = (s - prev_end) if prev_end is not None else 0
    rows.append((e.name[:50], (en - s), gap))
    prev_end = max(prev_end or 0, en)
big = sorted(rows, key=lambda r: -r[2])[:12]
for r in big: print("gap %.1f us before %-50s (dur %.1f us)" % (r[2], r[0], r[1]))
tot_gap = sum(r[2] for r in rows); tot_dur = sum(r[1] for r in rows)
print("sum durations %.1f ms, sum gaps %.1f ms" % (tot_dur / 1e3, tot_gap / 1e3))
import collections
c = collections.defaultdict(float)
for r in rows: c[r[0]] += r[1]
for k, v in sorted(c.items(), key=lambda kv: -kv[1])[:10]: print("  %-50s %.2f ms" % (k, v / 1e3))
