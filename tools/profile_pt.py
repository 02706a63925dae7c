"""Development aid: kernel timeline of a few C4 parallel-tempering iterations on this rank (torch.profiler /
CUPTI): per-kernel time, and the idle time between consecutive kernels on the device.

  [torchrun ...] python tools/profile_pt.py [niter]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import util
from gwin_b200 import distributions as D
from gwin_b200.sampler import CudaPTSampler

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
niter = int(sys.argv[1]) if len(sys.argv) > 1 else 8
graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
cfg = util.c3()
inj = cfg.inj
names = ['mass1', 'mass2', 'distance', 'inclination', 'coa_phase', 'ra', 'dec', 'polarization', 'tc']
prior = D.JointDistribution(
    names, D.Uniform(mass1=(1.2, 1.6)), D.Uniform(mass2=(1.2, 1.6)), D.UniformRadius(distance=(10., 100.)),
    D.SinAngle(inclination=None), D.UniformAngle(coa_phase=None), D.UniformAngle(ra=None), D.CosAngle(dec=None),
    D.UniformAngle(polarization=None), D.Uniform(tc=(inj['tc'] - 0.05, inj['tc'] + 0.05)))
static = {k: v for k, v in inj.items() if k not in names}
model = cfg.cuda_model(variable_params=tuple(names), static_params=static, prior=prior)
np.random.seed(3)
pt = CudaPTSampler(model, 8, 2048, seed=11)
pt.set_p0()
pt.run(4, use_graph=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pt.run(niter, use_graph=graph)
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t0, t1 = ev[0].time_range.start, ev[-1].time_range.end
    busy = sum(e.time_range.end - e.time_range.start for e in ev)
    print("%d device activities over %.1f us for %d iterations: busy %.1f us (%.0f %%), %.1f us per iteration"
          % (len(ev), t1 - t0, niter, busy, 100.0 * busy / (t1 - t0), (t1 - t0) / niter))
    agg = {}
    prev_end = None
    for e in ev:
        a = agg.setdefault(e.name[:48], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += e.time_range.end - e.time_range.start
        if prev_end is not None:
            a[2] += max(0.0, e.time_range.start - prev_end)
        prev_end = max(prev_end or 0, e.time_range.end)
    print("%-50s %6s %10s %14s" % ("kernel", "n/iter", "us/iter", "idle before us/iter"))
    for n, (c, t, g) in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
        print("%-50s %6.1f %10.1f %14.1f" % (n, c / niter, t / niter, g / niter))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
