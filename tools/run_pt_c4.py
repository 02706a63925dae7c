"""C4: emcee_pt parallel tempering on the BNS config -- 8 temperatures x 2048 walkers, walkers
sharded across the GPUs of one box with one NCCL all-gather per half-step (BASELINE.json
configs[3]).  Launch with torchrun (one rank per GPU) or plain python for 1 GPU.

Reports likelihood evals/s through the *whole* sampler path (prior, proposal, sharded batched
likelihood, all-gather, accept, temperature swaps)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import util
from gwin_b200 import distributions as D, models
from gwin_b200.distributed import WalkerSharding
from gwin_b200.pool import BatchPool
from gwin_b200.sampler import CudaPTSampler, EmceePTSampler

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
pool = BatchPool()
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pool = BatchPool(shard=WalkerSharding(device=torch.device("cuda", local)))
ntemps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nwalkers = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
niter = int(sys.argv[3]) if len(sys.argv) > 3 else 5
seglen = int(sys.argv[4]) if len(sys.argv) > 4 else 256
resident = (sys.argv[5] if len(sys.argv) > 5 else "device") == "device"

cfg = util.c3(seglen=seglen)
inj = cfg.inj
names = ['mass1', 'mass2', 'distance', 'inclination', 'coa_phase', 'ra', 'dec', 'polarization', 'tc']
prior = D.JointDistribution(
    names, D.Uniform(mass1=(1.2, 1.6), mass2=(1.2, 1.6)), D.UniformRadius(distance=(10., 100.)),
    D.SinAngle(inclination=None), D.UniformAngle(coa_phase=None), D.UniformSky(),
    D.UniformAngle(polarization=None), D.Uniform(tc=(inj['tc'] - 0.05, inj['tc'] + 0.05)))
model = cfg.cuda_model(variable_params=tuple(names), static_params={'spin1z': 0., 'spin2z': 0.},
                       prior=prior, device=local)
np.random.seed(20184)                       # replicated RNG: identical decisions on every rank
pt = CudaPTSampler(model, ntemps, nwalkers, seed=4242) if resident else EmceePTSampler(model, ntemps, nwalkers, pool=pool)
pt.set_p0()
pt.run(1)                                   # warm-up (also evaluates the initial ensemble)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
n0 = pool.npoints
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
timing = bool(os.environ.get("PT_KERNEL_TIMING"))     # the harvest synchronises the host: off by default
if timing:
    model.plan.kernel_timing(True)
e0.record()
t0 = time.perf_counter()
pt.run(niter)
t_issue = time.perf_counter() - t0
e1.record()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
kms, kn = model.plan.kernel_timing(False) if timing else (0.0, 0)
if rank == 0:
    print("   wall %.1f ms, host issue %.1f ms, device span %.1f ms, loglr kernel %.3f ms x %d = %.1f ms"
          % (dt * 1e3, t_issue * 1e3, e0.elapsed_time(e1), kms, kn, kms * kn))
evals = ntemps * nwalkers * niter           # two half-steps of ntemps*nwalkers/2 points per iteration
chk = float(np.sum(pt.chain[:, :, -1, :]))      # after the timed region: histories are flushed lazily
if world > 1:
    t = torch.tensor([dt, chk], dtype=torch.float64, device="cuda")
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    same = bool(tmax[1] == tmin[1])
    dt = float(tmax[0])
else:
    same = True
if rank == 0:
    sw = pt.tswap_acceptance_fraction if resident else pt.sampler.tswap_acceptance_fraction
    print("C4 PT (%s stepper): %d GPUs, %d temps x %d walkers, %d iterations, T=%d s: %.3f s -> %.4e likelihood evals/s "
          "(whole sampler path); chains identical on all ranks: %s (checksum %.9f); cold-chain acceptance %.2f; swap acceptance %s"
          % ("device-resident" if resident else "host", world, ntemps, nwalkers, niter, seglen, dt, evals / dt, same,
             chk, pt.acceptance_fraction[0].mean(), np.round(sw, 2)))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
