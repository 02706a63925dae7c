"""C3 batch of 16384 walkers with coalescence times spread over +-dt around the injection, dt from the
bench's 0.05 s to the whole segment: what the table path costs when the lanes of a warp no longer share
table rows (windows wider than TAB_SPAN rows are read lane by lane from L2), and parity of every case
against the direct kernel on a sample.

  python tools/tc_spread_sweep.py [W]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench
import util

W = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cfg = util.c3()
model = cfg.cuda_model()
plan = model.plan
rng = np.random.default_rng(20182)
base = bench.draw_walkers(rng, W)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("tc spread [s]   ms/batch (cold L2)   evals/s     slow-path   max |auto - direct| (256 walkers)")
for dt in (0.05, 0.2, 1.0, 5.0, 30.0, 120.0):
    P = base.copy()
    centre = cfg.inj["tc"] if dt <= 5.0 else 0.5 * cfg.seglen
    P[:, 10] = centre + rng.uniform(-dt, dt, W)
    Pd = torch.tensor(P, device="cuda")
    plan.configure(kernel=0)
    out = plan.loglr_device(Pd)
    torch.cuda.synchronize()
    plan.slow_path_walkers()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    reps = 10
    for _ in range(reps):
        flush.fill_(1)
        e0.record()
        plan.loglr_device(Pd, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1) / reps
    n_slow = plan.slow_path_walkers()
    sel = rng.choice(W, 256, replace=False)
    auto = out[0].cpu().numpy()[sel]
    plan.configure(kernel=1)
    direct = plan.loglr_host(P[sel])[0]
    plan.configure(kernel=0)
    print("%10.2f %16.3f %16.3e %10d %16.2e" % (dt, ms, W / ms * 1e3, n_slow // reps, np.max(np.abs(auto - direct))))
