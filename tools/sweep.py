"""C5: loglr throughput sweep -- batch 2^10..2^20 parameter points x segment 4..256 s @ 4096 Hz,
H1+L1+V1, BNS mass box.  CUDA-event timed, inputs resident on the device."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import util

segs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4, 8, 16, 32, 64, 128, 256]
Ws = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20]
print("%6s %9s %12s %12s %14s %10s" % ("T[s]", "W", "ms/batch", "evals/s", "bin-evals/s", "bins/eval"))
for T in segs:
    cfg = util.c3(seglen=T)
    model = cfg.cuda_model()
    for W in Ws:
        rng = np.random.default_rng(W + T)
        P = util.draw_bns(rng, W, cfg)
        fisco = 1.0 / (6 ** 1.5 * np.pi * (P[:, 0] + P[:, 1]) * 4.925491025543576e-06)
        bins = np.clip(np.minimum((fisco / cfg.delta_f + 1).astype(np.int64), model._kmax) - model._kmin, 0, None).mean()
        if W * bins > 6e11:
            continue
        model.plan.configure(kernel=0, mchirp_min=1.0)
        Pd = torch.tensor(P, device="cuda")
        out = model.plan.loglr_device(Pd)
        torch.cuda.synchronize()
        reps = 3 if W * bins > 1e10 else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            model.plan.loglr_device(Pd, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("%6d %9d %12.3f %12.4e %14.4e %10.0f" % (T, W, ms, W / (ms * 1e-3), W * bins / (ms * 1e-3), bins), flush=True)
        del Pd, out
    del model
    torch.cuda.empty_cache()
