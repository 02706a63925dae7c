#!/usr/bin/env python
"""bench.py -- loglr evals/sec of the batched TaylorF2 likelihood (BASELINE.json metric).

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): TaylorF2
BNS, 256 s @ 4096 Hz (524 289 frequency bins), H1+L1+V1, coloured Gaussian noise from the
analytic aLIGO ZDHP PSD plus a 1.4+1.4 Msun injection; GaussianNoise loglr for a batch of
16 384 walkers drawn in a box around the injection.

One "step" = one pass of the hot path over one batch of walkers.  With N GPUs every rank
holds a replica of the data tables and evaluates its own 16 384-walker slice (weak
scaling); the per-walker results [W, 1+3*n_det] are exchanged with one NCCL all-gather per
step, which is the only exchange the ensemble step needs.

  python bench.py [--gpus N] [--steps K] [--warmup W]           # CUDA arm
  python bench.py --impl reference [...]                        # CPU arm: the oracle port of the
                                                                # reference path on all host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "loglr evals/sec (TaylorF2, H1/L1/V1, 256 s @4096 Hz)"
UNIT = "evals/s"
SEGLEN, SRATE, F_LOWER = 256, 4096.0, 20.0
DETS = ["H1", "L1", "V1"]
WALKERS = 16384
MASS_BOX = (1.2, 1.6)                   # component masses of the walkers [Msun]
W_ALG_SLOTS = 32 + 12 * len(DETS)      # FP64 issue slots per (walker, bin): SURVEY.md 8d
TABLE_BYTES_PER_BIN = 16 * len(DETS) + 24
MTSUN = 4.925491025543575903411922162094833998e-6
PARAMS = ("mass1", "mass2", "spin1z", "spin2z", "distance", "inclination",
          "coa_phase", "ra", "dec", "polarization", "tc", "lambda1", "lambda2")
INJ = dict(mass1=1.4, mass2=1.4, spin1z=0.0, spin2z=0.0, distance=40.0,
           inclination=0.5, coa_phase=1.1, ra=1.3, dec=-0.4, polarization=0.7,
           tc=SEGLEN - 6.0, lambda1=0.0, lambda2=0.0)


def draw_walkers(rng, W):
    P = np.zeros((W, 13))
    P[:, 0] = rng.uniform(MASS_BOX[0], MASS_BOX[1], W)
    P[:, 1] = rng.uniform(MASS_BOX[0], MASS_BOX[1], W)
    P[:, 2] = 0.0
    P[:, 3] = 0.0
    P[:, 4] = rng.uniform(10.0, 100.0, W)
    P[:, 5] = np.arccos(rng.uniform(-1, 1, W))
    P[:, 6] = rng.uniform(0, 2 * np.pi, W)
    P[:, 7] = rng.uniform(0, 2 * np.pi, W)
    P[:, 8] = np.arcsin(rng.uniform(-1, 1, W))
    P[:, 9] = rng.uniform(0, 2 * np.pi, W)
    P[:, 10] = INJ["tc"] + rng.uniform(-0.05, 0.05, W)
    return P


def bins_used(P, kmin, kmax, delta_f):
    """Frequency bins each walker's inner products run over (per detector)."""
    fisco = 1.0 / (6.0 ** 1.5 * np.pi * (P[:, 0] + P[:, 1]) * MTSUN)
    n = (fisco / delta_f + 1).astype(np.int64)
    return np.clip(np.minimum(n, kmax) - kmin, 0, None)


# FP64 lane-operations per unit of work, counted in the source:
#   one (walker, sub-block) of the table kernel: Taylor coefficients 99, relative coefficients 21, the
#   series e_2..e_10 102, truncation estimate 12; per detector: grid position 10, weights 78, 13 taps x 8
#   complex moments 208, sum e_n M_n 28, anchor phase + sincos + accumulate 31
# This is the USEFUL work (what a walker's result needs).  What the launches EXECUTE is more: every lane of a
# warp runs a sub-block any of its lanes takes, and the taps x moments product runs on the FP64 tensor cores
# over the whole window of rows a warp's lanes touch (13 + their spread, zero weights elsewhere).  The executed
# count per launch comes from the committed ncu capture of this same command (profiles/r2_ncu_counts.json:
# FP64 warp instructions by opcode, a DMMA.884 = 8 multiply-adds per lane) as a ratio to the useful count.
C_SUB_SETUP, C_SUB_DET = 234, 355
C_RES_BIN = 14 + 33 * len(DETS)          # fine kernel: the bins after the finest sub-block, per-bin phase + sincos
C_REC_BIN = 6 + 8 * len(DETS) + 2        # recurrence kernel: 4 (cubic) / 8 (quartic) + 8 per detector, + block set-up


def ncu_counts():
    """Per launch of one step, from the committed capture: dict(kind -> kernel record) or {}."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_counts.json")))
    except (OSError, ValueError):
        return {}, None
    out, rec = {}, []
    for k in tj.get("kernels", []):
        nm = k["name"]
        if "loglr_table_kernel" in nm:
            out["table_fine" if "<3, 1" in nm else "table"] = k
        elif "loglr_recur_kernel" in nm:
            rec.append(k)
    # (with tables from kmin on there is no low-band launch: the one recurrence launch is the tails')
    if len(rec) == 1 and "table" in out:
        out["recurrence_tail"] = rec[0]
    elif rec:
        out["recurrence"] = rec[0]
        if len(rec) > 1:
            out["recurrence_tail"] = rec[1]
    return out, tj


def useful_work(plan, ke, band0):
    """Useful FP64 lane-operations of one batch per launch: dict(recurrence, table, table_fine, recurrence_tail),
    from the moment tables' schedule and the walkers' last bins."""
    start, m, _, fmin = plan.table_blocks()
    c_sub = C_SUB_SETUP + len(DETS) * C_SUB_DET
    out = dict(recurrence=0.0, table=0.0, table_fine=0.0, recurrence_tail=0.0)
    if len(start) == 0:
        out["recurrence"] = float(np.clip(ke - band0, 0, None).sum()) * C_REC_BIN
        return out
    k_split, end = int(start[0]), start + m
    out["recurrence"] = float(np.clip(np.minimum(ke, k_split) - band0, 0, None).sum()) * C_REC_BIN
    whole = np.searchsorted(end, ke, side="right")            # sub-blocks wholly below the last bin
    out["table"] = float(whole.sum()) * c_sub
    tail_s = np.where(whole < len(start), start[np.minimum(whole, len(start) - 1)], end[-1])
    tail_s = np.where(ke > k_split, tail_s, ke)
    left = np.clip(ke - tail_s, 0, None)
    fm = np.where(whole < len(start), fmin[np.minimum(whole, len(start) - 1)], 0)
    has_fine = (fm > 0) & (left > 0)
    units = np.where(has_fine, left // np.maximum(fm, 1), 0)
    out["table_fine"] = float(sum(bin(int(u)).count("1") for u in units)) * c_sub \
        + float(np.where(has_fine, left % np.maximum(fm, 1), 0).sum()) * C_RES_BIN
    out["recurrence_tail"] = float(np.where(has_fine, 0, left).sum()) * C_REC_BIN
    return out


def synthetic_noise(n_f, delta_f, seed=20181):
    from gwin_b200.psd import aLIGOZeroDetHighPower, colored_noise
    psd = aLIGOZeroDetHighPower(n_f, delta_f, F_LOWER)
    rng = np.random.default_rng(seed)
    return psd, {d: colored_noise(psd, delta_f, rng) for d in DETS}


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc = None
        self.device = device
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, power = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the power range
        pw = np.array(power)
        load = pw >= 0.5 * (pw.min() + pw.max()) if pw.max() > pw.min() else np.ones(len(pw), bool)
        return {"sm_mhz": float(np.median(np.array(sm)[load])), "sm_max_mhz": float(max(smax)),
                "power_w_max": float(pw.max()), "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# CPU arm / cpu_baseline: the oracle's C port of the reference path
# ---------------------------------------------------------------------------
def cpu_model(psd, data, delta_f):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gwin_oracle as O
    from c_oracle import COracle
    gen = O.FDomainDetFrameGenerator(0.0, DETS, delta_f, F_LOWER)
    wf = gen.generate(**INJ)
    d2 = {}
    for d in DETS:
        x = data[d].copy()
        n = min(len(wf[d]), len(x))
        x[:n] += wf[d][:n]
        d2[d] = x
    return COracle(d2, delta_f, 0.0, F_LOWER, psds={d: psd for d in DETS})


def cpu_time(co, P, threads, target_s):
    """Times a bounded sample of the workload; returns (evals/s, n_evaluated, seconds)."""
    probe = P[:max(threads, 8)]
    co.batch_loglr(probe, nthreads=threads)       # cold: page-faults the scratch arrays
    t0 = time.perf_counter()
    co.batch_loglr(probe, nthreads=threads)
    dt = time.perf_counter() - t0
    n = int(min(len(P), max(len(probe), target_s / (dt / len(probe)))))
    n = max(threads, n // threads * threads)
    t0 = time.perf_counter()
    co.batch_loglr(P[:n], nthreads=threads)
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_f = int(SEGLEN * SRATE) // 2 + 1
    delta_f = 1.0 / SEGLEN
    psd, data = synthetic_noise(n_f, delta_f)
    co = cpu_model(psd, data, delta_f)
    threads = os.cpu_count() or 1
    P = draw_walkers(np.random.default_rng(20182), WALKERS)
    # size one step to ~4 s of wall time
    rate, n, dt = cpu_time(co, P, threads, 4.0)
    per_step = n
    for _ in range(args.warmup):
        co.batch_loglr(P[:per_step], nthreads=threads)
    t0 = time.perf_counter()
    for s in range(args.steps):
        lo = (s * per_step) % (WALKERS - per_step + 1)
        co.batch_loglr(P[lo:lo + per_step], nthreads=threads)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = "%d of %d walkers per step, %d host threads, C port of the reference path (oracle/loglr_ref.c)" % (
        per_step, WALKERS, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config():
    """The same dict on both arms (the driver compares them): the workload is one step of 16 384 walkers per
    GPU; how much of it a CPU step samples is said in that line's ``cpu_baseline.sample``."""
    return {"workload": "C3: TaylorF2 BNS loglr (GaussianNoise), 256 s @ 4096 Hz, 524289 bins, "
                        "H1+L1+V1, coloured Gaussian noise + 1.4+1.4 Msun injection",
            "walkers_per_gpu_per_step": WALKERS, "f_lower_hz": F_LOWER,
            "mass_box_msun": list(MASS_BOX),
            "l2": "CUDA arm: flushed between timed steps (256 MiB write); CPU arm: tables far larger than the host caches",
            "kernel": "CUDA arm: plain model constructor, kernel AUTO (loglr_table_kernel over per-plan moment tables, "
                      "loglr_recur_kernel for whatever band the tables do not take); CPU arm: C port of the reference "
                      "path (oracle/loglr_ref.c) on all host threads, rank 0 only whatever N"}


# ---------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from gwin_b200 import _lib, models
    from gwin_b200.types import FrequencySeries
    from gwin_b200.waveform import FDomainCBCGenerator, FDomainDetFrameGenerator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stdout_fd = None
    if world > 1:
        # NCCL writes its version banner to file descriptor 1 when the communicator comes up;
        # stdout carries the one JSON line only, so fd 1 points at stderr until that line is printed
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    n_f = int(SEGLEN * SRATE) // 2 + 1
    delta_f = 1.0 / SEGLEN
    psd, noise = synthetic_noise(n_f, delta_f)
    gen = FDomainDetFrameGenerator(FDomainCBCGenerator, 0.0, detectors=DETS,
                                   variable_args=PARAMS, delta_f=delta_f, f_lower=F_LOWER,
                                   approximant="TaylorF2")
    inj = gen.generate(device=local, **INJ)       # CUDA generate_kernel
    data = {}
    for d in DETS:
        x = noise[d].copy()
        h = inj[d].numpy()
        x[:len(h)] += h[:n_f]
        data[d] = FrequencySeries(x, delta_f, 0.0)
    model = models.GaussianNoise(PARAMS, data, gen, F_LOWER, psds={d: psd for d in DETS},
                                 device=local)
    plan = model.plan          # plain constructor, kernel AUTO: nothing about the workload is told to the plan
    call = models.CallModel(model, "loglr", return_all_stats=False)

    W = args.walkers
    rng = np.random.default_rng(20182 + rank)
    import gwin_b200
    P_host = gwin_b200.pinned_empty((W, len(PARAMS)))      # the sampler's points live in page-locked host memory
    P_host[:] = draw_walkers(rng, W)
    bins = bins_used(P_host, plan.kmin, plan.kmax, delta_f)
    P_dev = torch.tensor(P_host, device=dev)
    nd = len(DETS)
    # one flat result buffer per rank, [loglr (W) | <d|h> (W x n_det x 2) | <h|h> (W x n_det)]: the kernels write
    # into its three segments and the all-gather sends it as it is (no packing copies)
    packed = torch.empty(W * (1 + 3 * nd), dtype=torch.float64, device=dev)
    out = (packed[:W], packed[W:W + 2 * nd * W].view(W, nd, 2), packed[W + 2 * nd * W:].view(W, nd))
    gathered = torch.empty((world, W * (1 + 3 * nd)), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        plan.loglr_device(P_dev, out=out)
        if world > 1:
            dist.all_gather_into_tensor(gathered, packed)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    plan.slow_path_walkers()                # reset the count: the timed steps must not use the slow path
    fp64_peak = _lib.fp64_peak(local)       # DFMA lane-ops/s, measured live on this GPU
    plan.kernel_timing(True)

    clocks = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    sync_all()
    for e0, e1 in ev:
        flush.fill_(1)                      # evict the tables from L2 between timed steps
        e0.record()
        step()
        e1.record()
    sync_all()
    clk = clocks.stop() if clocks else None
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    kern_ms, kern_n = plan.kernel_timing(False)
    kern_detail = plan.kernel_timing_detail()
    launches = plan.last_launches * args.steps
    n_slow = plan.slow_path_walkers()
    tab_info = plan.table_info()

    # end to end through the plugin call with HOST buffers (CallModel.batch -> C ABI host
    # entry: pinned staging, H2D of the parameter rows, kernels, D2H of loglr/hd/hh)
    # Two calls are kept in flight (CallModel.batch_async -> gwin_loglr_batch_points_submit / _wait), as a caller
    # with more than one batch per iteration does (the temperatures of a PT ensemble, pool.map over chunks): every
    # step still copies its own points in and its own loglr out inside the timed region, but the copies of step
    # n+1 overlap the kernels of step n.  The strictly serial figure (one call at a time) is reported beside it.
    call.batch(P_host)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = call.batch(P_host)
    e2e_sync_s = time.perf_counter() - t0
    assert np.all(np.isfinite(res.values))
    sync_all()
    P_alt = gwin_b200.pinned_empty((W, len(PARAMS)))        # two point sets alternate (nothing is cached)
    P_alt[:] = P_host[::-1]
    call.batch_async(P_host).result()
    sync_all()
    pend = None
    t0 = time.perf_counter()
    for i in range(args.steps):
        nxt = call.batch_async(P_alt if i & 1 else P_host)
        if pend is not None:
            res2 = pend.result()
        pend = nxt
    res2 = pend.result()
    e2e_s = time.perf_counter() - t0
    assert np.all(np.isfinite(res2.values))
    last = P_alt if (args.steps - 1) & 1 else P_host
    assert np.array_equal(res2.values, res.values[::-1] if last is P_alt else res.values)
    sync_all()

    t = torch.tensor([total_ms, e2e_s, float(bins.sum()), e2e_sync_s], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_s, e2e_sync_s = float(tmax[0]), float(tmax[1]), float(tmax[3])
        bins_total = float(tsum[2])
    else:
        bins_total = float(bins.sum())

    if rank == 0:
        evals = world * W * args.steps
        value = evals / (total_ms * 1e-3)
        e2e = evals / e2e_s
        # roofline (this rank): FP64 lane-operations the loglr launches execute / their own duration
        band0 = plan.kmin
        ke = np.minimum((1.0 / (6.0 ** 1.5 * np.pi * (P_host[:, 0] + P_host[:, 1]) * MTSUN) / delta_f + 1).astype(np.int64),
                        plan.kmax)
        work = useful_work(plan, ke, band0)
        names = {"recurrence": "loglr_recur_kernel<3,false> (band below the tables)",
                 "table": "loglr_table_kernel<3,false> (sub-blocks of the schedule)",
                 "table_fine": "loglr_table_kernel<3,true> (tails: fine sub-blocks + last bins)",
                 "recurrence_tail": "loglr_recur_kernel<3,false> (tails without fine sub-blocks)"}
        counts, counts_file = ncu_counts()
        executed = {}
        for k in work:
            c = counts.get(k)
            # executed / useful of this launch in the committed capture of the same (seeded) workload
            ratio = c["fp64_lane_ops_executed"] / c["useful_lane_ops"] if c and c.get("useful_lane_ops") else 1.0
            executed[k] = work[k] * ratio
        per_launch = []
        for k in ("recurrence", "table", "table_fine", "recurrence_tail"):
            if work[k] > 0 or kern_detail[k] > 2e-3:
                per_launch.append({"kernel": names[k], "ms": kern_detail[k], "fp64_lane_ops_executed": executed[k],
                                   "fp64_lane_ops_useful": work[k],
                                   "frac": executed[k] / max(kern_detail[k], 1e-9) / 1e-3 / fp64_peak,
                                   "useful_frac": work[k] / max(kern_detail[k], 1e-9) / 1e-3 / fp64_peak})
        lane_ops = sum(executed.values())
        achieved = 2.0 * lane_ops / (kern_ms * 1e-3) / 1e12       # TFLOP/s (1 lane-op = 2 flop)
        useful = 2.0 * sum(work.values()) / (kern_ms * 1e-3) / 1e12
        slots = float(bins.sum()) * W_ALG_SLOTS                   # the survey's work model, per launch
        peak = 2.0 * fp64_peak / 1e12
        nominal = 2.0 * 148 * 64 * 1.965e9 / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm = peaks.get("hbm_gbs", 6650.0)
        traffic, traffic_src = None, None
        if counts:       # dram bytes of the loglr launches of one step from the committed ncu --set full capture
            traffic = float(sum(c["dram_bytes_read"] + c["dram_bytes_write"] for c in counts.values()))
            traffic_src = ("profiles/r2_ncu_counts.json (%s; ncu --set full of the loglr launches of one timed step of `%s`) -- "
                           "reads = moment-table rows near the walkers' local times that are not in L2 after the flush (the "
                           "2.9 GB of tables are never streamed) plus the walkers' coefficient rows; writes = per-chunk partial "
                           "sums stay in L2 until the combine kernel reads them" % (counts_file["report"], counts_file["command"]))
        # every table byte has to come from HBM once per launch (L2 was flushed); the
        # 128 walker tiles that re-read it are served by the 126 MB L2
        alg_bytes = float(plan.kmax - plan.kmin) * TABLE_BYTES_PER_BIN + W * (13 + 1 + 3 * nd) * 8.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "detail": {"moment_tables": tab_info, "slow_path_walkers": n_slow, "mean_bins_per_eval": float(bins.mean()),
                       "walkers_this_run": W, "timed_region_ms": total_ms},
            "clocks": clk,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": world * W * 13 * 8,
                    "d2h_bytes_per_step": world * W * 8,
                    "path": "CallModel.batch_async(points) -> gwin_loglr_batch_points_submit / _wait, two calls in flight "
                            "(points in page-locked host memory; every step's H2D of its points, kernels and D2H of its "
                            "loglr inside the timed region; the copies of one step overlap the kernels of the other)",
                    "serial_value": evals / e2e_sync_s,
                    "serial_path": "CallModel.batch(points) -> gwin_loglr_batch_points_host, one call at a time"},
            "gpu_launches": launches,
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak,
                "peak_source": "live DFMA probe on this GPU (MEASURED_PEAKS.json has no FP64 entry); "
                               "nominal 148 SM x 64 lanes x 1.965 GHz = %.1f TFLOP/s" % nominal,
                "frac_of_nominal": achieved / nominal,
                "useful_achieved": useful, "useful_frac": useful / peak,
                "work_model": "frac: FP64 lane-operations the launches EXECUTE (warp instructions x 32 lanes, DMMA.884 = 8 "
                              "per lane; per launch = useful count x the executed/useful ratio of the committed ncu capture of "
                              "this command, profiles/r2_ncu_counts.json) / CUDA-event time / live DFMA peak.  useful_frac: only "
                              "what the walkers' results need, counted in the source: %d per (walker, sub-block) of the table "
                              "kernel (%d + %d per detector), %d per bin evaluated directly after the finest tail sub-block, %d "
                              "per bin of the recurrence kernel; units from the tables' schedule and the walkers' last bins"
                              % (C_SUB_SETUP + nd * C_SUB_DET, C_SUB_SETUP, C_SUB_DET, C_RES_BIN, C_REC_BIN),
                "launches": per_launch,
                "speedup_vs_work_model": 2.0 * slots / (kern_ms * 1e-3) / 1e12 / peak,
                "speedup_vs_work_model_note": "SURVEY.md 8d allots W_alg = 32 + 12*n_det = %d FP64 slots to every (walker, "
                                              "bin); this many times the live DFMA peak would be needed to do that work in "
                                              "the measured time" % W_ALG_SLOTS,
                "kernel": "the loglr launches of one step, timed as one bracket on the launching stream and launch by launch",
                "kernel_ms": kern_ms, "kernel_launches": kern_n,
                "kernel_share_of_step": kern_ms / (total_ms / args.steps),
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes,
                        "achieved_gbs": alg_bytes / (kern_ms * 1e-3) / 1e9,
                        "peak_gbs": hbm, "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / hbm},
                "traffic": traffic,
                "traffic_source": traffic_src,
            },
        }
        if world == 1 and not args.no_cpu:
            co = cpu_model(psd, noise, delta_f)
            threads = os.cpu_count() or 1
            rate, n, dt = cpu_time(co, P_host, threads, 12.0)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "%d of the %d walkers of one step in %.1f s on %d host threads; "
                          "C port of the reference path (oracle/loglr_ref.c), not pycbc/lalsimulation"
                          % (n, W, dt, threads)}
            # spot parity on the bench data itself: 256 walkers spanning the range of last bins
            sel = np.argsort(bins)[np.linspace(0, W - 1, 256).astype(int)]
            lr_c = co.batch_loglr(P_host[sel], nthreads=threads)[0]
            lr_g = res.values[sel]
            err = np.abs(lr_c - lr_g)
            line["parity_spot"] = {"walkers": len(sel), "max_abs_dloglr": float(err.max()),
                                   "max_over_tolerance": float(np.max(err / np.maximum(1e-6, 1e-9 * np.abs(lr_c)))),
                                   "against": "C port of the reference path (float64)"}
        if stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(stdout_fd, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--walkers", type=int, default=WALKERS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 1000 if args.impl == "cuda" else 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
