"""N>1 path on the CPU: world_size-2 gloo.  Walkers are sharded across ranks, each
rank evaluates its slice (oracle-backed model standing in for the GPU) and ONE
all_gather per half-step makes the results identical everywhere; the replicated
RNG then drives identical chains on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here):
        sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import util
    from gwin_b200 import distributions as D, models
    from gwin_b200.distributed import WalkerSharding
    from gwin_b200.pool import BatchPool
    from gwin_b200.sampler import EmceeEnsembleSampler, EmceePTSampler
    from oracle_model import OracleGaussianNoise
    cfg = util.c1()
    prior = D.JointDistribution(
        ['tc', 'distance'], D.Uniform(tc=(cfg.inj['tc'] - 0.05, cfg.inj['tc'] + 0.05)),
        D.UniformRadius(distance=(50., 1000.)))
    static = {k: v for k, v in cfg.inj.items() if k not in ('tc', 'distance')}
    m = OracleGaussianNoise(['tc', 'distance'], cfg, static_params=static, prior=prior, nthreads=1)
    shard = WalkerSharding()
    assert list(range(7))[shard.my_slice(7)] == ([0, 2, 4, 6], [1, 3, 5])[rank]
    assert shard.count(7) == (4, 3)[rank] and shard.count(1, rank=1) == 0
    pool = BatchPool(shard=shard)
    call = models.CallModel(m, 'logplr')
    rng = np.random.RandomState(0)      # same points on every rank
    pts = np.column_stack([cfg.inj['tc'] + rng.uniform(-0.06, 0.06, 9), rng.uniform(20., 900., 9)])
    sharded = pool.map(call, pts)
    local = BatchPool().map(call, pts)
    ok = all((a[0] == b[0]) and all((x == y) or (np.isnan(x) and np.isnan(y)) for x, y in zip(a[1], b[1]))
             for a, b in zip(sharded, local))
    np.random.seed(123)                 # replicated RNG
    s = EmceeEnsembleSampler(m, 20, pool=pool, model_call=call)
    s.set_p0()
    s.run(5)
    np.random.seed(321)
    pt = EmceePTSampler(m, 2, 20, pool=pool)
    pt.set_p0()
    pt.run(3)
    # option_utils.sampler_from_cli with use_mpi: the pool it builds shards over the ranks (reference
    # gwin/option_utils.py:150-185 with choose_pool(mpi=True)); same seed, same chain as above
    from types import SimpleNamespace
    from gwin_b200 import option_utils
    opts = SimpleNamespace(sampler="emcee", logpost_function="logplr", nprocesses=2, use_mpi=True, nwalkers=20)
    np.random.seed(123)
    s2 = option_utils.sampler_from_cli(opts, m)
    ok = ok and s2._pool.shard is not None and s2._pool.shard.world_size == world and s2._pool.count == 2
    s2.set_p0()
    s2.run(5)
    ok = ok and np.array_equal(s2.chain, s.chain)
    models._global_instance = None
    out.put((rank, ok, pool.count, s.chain.copy(), pt.chain.copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_walker_sharding_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([out.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1], "sharded map differs from the single-process map"
    assert res[0][2] == 2
    np.testing.assert_array_equal(res[0][3], res[1][3])     # identical chains on both ranks
    np.testing.assert_array_equal(res[0][4], res[1][4])
    # and identical to an un-sharded run with the same seed
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import util
    from gwin_b200 import distributions as D, models
    from gwin_b200.sampler import EmceeEnsembleSampler
    from oracle_model import OracleGaussianNoise
    cfg = util.c1()
    prior = D.JointDistribution(
        ['tc', 'distance'], D.Uniform(tc=(cfg.inj['tc'] - 0.05, cfg.inj['tc'] + 0.05)),
        D.UniformRadius(distance=(50., 1000.)))
    static = {k: v for k, v in cfg.inj.items() if k not in ('tc', 'distance')}
    m = OracleGaussianNoise(['tc', 'distance'], cfg, static_params=static, prior=prior, nthreads=1)
    np.random.seed(123)
    s = EmceeEnsembleSampler(m, 20, model_call=models.CallModel(m, 'logplr'))
    s.set_p0()
    s.run(5)
    np.testing.assert_array_equal(s.chain, res[0][3])
