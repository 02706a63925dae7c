"""Selecting the sampler and its pool from parsed options -- the one piece of
``gwin/option_utils.py`` on the hot path (``sampler_from_cli``, :150-185).

The reference wraps the model in ``CallModel(model, opts.logpost_function)``, parks it in
``models._global_instance`` when ``opts.nprocesses > 1`` (so that fork/MPI workers call it by
name), asks pycbc's ``choose_pool`` for a pool and hands everything to the sampler class'
``from_cli``.  Here the pool is always a ``BatchPool``: whatever ``nprocesses`` says, one
``pool.map`` becomes one batched likelihood call on this rank's GPU, and when
``torch.distributed`` is up (``opts.use_mpi`` in the reference's vocabulary: one process per
GPU) the batch is sharded over the ranks.  The argument parser, config files and data loading of
that module are outside the path (out of scope).
"""
from . import models, sampler
from .pool import BatchPool


def choose_pool(mpi=False, processes=1, device=None):
    """Counterpart of ``pycbc.pool.choose_pool`` as the reference calls it (:179): a ``BatchPool``
    whose ``count`` is what kombine reads (``sampler/kombine.py:81``); sharded over the
    ``torch.distributed`` ranks when asked for (``mpi``) and the process group exists."""
    shard = None
    if mpi:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from .distributed import WalkerSharding
            shard = WalkerSharding(device=device)
    pool = BatchPool(shard=shard)
    pool.count = max(int(processes), pool.count)
    return pool


def sampler_from_cli(opts, model, pool=None):
    """``opts`` needs what the reference's parser provides: ``sampler``, ``logpost_function``,
    ``nprocesses``, ``use_mpi`` and the chosen sampler's own options (``nwalkers``, ``ntemps``...).
    Returns the sampler instance; as in the reference the model reaches it wrapped in a
    ``CallModel`` and, with ``nprocesses > 1``, through ``models._call_global_model``."""
    model = models.CallModel(model, opts.logpost_function)
    if opts.nprocesses > 1:
        models._global_instance = model
        model_call = models._call_global_model
    else:
        model_call = None
    sclass = sampler.samplers[opts.sampler]
    if pool is None:
        pool = choose_pool(mpi=getattr(opts, "use_mpi", False), processes=opts.nprocesses)
    pool.count = max(int(opts.nprocesses), pool.count)
    return sclass.from_cli(opts, model, pool=pool, model_call=model_call)
