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


def _usable_samples(path):
    """Number of samples per walker in a result file that reads back consistently, else 0."""
    from .io import InferenceFile, check_integrity
    try:
        check_integrity(path)
        with InferenceFile(path, 'r') as fp:
            first = '%s/%s' % (fp.samples_group, fp.variable_params[0])
            return int(fp[first].size)
    except Exception:        # noqa: BLE001 -- anything a truncated / foreign file can raise on the way in
        return 0


def validate_checkpoint_files(checkpoint_file, backup_file):
    """The reference's rule for resuming (``gwin/option_utils.py:192-283``): a file is usable when it
    passes ``check_integrity`` and holds at least one sample; a backup that holds a different number
    of samples than a usable checkpoint is not.  When exactly one of the two is usable it is copied
    over the other, so that afterwards both are usable or neither is.  Returns whether a run may
    resume from them."""
    import shutil
    n_chk, n_bak = _usable_samples(checkpoint_file), _usable_samples(backup_file)
    chk_ok, bak_ok = n_chk > 0, n_bak > 0
    if chk_ok and bak_ok and n_bak != n_chk:
        bak_ok = False                              # the checkpoint is taken to be the complete one
    if chk_ok and not bak_ok:
        shutil.copy(checkpoint_file, backup_file)
    elif bak_ok and not chk_ok:
        shutil.copy(backup_file, checkpoint_file)
    return chk_ok or bak_ok
