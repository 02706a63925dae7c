"""``BatchPool``: the object handed to a sampler as ``pool=``.

The reference picks a fork/MPI pool (``gwin/option_utils.py:179-182``) and the
sampler library calls ``pool.map(lnprobfn, [p_0, ..., p_{W-1}])`` once per
ensemble half-step, one pickled task per walker.  ``BatchPool.map`` keeps that
protocol (``.map`` returning a list, ``.count`` as kombine reads at
``sampler/kombine.py:81``) but, when the mapped callable knows how to evaluate a
batch (``CallModel.batch`` and the PT wrappers), evaluates the whole list in ONE
kernel launch and hands back exactly what W scalar calls would have returned.
Optionally the batch is sharded over the ranks of a ``torch.distributed``
process group (one GPU per rank) and all-gathered.
"""
import numpy

from . import models as _models


class BatchPool(object):

    def __init__(self, shard=None):
        """shard : ``gwin_b200.distributed.WalkerSharding`` or None."""
        self.shard = shard
        self.count = 1 if shard is None else shard.world_size
        self.ncalls = 0
        self.npoints = 0

    @staticmethod
    def _resolve(fn):
        """The object with a ``batch(points)`` method behind ``fn`` (or None).

        Looks through the two indirections a sampler library puts in front of the model:
        emcee's ``_function_wrapper`` (attributes ``f``, ``args``, ``kwargs``; only when no extra
        arguments are bound) and gwin's module-level ``_call_global_model`` used with fork pools
        (reference ``models/__init__.py:28-34``)."""
        if hasattr(fn, "f") and not getattr(fn, "args", None) and not getattr(fn, "kwargs", None):
            fn = fn.f
        if fn is _models._call_global_model:
            fn = _models._global_instance
        return fn if hasattr(fn, "batch") else None

    def map(self, fn, iterable):
        points = list(iterable)
        target = self._resolve(fn)
        if target is None or len(points) == 0:
            return [fn(p) for p in points]
        pts = numpy.asarray(points, dtype=numpy.float64)
        self.ncalls += 1
        self.npoints += len(pts)
        if self.shard is not None:
            return self.shard.batch_rows(target, pts)
        return batch_rows(target, pts)

    def close(self):
        pass


def batch_rows(target, pts):
    """Evaluate ``target.batch(pts)`` and unpack to the list ``map`` returns."""
    res = target.batch(pts)
    return_all = getattr(target, "return_all_stats", True)
    return res.rows(return_all_stats=return_all)
