"""Walker sharding across the GPUs of one box.

Walkers of an ensemble half-step are independent: every rank holds a replica of
the data tables (the model), evaluates every ``world_size``-th point of the
proposal batch (round-robin: a tempered ensemble is ordered by temperature, and
hot chains propose far more points outside the prior -- skipped, free -- than
cold ones, so contiguous slices would be unbalanced) and the per-walker results ``[W_local, 1 + n_stats]`` are exchanged with ONE
``all_gather`` per half-step (NCCL over NVLink on GPUs, gloo in the CPU tests).
Every rank then holds identical results and can take the identical accept/reject
decisions from a replicated RNG -- positions need no traffic.
"""
import numpy


class WalkerSharding(object):

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        self.device = device

    def my_slice(self, W):
        """Round-robin share of W points: ``pts[my_slice(W)]`` = points rank, rank + world, ..."""
        return slice(self.rank, W, self.world_size)

    def count(self, W, rank=None):
        rank = self.rank if rank is None else rank
        return max(0, -(-(W - rank) // self.world_size))

    def all_gather_rows(self, local, W):
        """local: [W_local, C] float64 (rows of ``my_slice``) -> [W, C] in the original
        order on every rank (one collective)."""
        import torch
        wmax = -(-W // self.world_size)
        C = local.shape[1]
        dev = self.device if self.device is not None else "cpu"
        buf = torch.zeros((wmax, C), dtype=torch.float64, device=dev)
        buf[:local.shape[0]] = torch.as_tensor(local, dtype=torch.float64, device=dev)
        out = torch.empty((self.world_size * wmax, C), dtype=torch.float64, device=dev)
        self._dist.all_gather_into_tensor(out, buf, group=self.group)
        # row j of rank r is point j * world + r
        out = out.cpu().numpy().reshape(self.world_size, wmax, C)
        return numpy.ascontiguousarray(out.transpose(1, 0, 2)).reshape(-1, C)[:W]

    def evaluate(self, target, pts, callstat=None):
        """Sharded ``target.batch``: returns (values [W], stats dict name -> [W])
        identical on every rank."""
        W = len(pts)
        mine = numpy.asarray(pts)[self.my_slice(W)]
        res = target.batch(mine, callstat=callstat) if callstat is not None \
            else target.batch(mine)
        names = list(res.stats.keys())
        cols = [numpy.asarray(res.values, dtype=numpy.complex128)]
        cols += [numpy.asarray(res.stats[n], dtype=numpy.complex128) for n in names]
        packed = numpy.stack(cols, axis=1) if len(mine) else \
            numpy.zeros((0, 1 + len(names)), dtype=numpy.complex128)
        # complex stats (e.g. {det}_cplx_loglr) travel as (re, im) pairs
        local = numpy.concatenate([packed.real, packed.imag], axis=1)
        full = self.all_gather_rows(local, W)
        nc = 1 + len(names)
        re, im = full[:, :nc], full[:, nc:]
        values = re[:, 0]
        stats = {}
        for i, n in enumerate(names):
            is_cplx = numpy.iscomplexobj(res.stats[n])
            stats[n] = re[:, 1 + i] + 1j * im[:, 1 + i] if is_cplx else re[:, 1 + i]
        return values, stats, names

    def batch_rows(self, target, pts):
        values, stats, names = self.evaluate(target, pts)
        vals = values.tolist()
        if not getattr(target, "return_all_stats", True):
            return vals
        cols = [numpy.asarray(stats[n]).tolist() for n in names]
        blobs = list(zip(*cols)) if cols else [()] * len(vals)
        return list(zip(vals, blobs))
