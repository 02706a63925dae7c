"""Walker sharding across the GPUs of one box.

Walkers of an ensemble half-step are independent: every rank holds a replica of
the data tables (the model), evaluates a contiguous slice of the proposal batch
and the per-walker results ``[W_local, 1 + n_stats]`` are exchanged with ONE
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

    def bounds(self, W):
        """Contiguous, balanced slices: rank r gets [lo_r, hi_r)."""
        base, rem = divmod(W, self.world_size)
        lo = [r * base + min(r, rem) for r in range(self.world_size + 1)]
        return lo

    def my_slice(self, W):
        lo = self.bounds(W)
        return lo[self.rank], lo[self.rank + 1]

    def all_gather_rows(self, local, W):
        """local: [W_local, C] float64 -> [W, C] on every rank (one collective)."""
        import torch
        lo = self.bounds(W)
        wmax = max(lo[r + 1] - lo[r] for r in range(self.world_size))
        C = local.shape[1]
        dev = self.device if self.device is not None else "cpu"
        buf = torch.zeros((wmax, C), dtype=torch.float64, device=dev)
        buf[:local.shape[0]] = torch.as_tensor(local, dtype=torch.float64, device=dev)
        out = torch.empty((self.world_size * wmax, C), dtype=torch.float64, device=dev)
        self._dist.all_gather_into_tensor(out, buf, group=self.group)
        out = out.cpu().numpy().reshape(self.world_size, wmax, C)
        return numpy.concatenate([out[r, :lo[r + 1] - lo[r]]
                                  for r in range(self.world_size)], axis=0)

    def evaluate(self, target, pts, callstat=None):
        """Sharded ``target.batch``: returns (values [W], stats dict name -> [W])
        identical on every rank."""
        W = len(pts)
        a, b = self.my_slice(W)
        res = target.batch(pts[a:b], callstat=callstat) if callstat is not None \
            else target.batch(pts[a:b])
        names = list(res.stats.keys())
        cols = [numpy.asarray(res.values, dtype=numpy.complex128)]
        cols += [numpy.asarray(res.stats[n], dtype=numpy.complex128) for n in names]
        packed = numpy.stack(cols, axis=1) if b > a else \
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
