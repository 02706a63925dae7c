"""Device-resident ensemble / parallel-tempering samplers.

Same step as ``EmceeEnsembleSampler`` / ``EmceePTSampler`` (emcee 2.x semantics,
SURVEY.md appendix B) but the walkers never leave the GPU: proposal, boundary
conditions + prior, likelihood, accept and temperature swaps are kernels of
``libgwin_b200.so`` chained on one stream; with several GPUs each rank evaluates
its slice of the proposals and ONE NCCL all-gather of the log-likelihoods per
half-step (on device buffers) is the only communication -- the counter-based
RNG regenerates identical proposals and decisions on every rank.

torch is used for what the task allows it: owning device memory, the stream and
``torch.distributed``.  Restrictions (use the host steppers otherwise): priors
must be a ``JointDistribution`` of Uniform / UniformAngle / SinAngle / CosAngle /
UniformRadius / UniformSky / Gaussian without constraints, at most 64 sampled
dimensions; the only sampling transform understood is (mass1, mass2) -> (mchirp, q)
(log-Jacobian on the device), the only waveform transform (mchirp, q) -> (mass1, mass2).  ``recalib_*`` parameters of a calibrated model
(``gwin/calibration.py``) are sampled like any other: the prior kernel fills the
kernels' calibration matrix (values are not checked against the regime in which the
least-squares cubic equals scipy's smoothing spline -- keep their priors inside
sum(values^2) <= n_points per detector and quantity, as ``calibration.values_matrix`` demands on the host path).
"""
import ctypes as C
import os

import numpy

from .. import _lib
from .. import distributions as D
from .. import transforms as T
from ..distributions import FieldArray
from ..waveform import WAVEFORM_PARAMS
from .base import BaseMCMCSampler
from .ensemble import default_beta_ladder

_TYPES = {D.Uniform: _lib.PRIOR_UNIFORM, D.UniformAngle: _lib.PRIOR_UNIFORM_ANGLE,
          D.SinAngle: _lib.PRIOR_SIN_ANGLE, D.CosAngle: _lib.PRIOR_COS_ANGLE,
          D.UniformRadius: _lib.PRIOR_UNIFORM_RADIUS, D.Gaussian: _lib.PRIOR_GAUSSIAN}


def prior_spec_from_model(model):
    """Translate ``model.prior_distribution`` + static params into the
    ``gwin_prior_spec`` the device kernels understand."""
    sampled_mchirp_q = False
    if model.sampling_transforms is not None:
        st = model.sampling_transforms.sampling_transforms
        if len(st) == 1 and isinstance(st[0], T.Mass1Mass2ToMchirpQ):
            sampled_mchirp_q = True             # the walkers move in (mchirp, q), the prior is on the masses
        else:
            raise NotImplementedError("the only sampling transform supported on the device path is "
                                      "mass1, mass2 -> mchirp, q")
    wt = getattr(model, "_waveform_transforms", None)
    mchirp_q = False
    if wt:
        if len(wt) == 1 and isinstance(wt[0], T.MchirpQToMass1Mass2):
            mchirp_q = True
        else:
            raise NotImplementedError("only the mchirp,q -> mass1,mass2 waveform transform "
                                      "is supported on the device path")
    prior = model.prior_distribution
    if not isinstance(prior, D.JointDistribution) or prior._constraints:
        raise NotImplementedError("the device path needs a JointDistribution without constraints")
    per_param = {}
    for dist in prior.distributions:
        parts = [dist._az, dist._po] if isinstance(dist, D.UniformSky) else [dist]
        for part in parts:
            if type(part) not in _TYPES:
                raise NotImplementedError("prior %s is not supported on the device path"
                                          % type(part).__name__)
            for name, b in part.bounds.items():
                per_param[name] = (_TYPES[type(part)], b.min, b.max)
                if isinstance(part, D.Gaussian):
                    per_param[name] += (part._mean[name], float(numpy.sqrt(part._var[name])))
    names = list(model.sampling_params)
    if len(names) > _lib.ENS_MAXDIM:
        raise NotImplementedError("more than %d sampling dimensions" % _lib.ENS_MAXDIM)
    wf_names = list(WAVEFORM_PARAMS.keys())
    # recalib_* parameters: value v = (2 det + which) n_points + i of the kernels' calibration matrix
    calib_names = []
    recalib = getattr(model, "_recalib", None)
    if recalib:
        for det in model.detectors:
            calib_names += list(recalib[det].param_names)
    spec = _lib.PriorSpec()
    spec.ndim = len(names)
    if sampled_mchirp_q:
        if mchirp_q:
            raise NotImplementedError("mchirp, q as both a sampling and a waveform transform")
        spec.derived_masses = 1
        for i, name in enumerate(("mass1", "mass2")):
            kind, lo, hi = per_param[name][:3]
            spec.derived_type[i], spec.derived_lo[i], spec.derived_hi[i] = kind, lo, hi
            spec.derived_sigma[i] = 1.0
            if kind == _lib.PRIOR_GAUSSIAN:
                spec.derived_mu[i], spec.derived_sigma[i] = per_param[name][3:5]
        # the sampling parameters themselves: no prior of their own, wide bounds for the coverage draw
        m1, m2 = per_param["mass1"], per_param["mass2"]
        mc = lambda a, b: (a * b) ** 0.6 / (a + b) ** 0.2
        per_param["mchirp"] = (_lib.PRIOR_NONE, mc(m1[1], m2[1]), mc(m1[2], m2[2]))
        per_param["q"] = (_lib.PRIOR_NONE, m1[1] / m2[2], m1[2] / m2[1])
    for d, name in enumerate(names):
        kind, lo, hi = per_param[name][:3]
        spec.type[d], spec.lo[d], spec.hi[d] = kind, lo, hi
        spec.sigma[d] = 1.0
        if kind == _lib.PRIOR_GAUSSIAN:
            spec.mu[d], spec.sigma[d] = per_param[name][3:5]
        if name in calib_names:
            spec.target[d] = _lib.TARGET_CALIB0 + calib_names.index(name)
        elif name in wf_names:
            spec.target[d] = wf_names.index(name)
        elif (mchirp_q or sampled_mchirp_q) and name == "mchirp":
            spec.target[d] = _lib.TARGET_MCHIRP
        elif (mchirp_q or sampled_mchirp_q) and name == "q":
            spec.target[d] = _lib.TARGET_Q
        else:
            raise NotImplementedError("sampling parameter %r does not map to a waveform parameter" % name)
    static = dict(model.waveform_generator.frozen_params)
    static.update(model.static_params)
    for i, (name, default) in enumerate(WAVEFORM_PARAMS.items()):
        if name in static:
            spec.fixed[i] = float(static[name])
        elif default is not None:
            spec.fixed[i] = float(default)
        elif name in names or ((mchirp_q or sampled_mchirp_q) and name in ("mass1", "mass2")):
            spec.fixed[i] = 0.0
        else:
            raise ValueError("waveform parameter %r is neither sampled nor static" % name)
    return spec


def _min_chirp_mass(spec):
    """Smallest chirp mass inside the prior box (None if it cannot be told)."""
    lo = {}
    for d in range(spec.ndim):
        lo[spec.target[d]] = spec.lo[d]
    if _lib.TARGET_MCHIRP in lo:
        return lo[_lib.TARGET_MCHIRP] if lo[_lib.TARGET_MCHIRP] > 0 else None
    m1 = lo.get(0, spec.fixed[0])
    m2 = lo.get(1, spec.fixed[1])
    if m1 > 0 and m2 > 0:
        return (m1 * m2) ** 0.6 / (m1 + m2) ** 0.2
    return None


class _DeviceEnsemble(BaseMCMCSampler):
    """Common engine: ``ntemps`` tempered ensembles of ``nwalkers`` walkers."""

    def __init__(self, model, ntemps, nwalkers, betas, move, interleaved, seed=None,
                 process_group=None, a=2.0):
        import torch
        if nwalkers % 2:
            raise ValueError("The number of walkers must be even.")
        ndim = len(model.sampling_params)
        if nwalkers < 2 * ndim:
            raise ValueError("The number of walkers needs to be more than twice "
                             "the dimension of your parameter space.")
        self._torch = torch
        self._lib = _lib.load()
        self._spec = prior_spec_from_model(model)
        super(_DeviceEnsemble, self).__init__(None, model)
        self._nwalkers, self._ntemps, self._ndim = nwalkers, ntemps, ndim
        self._move, self._interleaved, self._a = move, interleaved, float(a)
        self._seed = int(numpy.random.randint(0, 2 ** 31 - 1) if seed is None else seed)
        self._plan = model.plan
        self._dev = torch.device("cuda", self._plan.device)
        self._betas_host = numpy.asarray(betas, dtype=float)
        f64 = dict(dtype=torch.float64, device=self._dev)
        self._betas = torch.tensor(self._betas_host, **f64)
        nh = nwalkers // 2
        self._Wh = ntemps * nh
        self._q = torch.empty((self._Wh, ndim), **f64)
        self._logfac = torch.empty(self._Wh, **f64)
        self._q_logp = torch.empty(self._Wh, **f64)
        self._skip = torch.empty(self._Wh, dtype=torch.uint8, device=self._dev)
        self._params = torch.empty((_lib.GWIN_NPARAM, self._Wh), **f64)
        # calibrated model: the recalib_* values of the proposals, value-major [n_values][W]; values that
        # are static keep what is written here once
        self._calib_names = []
        recalib = getattr(model, "_recalib", None)
        if recalib:
            for det in model.detectors:
                self._calib_names += list(recalib[det].param_names)
        self._calib = None
        if self._calib_names:
            self._calib = self._calib_buffer(self._Wh)
        self._naccept = torch.zeros((ntemps, nwalkers), dtype=torch.int32, device=self._dev)
        self._nswap_acc = torch.zeros(ntemps, dtype=torch.int32, device=self._dev)
        self._nswap = 0
        # walker sharding
        self._group = process_group
        self._world, self._rank = 1, 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self._world = torch.distributed.get_world_size(process_group)
            self._rank = torch.distributed.get_rank(process_group)
        self._wmax = -(-self._Wh // self._world)
        self._gath = torch.empty(self._world * self._wmax, **f64)
        self._local = torch.zeros(self._wmax, **f64)
        self._skip_l = torch.zeros(self._wmax, dtype=torch.uint8, device=self._dev)
        self._p = None
        self._iterations = 0
        self._iter0 = 0              # RNG counter offset (resume / cleared chains)
        self._chain = self._lnprob_h = self._logl_h = None
        self._pending = []
        self.likelihood_evals = 0
        # the block schedule of the recurrence kernel must cover the smallest chirp mass the
        # prior allows (the device entry point cannot look at the batch without a sync)
        mc = _min_chirp_mass(self._spec)
        if mc is not None:
            self._plan.configure(mchirp_min=2.0 ** (numpy.floor(4.0 * numpy.log2(mc)) / 4.0))

    # -- emcee-like attributes --------------------------------------------
    @property
    def iterations(self):
        return self._iterations

    @property
    def niterations(self):
        return self._iterations + self.lastclear

    @property
    def ntemps(self):
        return self._ntemps

    @property
    def betas(self):
        return self._betas_host

    @property
    def acceptance_fraction(self):
        return self._naccept.cpu().numpy() / max(self._iterations, 1)

    @property
    def tswap_acceptance_fraction(self):
        """Accepted / attempted swaps per temperature (end temperatures take part in one swap
        per iteration, inner ones in two)."""
        per_iter = numpy.full(self._ntemps, 2.0)
        per_iter[0] = per_iter[-1] = 1.0
        return self._nswap_acc.cpu().numpy() / numpy.maximum(self._nswap * per_iter, 1)

    def _hist(self, t):
        return numpy.empty((self._ntemps, self._nwalkers, 0) + ((self._ndim,) if t else ())) \
            if self._chain is None else None

    def _eval_local(self, W, q, q_logp, skip, params, out_loglr):
        """prior for W points in q and the likelihood of this rank's share of them: all W into
        out_loglr on one GPU, else points rank, rank + world, ... into ``self._local``."""
        torch, lib = self._torch, self._lib
        st = torch.cuda.current_stream(self._dev).cuda_stream
        calib = None
        if self._calib_names:
            calib = self._calib if W == self._calib.shape[1] else self._calib_buffer(W)
        _lib.check(lib.gwin_ens_prior_calib(C.byref(self._spec), W, q.data_ptr(), q_logp.data_ptr(),
                                            skip.data_ptr(), params.data_ptr(),
                                            calib.data_ptr() if calib is not None else None,
                                            len(self._calib_names), st))
        if self._world == 1 and calib is not None:
            _lib.check(lib.gwin_loglr_batch_calib(
                self._plan._h, self.model._mode, W, params.data_ptr(), 1, W, calib.data_ptr(), 1, W,
                skip.data_ptr(), out_loglr.data_ptr(), None, None, st))
        elif self._world == 1:
            _lib.check(lib.gwin_loglr_batch_strided(
                self._plan._h, self.model._mode, W, params.data_ptr(), 1, W, skip.data_ptr(),
                out_loglr.data_ptr(), None, None, st))
        else:
            # round-robin shard: this rank evaluates points rank, rank + world, ... (hot chains
            # propose many more points outside the prior than cold ones; contiguous slices of a
            # temperature-ordered batch would be unbalanced).  params is [13][W]: the strided
            # entry point walks it in place; the skip mask is compacted.
            world, rank = self._world, self._rank
            wmax = -(-W // world)
            n = max(0, -(-(W - rank) // world))
            local = self._local[:wmax]
            if n > 0:
                self._skip_l[:n].copy_(skip[rank:W:world])
                if calib is not None:
                    _lib.check(lib.gwin_loglr_batch_calib(
                        self._plan._h, self.model._mode, n, params.data_ptr() + 8 * rank, world, W,
                        calib.data_ptr() + 8 * rank, world, W,
                        self._skip_l.data_ptr(), local.data_ptr(), None, None, st))
                else:
                    _lib.check(lib.gwin_loglr_batch_strided(
                        self._plan._h, self.model._mode, n, params.data_ptr() + 8 * rank, world, W,
                        self._skip_l.data_ptr(), local.data_ptr(), None, None, st))
        self.likelihood_evals += W

    def _calib_buffer(self, W):
        """[n_values][W] with the static recalib_* values filled in (the sampled ones are written by the
        prior kernel for every batch)."""
        torch = self._torch
        static = dict(self.model.waveform_generator.frozen_params)
        static.update(self.model.static_params)
        buf = torch.zeros((len(self._calib_names), W), dtype=torch.float64, device=self._dev)
        sampled = set(self.model.sampling_params)
        for v, name in enumerate(self._calib_names):
            if name in sampled:
                continue
            if name not in static:
                raise KeyError(name)
            buf[v] = float(static[name])
        return buf

    def _gather(self, W):
        """The one collective of a half-step: every rank's share of the log-likelihoods."""
        wmax = -(-W // self._world)
        self._torch.distributed.all_gather_into_tensor(self._gath[:self._world * wmax], self._local[:wmax],
                                                       group=self._group)

    def _scatter_back(self, W, out_loglr):
        # entry j of rank r is point j * world + r
        world = self._world
        wmax = -(-W // world)
        out_loglr[:W].copy_(self._gath[:world * wmax].view(world, wmax).t().reshape(-1)[:W])

    def _eval_q(self, W, q, q_logp, skip, params, out_loglr):
        """prior + sharded likelihood for W points in q; result in out_loglr[:W]."""
        self._eval_local(W, q, q_logp, skip, params, out_loglr)
        if self._world > 1:
            self._gather(W)
            self._scatter_back(W, out_loglr)

    def set_p0(self, samples_file=None, prior=None):
        nt, nw = self._ntemps, self._nwalkers
        if samples_file is not None:
            samples = samples_file.read_samples(self.sampling_params, iteration=-1,
                                                flatten=False, group='sampling_samples')
        else:
            samples = self.model.prior_rvs(size=nw * nt, prior=prior)
        self._p = None
        p0 = numpy.ones((nt, nw, self._ndim))
        for i, param in enumerate(self.sampling_params):
            p0[..., i] = numpy.asarray(samples[param]).reshape((nt, nw))
        self._p0 = p0 if nt > 1 or self._keep_temp_axis else p0[0]
        return self._p0

    _keep_temp_axis = True
    # one CUDA graph per iteration (see run()); GWIN_SAMPLER_GRAPH=0 turns it off
    use_graph = os.environ.get("GWIN_SAMPLER_GRAPH", "1") != "0"

    def _initialise(self, p0):
        torch = self._torch
        nt, nw, nd = self._ntemps, self._nwalkers, self._ndim
        p = torch.tensor(numpy.asarray(p0, dtype=float).reshape(nt, nw, nd), dtype=torch.float64,
                         device=self._dev).contiguous()
        W = nt * nw
        f64 = dict(dtype=torch.float64, device=self._dev)
        logp = torch.empty(W, **f64)
        skip = torch.empty(W, dtype=torch.uint8, device=self._dev)
        params = torch.empty((_lib.GWIN_NPARAM, W), **f64)
        loglr = torch.empty(W, **f64)
        if self._world > 1 and self._gath.numel() < self._world * (-(-W // self._world)):
            self._gath = torch.empty(self._world * (-(-W // self._world)), **f64)
            self._local = torch.zeros(-(-W // self._world), **f64)
            self._skip_l = torch.zeros(-(-W // self._world), dtype=torch.uint8, device=self._dev)
        self._declare_coverage()
        self._eval_q(W, p.view(W, nd), logp, skip, params, loglr)
        logl = torch.where(torch.isneginf(logp), torch.zeros_like(loglr), loglr).view(nt, nw)
        lnprob = logl * self._betas.view(-1, 1) + logp.view(nt, nw)
        if bool(torch.isposinf(lnprob).any()) or bool(torch.isnan(lnprob).any()):
            raise ValueError("The initial lnprob was NaN or +inf.")
        self._p, self._logl, self._lnprob = p, logl.contiguous(), lnprob.contiguous()

    def _declare_coverage(self, n=4096):
        """The moment tables of the likelihood must cover what the sampler can propose: points
        spread over the prior box go through the device prior kernel (boundary conditions,
        mchirp/q -> masses) and are declared to the plan (``gwin_plan_cover_device``).  Inside
        a recorded CUDA graph the library could not measure a batch itself."""
        torch, lib = self._torch, self._lib
        spec = self._spec
        nd = self._ndim
        lo = numpy.array([spec.lo[i] for i in range(nd)])
        hi = numpy.array([spec.hi[i] for i in range(nd)])
        for i in range(nd):                 # unbounded normals: five standard deviations
            if spec.type[i] == _lib.PRIOR_GAUSSIAN:
                lo[i] = max(lo[i], spec.mu[i] - 5.0 * spec.sigma[i])
                hi[i] = min(hi[i], spec.mu[i] + 5.0 * spec.sigma[i])
        rs = numpy.random.RandomState(12345)
        q = torch.tensor(lo + (hi - lo) * rs.uniform(size=(n, nd)), dtype=torch.float64, device=self._dev)
        logp = torch.empty(n, dtype=torch.float64, device=self._dev)
        skip = torch.empty(n, dtype=torch.uint8, device=self._dev)
        params = torch.empty((_lib.GWIN_NPARAM, n), dtype=torch.float64, device=self._dev)
        st = torch.cuda.current_stream(self._dev).cuda_stream
        _lib.check(lib.gwin_ens_prior(C.byref(spec), n, q.data_ptr(), logp.data_ptr(),
                                      skip.data_ptr(), params.data_ptr(), st))
        self._plan.cover(params, param_major=True)

    def run(self, niterations, **kwargs):
        """``niterations`` steps.  All arguments of one iteration are constant -- the
        iteration number lives in a device counter -- so the iteration is recorded once as a
        CUDA graph and replayed (``use_graph=False`` or ``GWIN_SAMPLER_GRAPH=0`` launches
        every kernel from the host instead; the chains are identical)."""
        torch, lib = self._torch, self._lib
        if self._p is None:
            self._initialise(self.p0)
        nt, nw, nd = self._ntemps, self._nwalkers, self._ndim
        W = nt * nw
        f64 = dict(dtype=torch.float64, device=self._dev)
        chain = torch.empty((nt, nw, niterations, nd), **f64)
        lnp_h = torch.empty((nt, nw, niterations), **f64)
        lnl_h = torch.empty((nt, nw, niterations), **f64)
        q_loglr = torch.empty(self._Wh, **f64)
        g0 = self._iterations + self._iter0          # global iteration of this call's first step
        perms = None
        if nt > 1:
            # pairings of the temperature swaps: per global iteration and level two random permutations of
            # the walkers, keyed by (seed, global iteration) -- identical on every rank and independent of how
            # a run is split into run() calls.  Drawn on the device (argsort of uniform keys from torch's
            # Philox generator): on the host this cost 0.5 ms per iteration, half of a C4 iteration.
            gen = torch.Generator(device=self._dev)
            keys = torch.empty((niterations, nt - 1, 2, nw), dtype=torch.float64, device=self._dev)
            for it in range(niterations):
                gen.manual_seed((self._seed * 1000003 + g0 + it) % (2 ** 62))
                torch.rand((nt - 1, 2, nw), generator=gen, dtype=torch.float64, device=self._dev, out=keys[it])
            perms = keys.argsort(dim=-1).to(torch.int32).contiguous()
            del keys
        ctr = torch.zeros(1, dtype=torch.int32, device=self._dev)   # iteration within this call

        def pre(half):
            st = torch.cuda.current_stream(self._dev).cuda_stream
            _lib.check(lib.gwin_ens_propose(
                self._seed, g0, ctr.data_ptr(), half, nt, nw, nd, self._a, self._move,
                self._interleaved, self._p.data_ptr(), self._q.data_ptr(),
                self._logfac.data_ptr(), st))
            self._eval_local(self._Wh, self._q, self._q_logp, self._skip, self._params, q_loglr)

        def post(half):
            st = torch.cuda.current_stream(self._dev).cuda_stream
            if self._world > 1:
                self._scatter_back(self._Wh, q_loglr)
            _lib.check(lib.gwin_ens_accept(
                self._seed, g0, ctr.data_ptr(), half, nt, nw, nd, self._interleaved,
                self._betas.data_ptr(), self._q.data_ptr(), self._logfac.data_ptr(),
                q_loglr.data_ptr(), self._q_logp.data_ptr(), self._p.data_ptr(),
                self._logl.data_ptr(), self._lnprob.data_ptr(), self._naccept.data_ptr(), st))
            if half == 0:
                return
            if nt > 1:
                _lib.check(lib.gwin_ens_swap_all(
                    self._seed, g0, ctr.data_ptr(), nt, nw, nd, self._betas.data_ptr(),
                    perms.data_ptr(), (nt - 1) * 2 * nw, self._p.data_ptr(),
                    self._logl.data_ptr(), self._lnprob.data_ptr(), self._nswap_acc.data_ptr(), st))
            _lib.check(lib.gwin_ens_record(
                W, nd, niterations, 0, ctr.data_ptr(), self._p.data_ptr(), self._lnprob.data_ptr(),
                self._logl.data_ptr(), chain.data_ptr(), lnp_h.data_ptr(), lnl_h.data_ptr(), st))
            _lib.check(lib.gwin_ens_tick(ctr.data_ptr(), st))

        def iteration():
            for half in (0, 1):
                pre(half)
                if self._world > 1:
                    self._gather(self._Wh)
                post(half)

        # recording costs ~50 ms: by default only runs long enough to repay it are recorded
        if "use_graph" in kwargs:
            use_graph = bool(kwargs["use_graph"]) and niterations >= 4
        else:
            use_graph = self.use_graph and niterations >= 64
        if use_graph:
            # the first iteration runs eagerly on the capture stream (workspaces, the NCCL
            # communicator and the block schedule come into being here), the second is
            # recorded, the rest are replays of the record
            cur = torch.cuda.current_stream(self._dev)
            side = torch.cuda.Stream(self._dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                iteration()
            cur.wait_stream(side)
            torch.cuda.synchronize(self._dev)
        # several GPUs: GWIN_SAMPLER_GRAPH_NCCL=1 records the all-gathers too (one graph for everything);
        # otherwise they stay outside and everything between two of them is one graph
        nccl_in_graph = self._world > 1 and os.environ.get("GWIN_SAMPLER_GRAPH_NCCL", "0") == "1"
        if use_graph and (self._world == 1 or nccl_in_graph):
            # a replay costs the same whatever the graph holds (~0.25 ms between two replays on this stack):
            # several iterations are recorded per graph; what does not fill a graph runs kernel by kernel
            per_graph = max(1, min(int(os.environ.get("GWIN_SAMPLER_GRAPH_ITERS", "16")), niterations - 1))
            graph = torch.cuda.CUDAGraph()
            evals0 = self.likelihood_evals
            with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
                for _ in range(per_graph):
                    iteration()
            self.likelihood_evals = evals0 + 2 * self._Wh          # (the accounting below expects one recorded iteration)
            n_replay, n_rest = divmod(niterations - 1, per_graph)
            for _ in range(n_replay):
                graph.replay()
            for _ in range(n_rest):
                iteration()
            self.likelihood_evals -= 2 * self._Wh * n_rest          # (counted again below)
            self._graph_keepalive = (graph, chain, lnp_h, lnl_h, perms, ctr, q_loglr)
        elif use_graph:
            graphs = []
            for half in (0, 1):
                for part in (pre, post):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        part(half)
                    graphs.append(g)
            for _ in range(niterations - 1):
                for half in (0, 1):
                    graphs[2 * half].replay()
                    self._gather(self._Wh)
                    graphs[2 * half + 1].replay()
            self._graph_keepalive = (graphs, chain, lnp_h, lnl_h, perms, ctr, q_loglr)
        else:
            for _ in range(niterations):
                iteration()
        if use_graph:       # the record pass counted one replay; the others ran without Python
            self.likelihood_evals += 2 * self._Wh * (niterations - 2)
        if nt > 1:
            self._nswap += nw * niterations
        self._iterations += niterations
        # histories stay on the device until somebody reads them (chain / lnpost / model_stats)
        self._pending.append((chain, lnp_h, lnl_h))
        self._pos = None
        return self._p, self._lnprob, (self._seed, self._iterations)

    def _flush(self):
        """Bring pending device histories to the host (one D2H per history per run())."""
        for chain, lnp_h, lnl_h in self._pending:
            c, a, b = chain.cpu().numpy(), lnp_h.cpu().numpy(), lnl_h.cpu().numpy()
            if self._chain is None:
                self._chain, self._lnprob_h, self._logl_h = c, a, b
            else:
                self._chain = numpy.concatenate((self._chain, c), axis=2)
                self._lnprob_h = numpy.concatenate((self._lnprob_h, a), axis=2)
                self._logl_h = numpy.concatenate((self._logl_h, b), axis=2)
        self._pending = []

    @property
    def pos(self):
        """Current positions (host copy)."""
        return None if self._p is None else self._p.cpu().numpy()

    def write_state(self, fp):
        """The RNG is counter based: seed + iteration counter are the whole state."""
        fp.attrs['philox_seed'] = int(self._seed)
        fp.attrs['philox_iteration'] = int(self._iterations + self._iter0)

    def set_state_from_file(self, fp):
        self._seed = int(fp.attrs['philox_seed'])
        self._iter0 = int(fp.attrs['philox_iteration']) - self._iterations
        self.lastclear = int(fp.attrs.get('niterations', 0))

    def clear_chain(self):
        self.lastclear = self.niterations
        self._iter0 += self._iterations
        self._iterations = 0
        self._chain = self._lnprob_h = self._logl_h = None
        self._pending = []
        self._naccept.zero_()
        self._nswap_acc.zero_()
        self._nswap = 0


class CudaPTSampler(_DeviceEnsemble):
    """Device-resident counterpart of ``EmceePTSampler`` (``gwin/sampler/emcee.py:253-432``).
    The tempered quantity is ``loglr`` (``lognl`` is a constant that cancels in every
    acceptance ratio)."""
    name = "emcee_pt_cuda"

    def __init__(self, model, ntemps, nwalkers, pool=None, model_call=None, betas=None,
                 seed=None, process_group=None):
        ndim = len(model.sampling_params)
        if betas is None:
            betas = default_beta_ladder(ndim, ntemps)
        super(CudaPTSampler, self).__init__(model, ntemps, nwalkers, betas, move=1, interleaved=1,
                                            seed=seed, process_group=process_group)

    @classmethod
    def from_cli(cls, opts, model, pool=None, model_call=None):
        return cls(model, opts.ntemps, opts.nwalkers)

    @property
    def chain(self):
        """ntemps x nwalkers x niterations x ndim."""
        self._flush()
        return self._chain if self._chain is not None else \
            numpy.empty((self._ntemps, self._nwalkers, 0, self._ndim))

    @property
    def lnpost(self):
        self._flush()
        return self._lnprob_h

    @property
    def model_stats(self):
        self._flush()
        logl = self._logl_h
        logp = self._lnprob_h - logl * self._betas_host.reshape(-1, 1, 1)
        return FieldArray.from_kwargs(loglr=logl, logprior=logp)


class CudaEnsembleSampler(_DeviceEnsemble):
    """Device-resident counterpart of ``EmceeEnsembleSampler`` (``gwin/sampler/emcee.py:46-194``):
    stretch move on two contiguous half-ensembles, target ``logprior + loglr`` (= ``logplr``)."""
    name = "emcee_cuda"
    _keep_temp_axis = False

    def __init__(self, model, nwalkers, pool=None, model_call=None, seed=None, process_group=None):
        super(CudaEnsembleSampler, self).__init__(model, 1, nwalkers, [1.0], move=0, interleaved=0,
                                                  seed=seed, process_group=process_group)

    @classmethod
    def from_cli(cls, opts, model, pool=None, model_call=None):
        return cls(model, opts.nwalkers)

    @property
    def chain(self):
        """nwalkers x niterations x ndim."""
        self._flush()
        return self._chain[0] if self._chain is not None else \
            numpy.empty((self._nwalkers, 0, self._ndim))

    @property
    def lnpost(self):
        self._flush()
        return self._lnprob_h[0]

    @property
    def acceptance_fraction(self):
        return super(CudaEnsembleSampler, self).acceptance_fraction[0]

    @property
    def model_stats(self):
        self._flush()
        logl = self._logl_h[0]
        return FieldArray.from_kwargs(loglr=logl, logprior=self._lnprob_h[0] - logl)
