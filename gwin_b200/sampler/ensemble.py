"""Native ensemble steppers with emcee-2.x semantics.

The reference delegates the ensemble step to ``emcee.EnsembleSampler`` /
``emcee.PTSampler`` (``gwin/sampler/emcee.py:74-76, 286-289``), which evaluate
``k/2`` (or ``ntemps * k/2``) proposals per half-step through ``pool.map``.  These
classes reproduce that step -- stretch move with ``a = 2``, two half-ensembles,
blobs carried with accepted positions; for PT the interleaved halves, the
``beta * logl + logp`` acceptance and the adjacent-temperature swaps -- but hand
each half-step's proposals to ONE batched likelihood call.

``lnpostfn(points[W, ndim]) -> (lnprob[W], blobs)`` where ``blobs`` is a list of W
stat tuples or None; the PT variant takes ``logl_logp(points) -> (logl, logp)``.

Attribution: the step order, the use of the random stream and the swap rule follow
emcee 2.x (``emcee/ensemble.py``, ``emcee/ptsampler.py``; Copyright 2010-2013 Daniel
Foreman-Mackey & contributors, MIT licence) so that chains drawn here are
RNG-stream-identical to chains drawn by the reference through emcee; emcee is a
third-party dependency of the reference (``setup.py``), not part of /root/reference.
"""
import numpy


class EnsembleStepper(object):
    """Affine-invariant stretch-move ensemble sampler (Goodman & Weare)."""

    def __init__(self, nwalkers, dim, lnpostfn, a=2.0, random_state=None):
        if nwalkers % 2 != 0:
            raise ValueError("The number of walkers must be even.")
        if nwalkers < 2 * dim:
            raise ValueError("The number of walkers needs to be more than twice "
                             "the dimension of your parameter space.")
        self.k = nwalkers
        self.dim = dim
        self.a = a
        self.lnpostfn = lnpostfn
        self._random = numpy.random.mtrand.RandomState()
        if random_state is not None:
            self.random_state = random_state
        self.reset()

    # -- RNG state, as emcee exposes it -------------------------------------
    @property
    def random_state(self):
        return self._random.get_state()

    @random_state.setter
    def random_state(self, state):
        self._random.set_state(state)

    def reset(self):
        self.iterations = 0
        self.naccepted = numpy.zeros(self.k)
        self._chain = numpy.empty((self.k, 0, self.dim))
        self._lnprob = numpy.empty((self.k, 0))
        self._blobs = []
        self._last = None

    def clear_blobs(self):
        self._blobs = []

    @property
    def chain(self):
        return self._chain

    @property
    def lnprobability(self):
        return self._lnprob

    @property
    def blobs(self):
        """[niterations][nwalkers] stat tuples."""
        return self._blobs

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(self.iterations, 1)

    def _eval(self, pts):
        lnprob, blobs = self.lnpostfn(pts)
        lnprob = numpy.asarray(lnprob, dtype=float)
        if numpy.any(numpy.isnan(lnprob)):
            raise ValueError("lnprob returned NaN.")       # as emcee does
        return lnprob, blobs

    def run_mcmc(self, p0, niterations, lnprob0=None, blobs0=None):
        p = numpy.array(p0, dtype=float)
        if p.shape != (self.k, self.dim):
            raise ValueError("p0 must have shape (nwalkers, dim)")
        lnprob, blobs = lnprob0, blobs0
        if lnprob is None:
            lnprob, blobs = self._eval(p)
        if numpy.any(numpy.isinf(lnprob) & (lnprob > 0)):
            raise ValueError("The initial lnprob was +inf.")
        lnprob = numpy.array(lnprob, dtype=float)
        blobs = list(blobs) if blobs is not None else None
        halfk = self.k // 2
        first = numpy.arange(self.k) < halfk
        chain = numpy.empty((self.k, niterations, self.dim))
        lnp_hist = numpy.empty((self.k, niterations))
        for it in range(niterations):
            for S0, S1 in ((first, ~first), (~first, first)):
                s, c = p[S0], p[S1]
                Ns, Nc = len(s), len(c)
                zz = ((self.a - 1.) * self._random.rand(Ns) + 1) ** 2. / self.a
                rint = self._random.randint(Nc, size=(Ns,))
                q = c[rint] - zz[:, numpy.newaxis] * (c[rint] - s)
                newlnprob, newblobs = self._eval(q)
                lnpdiff = (self.dim - 1.) * numpy.log(zz) + newlnprob - lnprob[S0]
                accept = lnpdiff > numpy.log(self._random.rand(len(lnpdiff)))
                idx = numpy.flatnonzero(S0)[accept]
                p[idx] = q[accept]
                lnprob[idx] = newlnprob[accept]
                self.naccepted[idx] += 1
                if blobs is not None and newblobs is not None:
                    for i, j in zip(idx, numpy.flatnonzero(accept)):
                        blobs[i] = newblobs[j]
            chain[:, it, :] = p
            lnp_hist[:, it] = lnprob
            if blobs is not None:
                self._blobs.append(list(blobs))
            self.iterations += 1
        self._chain = numpy.concatenate((self._chain, chain), axis=1)
        self._lnprob = numpy.concatenate((self._lnprob, lnp_hist), axis=1)
        self._last = (p, lnprob, blobs)
        if blobs is not None:
            return p, lnprob, self.random_state, blobs
        return p, lnprob, self.random_state


_TSTEP = numpy.array([25.2741, 7., 4.47502, 3.5236, 3.0232, 2.71225, 2.49879,
                      2.34226, 2.22198, 2.12628])


def default_beta_ladder(ndim, ntemps, Tmax=None):
    """Geometric ladder ``beta_i = tstep^-i`` with the dimension-dependent
    spacing emcee uses for ndim <= 10 (constant ~25 % swap acceptance for a
    Gaussian); beyond that the spacing falls off as ``1 + 2.66/sqrt(ndim)``."""
    tstep = _TSTEP[ndim - 1] if ndim <= len(_TSTEP) else 1.0 + 2.66 / numpy.sqrt(ndim)
    if Tmax is not None:
        ntemps = int(numpy.log(Tmax) / numpy.log(tstep) + 2)
    return numpy.exp(numpy.linspace(0, -(ntemps - 1) * numpy.log(tstep), ntemps))


class PTStepper(object):
    """Parallel-tempered ensemble sampler with emcee ``PTSampler`` semantics."""

    def __init__(self, ntemps, nwalkers, dim, logl_logp, betas=None, a=2.0,
                 random_state=None):
        if nwalkers % 2 != 0:
            raise ValueError("The number of walkers must be even.")
        if nwalkers < 2 * dim:
            raise ValueError("The number of walkers must be greater than 2*dimension.")
        self.ntemps, self.nwalkers, self.dim, self.a = ntemps, nwalkers, dim, a
        self.logl_logp = logl_logp
        self.betas = numpy.asarray(default_beta_ladder(dim, ntemps) if betas is None
                                   else betas, dtype=float).copy()
        self._random = numpy.random.mtrand.RandomState()
        if random_state is not None:
            self.random_state = random_state
        self.reset()

    @property
    def random_state(self):
        return self._random.get_state()

    @random_state.setter
    def random_state(self, state):
        self._random.set_state(state)

    def reset(self):
        self.iterations = 0
        self.nswap = numpy.zeros(self.ntemps)
        self.nswap_accepted = numpy.zeros(self.ntemps)
        self.nprop = numpy.zeros((self.ntemps, self.nwalkers))
        self.nprop_accepted = numpy.zeros((self.ntemps, self.nwalkers))
        self._chain = numpy.empty((self.ntemps, self.nwalkers, 0, self.dim))
        self._lnprob = numpy.empty((self.ntemps, self.nwalkers, 0))
        self._lnlikelihood = numpy.empty((self.ntemps, self.nwalkers, 0))

    chain = property(lambda self: self._chain)
    lnprobability = property(lambda self: self._lnprob)
    lnlikelihood = property(lambda self: self._lnlikelihood)

    @property
    def acceptance_fraction(self):
        return self.nprop_accepted / numpy.maximum(self.nprop, 1)

    @property
    def tswap_acceptance_fraction(self):
        return self.nswap_accepted / numpy.maximum(self.nswap, 1)

    def _eval(self, pts):
        logl, logp = self.logl_logp(pts.reshape(-1, self.dim))
        logl = numpy.asarray(logl, dtype=float).reshape(pts.shape[:-1])
        logp = numpy.asarray(logp, dtype=float).reshape(pts.shape[:-1])
        # emcee: points with logp = -inf are never evaluated and get logl = 0
        logl = numpy.where(numpy.isneginf(logp), 0., logl)
        if numpy.any(numpy.isnan(logl)) or numpy.any(numpy.isnan(logp)):
            raise ValueError("log-likelihood / log-prior returned NaN.")
        return logl, logp

    def run_mcmc(self, p0, niterations, lnprob0=None, lnlike0=None):
        p = numpy.array(p0, dtype=float)
        if p.shape != (self.ntemps, self.nwalkers, self.dim):
            raise ValueError("p0 must have shape (ntemps, nwalkers, dim)")
        if lnprob0 is None or lnlike0 is None:
            logl, logp = self._eval(p)
            lnprob = logl * self.betas.reshape(-1, 1) + logp
        else:
            lnprob, logl = numpy.array(lnprob0, float), numpy.array(lnlike0, float)
        nr = self._random
        half = self.nwalkers // 2
        chain = numpy.empty((self.ntemps, self.nwalkers, niterations, self.dim))
        lnp_h = numpy.empty((self.ntemps, self.nwalkers, niterations))
        lnl_h = numpy.empty((self.ntemps, self.nwalkers, niterations))
        for it in range(niterations):
            for j in (0, 1):
                jupdate, jsample = j, (j + 1) % 2
                pupdate = p[:, jupdate::2, :]
                psample = p[:, jsample::2, :]
                zs = numpy.exp(nr.uniform(low=-numpy.log(self.a), high=numpy.log(self.a),
                                          size=(self.ntemps, half)))
                qs = numpy.zeros((self.ntemps, half, self.dim))
                for k in range(self.ntemps):
                    js = nr.randint(0, high=half, size=half)
                    qs[k] = psample[k, js, :] + zs[k, :].reshape((half, 1)) * \
                        (pupdate[k] - psample[k, js, :])
                qslogl, qslogp = self._eval(qs)
                qslnprob = qslogl * self.betas.reshape(-1, 1) + qslogp
                logpaccept = self.dim * numpy.log(zs) + qslnprob - lnprob[:, jupdate::2]
                logr = numpy.log(nr.uniform(low=0.0, high=1.0, size=(self.ntemps, half)))
                accepts = logr < logpaccept
                pupdate[accepts] = qs[accepts]
                lnprob[:, jupdate::2][accepts] = qslnprob[accepts]
                logl[:, jupdate::2][accepts] = qslogl[accepts]
                self.nprop[:, jupdate::2] += 1.0
                self.nprop_accepted[:, jupdate::2] += accepts
            p, lnprob, logl = self._temperature_swaps(p, lnprob, logl)
            chain[:, :, it, :] = p
            lnp_h[:, :, it] = lnprob
            lnl_h[:, :, it] = logl
            self.iterations += 1
        self._chain = numpy.concatenate((self._chain, chain), axis=2)
        self._lnprob = numpy.concatenate((self._lnprob, lnp_h), axis=2)
        self._lnlikelihood = numpy.concatenate((self._lnlikelihood, lnl_h), axis=2)
        return p, lnprob, logl

    def _temperature_swaps(self, p, lnprob, logl):
        nr = self._random
        for i in range(self.ntemps - 1, 0, -1):
            bi, bi1 = self.betas[i], self.betas[i - 1]
            dbeta = bi1 - bi
            iperm = nr.permutation(self.nwalkers)
            i1perm = nr.permutation(self.nwalkers)
            raccept = numpy.log(nr.uniform(size=self.nwalkers))
            paccept = dbeta * (logl[i, iperm] - logl[i - 1, i1perm])
            self.nswap[i] += self.nwalkers
            self.nswap[i - 1] += self.nwalkers
            asel = paccept > raccept
            nacc = numpy.sum(asel)
            self.nswap_accepted[i] += nacc
            self.nswap_accepted[i - 1] += nacc
            ptemp = numpy.copy(p[i, iperm[asel], :])
            ltemp = numpy.copy(logl[i, iperm[asel]])
            prtemp = numpy.copy(lnprob[i, iperm[asel]])
            p[i, iperm[asel], :] = p[i - 1, i1perm[asel], :]
            logl[i, iperm[asel]] = logl[i - 1, i1perm[asel]]
            lnprob[i, iperm[asel]] = lnprob[i - 1, i1perm[asel]] - \
                dbeta * logl[i - 1, i1perm[asel]]
            p[i - 1, i1perm[asel], :] = ptemp
            logl[i - 1, i1perm[asel]] = ltemp
            lnprob[i - 1, i1perm[asel]] = prtemp + dbeta * ltemp
        return p, lnprob, logl

    def thermodynamic_integration_log_evidence(self, fburnin=0.1):
        """log Z by thermodynamic integration over beta (emcee's estimator; the
        reference calls it at ``gwin/sampler/emcee.py:902-954``).  Returns
        (lnZ, dlnZ)."""
        istart = int(self._lnlikelihood.shape[2] * fburnin + 0.5)
        mean_logls = numpy.mean(numpy.mean(self._lnlikelihood, axis=1)[:, istart:], axis=1)
        betas = numpy.asarray(self.betas)
        if betas[-1] != 0:
            betas = numpy.concatenate((betas, [0]))
            mean_logls = numpy.concatenate((mean_logls, [mean_logls[-1]]))
        order = numpy.argsort(betas)[::-1]
        betas, mean_logls = betas[order], mean_logls[order]
        lnZ = -numpy.trapz(mean_logls, betas)
        betas2 = numpy.concatenate((betas[::2], [0])) if betas[::2][-1] != 0 else betas[::2]
        logls2 = numpy.concatenate((mean_logls[::2], [mean_logls[-1]])) \
            if len(betas2) != len(mean_logls[::2]) else mean_logls[::2]
        lnZ2 = -numpy.trapz(logls2, betas2)
        return lnZ, numpy.abs(lnZ - lnZ2)
