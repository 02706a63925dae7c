"""Vectorised priors for the batched model call.

The reference evaluates ``self.prior_distribution(**params)`` one walker at a
time through ``pycbc.distributions.JointDistribution`` (``gwin/models/base.py:
543-554``) and applies its boundary conditions in ``_transform_params``
(``base.py:602-621``).  The classes below keep that protocol -- ``params``,
``bounds``, ``__call__ == logpdf``, ``pdf``, ``rvs``,
``apply_boundary_conditions``, ``variable_args`` -- but accept arrays of W
walkers so the step before the kernel costs one numpy pass instead of W Python
calls.  Covered: the distributions the reference's docs and sampler tests use
(``docs/command-line/gwin.rst:243-312``, ``test/test_sampler.py:74-98``).
"""
from collections import OrderedDict, namedtuple

import numpy

Bounds = namedtuple("Bounds", ["min", "max"])


class FieldArray(numpy.recarray):
    """numpy record array with the two conveniences the reference relies on
    (``fieldnames``, ``from_arrays``; ``gwin/models/base.py:597``)."""

    @property
    def fieldnames(self):
        return self.dtype.names

    @classmethod
    def from_arrays(cls, arrays, names):
        arrays = [numpy.asarray(a) for a in arrays]
        dtype = [(str(n), a.dtype) for n, a in zip(names, arrays)]
        out = numpy.empty(arrays[0].shape, dtype=dtype).view(cls)
        for n, a in zip(names, arrays):
            out[n] = a
        return out

    @classmethod
    def from_kwargs(cls, **kwargs):
        return cls.from_arrays(list(kwargs.values()), list(kwargs.keys()))


class _BoundedDist(object):
    name = None

    def __init__(self, **params):
        self._bounds = OrderedDict()
        for p, b in params.items():
            if b is None:
                b = self._default_bounds
            self._bounds[p] = Bounds(float(b[0]), float(b[1]))
        self._params = list(self._bounds.keys())

    @property
    def params(self):
        return self._params

    @property
    def bounds(self):
        return self._bounds

    def _inside(self, kwargs):
        ok = True
        for p, b in self._bounds.items():
            x = numpy.asarray(kwargs[p], dtype=float)
            ok = ok & (x >= b.min) & (x < b.max)     # [UPSTREAM] pycbc Bounds: closed/open
        return ok

    def apply_boundary_conditions(self, **kwargs):
        return kwargs

    def logpdf(self, **kwargs):
        with numpy.errstate(divide="ignore"):
            return numpy.log(self.pdf(**kwargs))

    def __call__(self, **kwargs):
        return self.logpdf(**kwargs)

    def rvs(self, size=1, rng=None):
        rng = numpy.random if rng is None else rng
        arrs = [self._draw(p, b, size, rng) for p, b in self._bounds.items()]
        return FieldArray.from_arrays(arrs, self._params)


class Uniform(_BoundedDist):
    """``pycbc.distributions.Uniform(param=(min, max), ...)``."""
    name = "uniform"

    def pdf(self, **kwargs):
        norm = 1.0
        for b in self._bounds.values():
            norm *= (b.max - b.min)
        return numpy.where(self._inside(kwargs), 1.0 / norm, 0.0)

    def _draw(self, p, b, size, rng):
        return rng.uniform(b.min, b.max, size=size)


class UniformAngle(Uniform):
    """Uniform on a cyclic angle, default ``[0, 2 pi)``; boundary conditions wrap
    values into the domain (``pycbc.distributions.UniformAngle``)."""
    name = "uniform_angle"
    _default_bounds = (0.0, 2.0 * numpy.pi)

    def apply_boundary_conditions(self, **kwargs):
        out = dict(kwargs)
        for p, b in self._bounds.items():
            if p in out:
                x = numpy.asarray(out[p], dtype=float)
                w = b.max - b.min
                y = (x - b.min) % w
                # (a value a rounding error below the lower bound wraps to the upper bound itself, which is
                #  outside the half-open domain: it is the lower bound)
                out[p] = numpy.where(y >= w, 0.0, y) + b.min
        return out


class SinAngle(UniformAngle):
    """p(theta) ~ sin(theta) on ``[0, pi]`` (polar angle, e.g. inclination);
    reflected boundaries (``pycbc.distributions.SinAngle``)."""
    name = "sin_angle"
    _default_bounds = (0.0, numpy.pi)
    _func = staticmethod(numpy.cos)
    _dfunc = staticmethod(numpy.sin)
    _arcfunc = staticmethod(numpy.arccos)

    def apply_boundary_conditions(self, **kwargs):
        out = dict(kwargs)
        lo, hi = self._default_bounds
        for p in self._bounds:
            if p in out:
                x = numpy.asarray(out[p], dtype=float) - lo
                period = 2.0 * (hi - lo)
                x = x % period
                x = numpy.where(x > (hi - lo), period - x, x)
                out[p] = x + lo
        return out

    def pdf(self, **kwargs):
        val = 1.0
        for p, b in self._bounds.items():
            x = numpy.asarray(kwargs[p], dtype=float)
            norm = abs(self._func(b.min) - self._func(b.max))
            val = val * numpy.abs(self._dfunc(x)) / norm
        return numpy.where(self._inside(kwargs), val, 0.0)

    def _draw(self, p, b, size, rng):
        a, c = self._func(b.min), self._func(b.max)
        return self._arcfunc(rng.uniform(min(a, c), max(a, c), size=size))


class CosAngle(SinAngle):
    """p(theta) ~ cos(theta) on ``[-pi/2, pi/2]`` (e.g. declination)."""
    name = "cos_angle"
    _default_bounds = (-numpy.pi / 2.0, numpy.pi / 2.0)
    _func = staticmethod(numpy.sin)
    _dfunc = staticmethod(numpy.cos)
    _arcfunc = staticmethod(numpy.arcsin)


class UniformRadius(_BoundedDist):
    """Uniform in volume: p(r) ~ r^2 (``pycbc.distributions.UniformRadius``)."""
    name = "uniform_radius"

    def pdf(self, **kwargs):
        val = 1.0
        for p, b in self._bounds.items():
            x = numpy.asarray(kwargs[p], dtype=float)
            val = val * 3.0 * x * x / (b.max ** 3 - b.min ** 3)
        return numpy.where(self._inside(kwargs), val, 0.0)

    def _draw(self, p, b, size, rng):
        u = rng.uniform(0.0, 1.0, size=size)
        return (b.min ** 3 + u * (b.max ** 3 - b.min ** 3)) ** (1.0 / 3.0)


class Gaussian(_BoundedDist):
    """Independent (optionally truncated) normals:
    ``Gaussian(x=(min, max), x_mean=.., x_var=..)``."""
    name = "gaussian"

    def __init__(self, **params):
        means = {k[:-5]: params.pop(k) for k in list(params) if k.endswith("_mean")}
        vars_ = {k[:-4]: params.pop(k) for k in list(params) if k.endswith("_var")}
        for p in params:
            if params[p] is None:
                params[p] = (-numpy.inf, numpy.inf)
        super(Gaussian, self).__init__(**params)
        self._mean = {p: float(means.get(p, 0.0)) for p in self._params}
        self._var = {p: float(vars_.get(p, 1.0)) for p in self._params}

    def _norm(self, p, b):
        """Probability mass of the untruncated normal inside the bounds; formed from complementary error
        functions on the side of the mean the interval lies on, so that an interval far out in a tail keeps
        its relative accuracy (erf differences there cancel)."""
        from scipy.special import erf, erfc
        s = numpy.sqrt(2.0 * self._var[p])
        lo, hi = (b.min - self._mean[p]) / s, (b.max - self._mean[p]) / s
        if lo > 0.0:
            return 0.5 * (erfc(lo) - erfc(hi))
        if hi < 0.0:
            return 0.5 * (erfc(-hi) - erfc(-lo))
        return 0.5 * (erf(hi) - erf(lo))

    def pdf(self, **kwargs):
        val = 1.0
        for p, b in self._bounds.items():
            x = numpy.asarray(kwargs[p], dtype=float)
            v = self._var[p]
            val = val * numpy.exp(-0.5 * (x - self._mean[p]) ** 2 / v) / \
                numpy.sqrt(2.0 * numpy.pi * v) / self._norm(p, b)
        return numpy.where(self._inside(kwargs), val, 0.0)

    def _draw(self, p, b, size, rng):
        out = numpy.empty(size)
        n = 0
        while n < size:
            x = rng.normal(self._mean[p], numpy.sqrt(self._var[p]), size=size)
            x = x[(x >= b.min) & (x <= b.max)][:size - n]
            out[n:n + len(x)] = x
            n += len(x)
        return out


class UniformSky(object):
    """Isotropic sky position: uniform ``ra`` on [0, 2pi) x cos-weighted ``dec``
    (``pycbc.distributions.UniformSky``)."""
    name = "uniform_sky"

    def __init__(self, polar_angle="dec", azimuthal_angle="ra"):
        self._az = UniformAngle(**{azimuthal_angle: None})
        self._po = CosAngle(**{polar_angle: None})
        self._params = [azimuthal_angle, polar_angle]
        self._bounds = OrderedDict(list(self._az.bounds.items()) +
                                   list(self._po.bounds.items()))

    params = property(lambda self: self._params)
    bounds = property(lambda self: self._bounds)

    def apply_boundary_conditions(self, **kwargs):
        return self._po.apply_boundary_conditions(
            **self._az.apply_boundary_conditions(**kwargs))

    def pdf(self, **kwargs):
        return self._az.pdf(**kwargs) * self._po.pdf(**kwargs)

    def logpdf(self, **kwargs):
        with numpy.errstate(divide="ignore"):
            return numpy.log(self.pdf(**kwargs))

    __call__ = logpdf

    def rvs(self, size=1, rng=None):
        a = self._az.rvs(size, rng)
        b = self._po.rvs(size, rng)
        return FieldArray.from_arrays([a[self._params[0]], b[self._params[1]]],
                                      self._params)


class JointDistribution(object):
    """``pycbc.distributions.JointDistribution(variable_args, *dists,
    constraints=...)``: callable log-prior over named parameters.

    ``constraints`` is an optional list of callables ``f(params_dict) -> bool
    array``; points failing any of them get ``-inf``.
    """
    name = "joint"

    def __init__(self, variable_args, *distributions, **kwargs):
        self.variable_args = tuple(variable_args)
        self.distributions = distributions
        self._constraints = kwargs.pop("constraints", None) or []
        if kwargs:
            raise TypeError("unexpected keyword(s) %s" % list(kwargs))
        covered = [p for d in distributions for p in d.params]
        if sorted(covered) != sorted(self.variable_args):
            raise ValueError("distributions cover %s but variable_args are %s"
                             % (sorted(covered), sorted(self.variable_args)))

    def apply_boundary_conditions(self, **params):
        for d in self.distributions:
            params.update(d.apply_boundary_conditions(
                **{p: params[p] for p in d.params if p in params}))
        return params

    def __call__(self, **params):
        logp = 0.0
        for d in self.distributions:
            logp = logp + d(**{p: params[p] for p in d.params})
        for c in self._constraints:
            logp = numpy.where(c(params), logp, -numpy.inf)
        return logp

    def rvs(self, size=1, rng=None):
        draws = {}
        for d in self.distributions:
            r = d.rvs(size, rng)
            for p in d.params:
                draws[p] = r[p]
        out = FieldArray.from_arrays([draws[p] for p in self.variable_args],
                                     self.variable_args)
        if self._constraints:
            # redraw rejected points
            for _ in range(1000):
                ok = numpy.ones(size, dtype=bool)
                for c in self._constraints:
                    ok &= numpy.asarray(c({p: out[p] for p in self.variable_args}))
                bad = numpy.flatnonzero(~ok)
                if len(bad) == 0:
                    break
                for d in self.distributions:
                    r = d.rvs(len(bad), rng)
                    for p in d.params:
                        out[p][bad] = r[p]
        return out


distribs = {c.name: c for c in (Uniform, UniformAngle, SinAngle, CosAngle,
                                UniformRadius, Gaussian, UniformSky)}
