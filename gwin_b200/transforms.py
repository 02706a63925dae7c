"""Vectorised parameter transforms (the subset of ``pycbc.transforms`` the
reference's sampler tests and docs exercise; ``gwin/models/base.py:126-184``,
``base_data.py:153-161``).  Each transform maps a dict of arrays to a dict of
arrays and knows the Jacobian of its forward and inverse maps."""
import numpy


class BaseTransform(object):
    name = None
    inputs = ()
    outputs = ()

    def transform(self, maps):
        raise NotImplementedError

    def inverse_transform(self, maps):
        raise NotImplementedError

    def jacobian(self, maps):
        raise NotImplementedError

    def inverse_jacobian(self, maps):
        raise NotImplementedError

    @property
    def inverse(self):
        return _Inverse(self)


class _Inverse(BaseTransform):
    def __init__(self, t):
        self._t = t
        self.name = "inverse_" + str(t.name)
        self.inputs, self.outputs = t.outputs, t.inputs

    def transform(self, maps):
        return self._t.inverse_transform(maps)

    def inverse_transform(self, maps):
        return self._t.transform(maps)

    def jacobian(self, maps):
        return self._t.inverse_jacobian(maps)

    def inverse_jacobian(self, maps):
        return self._t.jacobian(maps)


class MchirpQToMass1Mass2(BaseTransform):
    """(mchirp, q = m1/m2 >= 1) -> (mass1, mass2)."""
    name = "mchirp_q_to_mass1_mass2"
    inputs = ("mchirp", "q")
    outputs = ("mass1", "mass2")

    def transform(self, maps):
        mc = numpy.asarray(maps["mchirp"], dtype=float)
        q = numpy.asarray(maps["q"], dtype=float)
        m2 = mc * (1.0 + q) ** 0.2 / q ** 0.6
        out = dict(maps)
        out["mass1"] = q * m2
        out["mass2"] = m2
        return out

    def inverse_transform(self, maps):
        m1 = numpy.asarray(maps["mass1"], dtype=float)
        m2 = numpy.asarray(maps["mass2"], dtype=float)
        out = dict(maps)
        out["mchirp"] = (m1 * m2) ** 0.6 / (m1 + m2) ** 0.2
        out["q"] = m1 / m2
        return out

    def jacobian(self, maps):
        # |d(m1, m2)/d(mchirp, q)| = mchirp ((1+q)/q^3)^(2/5)
        mc = numpy.asarray(maps["mchirp"], dtype=float)
        q = numpy.asarray(maps["q"], dtype=float)
        return mc * ((1.0 + q) / q ** 3) ** 0.4

    def inverse_jacobian(self, maps):
        m1 = numpy.asarray(maps["mass1"], dtype=float)
        m2 = numpy.asarray(maps["mass2"], dtype=float)
        mc = (m1 * m2) ** 0.6 / (m1 + m2) ** 0.2
        return mc / m2 ** 2


class Mass1Mass2ToMchirpQ(_Inverse):
    name = "mass1_mass2_to_mchirp_q"

    def __init__(self):
        super(Mass1Mass2ToMchirpQ, self).__init__(MchirpQToMass1Mass2())
        self.name = "mass1_mass2_to_mchirp_q"


class Logit(BaseTransform):
    """x in (min, max) -> logit((x-min)/(max-min)); the sampling transform the
    reference docs use to remove hard prior edges."""
    name = "logit"

    def __init__(self, inputvar, outputvar, domain=(0., 1.)):
        self.inputs = (inputvar,)
        self.outputs = (outputvar,)
        self._a, self._b = float(domain[0]), float(domain[1])

    def transform(self, maps):
        x = numpy.asarray(maps[self.inputs[0]], dtype=float)
        u = (x - self._a) / (self._b - self._a)
        out = dict(maps)
        with numpy.errstate(divide="ignore", invalid="ignore"):
            out[self.outputs[0]] = numpy.log(u) - numpy.log1p(-u)
        return out

    def inverse_transform(self, maps):
        y = numpy.asarray(maps[self.outputs[0]], dtype=float)
        out = dict(maps)
        out[self.inputs[0]] = self._a + (self._b - self._a) / (1.0 + numpy.exp(-y))
        return out

    def jacobian(self, maps):
        x = numpy.asarray(maps[self.inputs[0]], dtype=float)
        return (self._b - self._a) / ((x - self._a) * (self._b - x))

    def inverse_jacobian(self, maps):
        y = numpy.asarray(maps[self.outputs[0]], dtype=float)
        e = numpy.exp(-numpy.abs(y))
        return (self._b - self._a) * e / (1.0 + e) ** 2


def apply_transforms(samples, transforms, inverse=False):
    """``pycbc.transforms.apply_transforms``: forward order, or reversed order
    with each transform inverted."""
    if inverse:
        transforms = [t.inverse for t in transforms[::-1]]
    out = dict(samples)
    for t in transforms:
        out = t.transform(out)
    return out


def compute_jacobian(samples, transforms, inverse=False):
    """Product of the (inverse) Jacobians along the chain."""
    j = 1.0
    if inverse:
        for t in transforms:
            j = j * t.inverse_jacobian(samples)
    else:
        for t in transforms:
            j = j * t.jacobian(samples)
    return j
