"""Detector geometry handed to the CUDA plan.

Only constants live here: vertex location and response tensor of each
interferometer (values of LALDetectors.h: ``response = (x(x)x - y(x)y)/2`` from
the arm unit vectors, stored by LAL as REAL4 -- we round to float32 the same
way).  The antenna pattern and time delay themselves
(``pycbc.detector.Detector.antenna_pattern`` /
``time_delay_from_earth_center``, called inside
``FDomainDetFrameGenerator.generate`` -> reference ``gaussian_noise.py:322``)
are evaluated per walker on the GPU (``walker_setup_kernel``).
"""
import numpy

_GEOMETRY = {
    # name: (vertex [m], x-arm unit vector, y-arm unit vector), Earth-fixed frame
    "H1": ((-2.16141492636e+06, -3.83469517889e+06, 4.60035022664e+06),
           (-0.22389266154, +0.79983062746, +0.55690487831),
           (-0.91397818574, +0.02609403989, -0.40492342125)),
    "L1": ((-7.42760447238e+04, -5.49628371971e+06, 3.22425701744e+06),
           (-0.95457412153, -0.14158077340, -0.26218911324),
           (+0.29774156894, -0.48791033647, -0.82054461286)),
    "V1": ((4.54637409900e+06, 8.42989697626e+05, 4.37857696241e+06),
           (-0.70045821479, +0.20848948619, +0.68256166277),
           (-0.05379255368, -0.96908180549, +0.24080451708)),
    # KAGRA (LAL_KAGRA_*), used for the 4-detector code path
    "K1": ((-3777336.024, 3484898.411, 3765313.697),
           (-0.3759040, -0.8361583, 0.3994189),
           (0.7164378, 0.01114076, 0.6975620)),
}


class Detector(object):
    """Holds ``name``, ``location`` [3] and ``response`` [3, 3]."""

    def __init__(self, name):
        try:
            loc, xarm, yarm = _GEOMETRY[name]
        except KeyError:
            raise ValueError("unknown detector %r (known: %s)"
                             % (name, ", ".join(sorted(_GEOMETRY))))
        self.name = name
        self.location = numpy.array(loc, dtype=numpy.float64)
        xarm = numpy.array(xarm)
        yarm = numpy.array(yarm)
        resp = 0.5 * (numpy.outer(xarm, xarm) - numpy.outer(yarm, yarm))
        self.response = resp.astype(numpy.float32).astype(numpy.float64)

    def __repr__(self):
        return "Detector(%r)" % self.name


def available_detectors():
    return sorted(_GEOMETRY)
