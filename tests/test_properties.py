"""Property tests (hypothesis) of the host pieces the device kernels mirror: every distribution the
device prior kernel knows integrates to one and its boundary conditions are idempotent; the mass
transforms invert each other and their Jacobians are reciprocal; the closed-form choice of the tails'
fine sub-blocks tiles any stretch; the sharding helper covers every point once."""
import numpy as np
from hypothesis import assume, given, settings, strategies as st
from scipy import integrate

from gwin_b200 import distributions as D
from gwin_b200 import transforms as T
from kernel_model import PlanModel

finite = dict(allow_nan=False, allow_infinity=False)


@settings(max_examples=40, deadline=None)
@given(lo=st.floats(-3.0, 3.0, **finite), width=st.floats(0.05, 4.0, **finite),
       mu=st.floats(-2.0, 2.0, **finite), sd=st.floats(0.05, 3.0, **finite))
def test_densities_integrate_to_one(lo, width, mu, sd):
    hi = lo + width
    # (a normal truncated to an interval more than ~35 sigma out underflows in linear space: not a prior)
    assume(max(lo - mu, mu - hi, 0.0) <= 20.0 * sd)
    cases = [(D.Uniform(x=(lo, hi)), lo, hi),
             (D.Gaussian(x=(lo, hi), x_mean=mu, x_var=sd * sd), lo, hi),
             (D.UniformRadius(x=(abs(lo) + 0.1, abs(lo) + 0.1 + width)), abs(lo) + 0.1, abs(lo) + 0.1 + width),
             (D.SinAngle(x=None), 0.0, np.pi), (D.CosAngle(x=None), -np.pi / 2, np.pi / 2),
             (D.UniformAngle(x=None), 0.0, 2 * np.pi)]
    for dist, a, b in cases:
        val, _ = integrate.quad(lambda t: float(np.exp(dist(x=np.array([t]))[0])), a, b, limit=200)
        assert abs(val - 1.0) < 1e-7, (type(dist).__name__, val)
        # nothing outside the support
        assert dist(x=np.array([a - 0.3 * (b - a)]))[0] == -np.inf or isinstance(dist, (D.UniformAngle,))


@settings(max_examples=60, deadline=None)
@given(x=st.floats(-40.0, 40.0, **finite))
def test_boundary_conditions_are_idempotent_and_land_in_the_support(x):
    for dist, a, b in ((D.SinAngle(t=None), 0.0, np.pi), (D.CosAngle(t=None), -np.pi / 2, np.pi / 2),
                       (D.UniformAngle(t=None), 0.0, 2 * np.pi)):
        once = dist.apply_boundary_conditions(t=np.array([x]))['t']
        twice = dist.apply_boundary_conditions(t=once)['t']
        assert a - 1e-12 <= once[0] <= b + 1e-12
        assert np.allclose(once, twice, atol=1e-12)
        assert np.isfinite(dist(t=once)[0]) or dist.pdf(t=once)[0] == 0.0


@settings(max_examples=60, deadline=None)
@given(mc=st.floats(0.5, 60.0, **finite), q=st.floats(1.0, 12.0, **finite))
def test_mass_transforms_invert_and_jacobians_are_reciprocal(mc, q):
    fwd, inv = T.MchirpQToMass1Mass2(), T.Mass1Mass2ToMchirpQ()
    m = fwd.transform({'mchirp': np.array([mc]), 'q': np.array([q])})
    back = inv.transform(m)
    assert np.allclose(back['mchirp'], mc, rtol=1e-12) and np.allclose(back['q'], q, rtol=1e-12)
    assert m['mass1'][0] >= m['mass2'][0] * (1 - 1e-12)
    j_f = fwd.jacobian({'mchirp': np.array([mc]), 'q': np.array([q])})
    j_b = fwd.inverse_jacobian(m)
    assert np.allclose(j_f * j_b, 1.0, rtol=1e-10)
    # numerical determinant of d(m1, m2)/d(mchirp, q)
    h = 1e-6
    cols = []
    for dmc, dq in ((h, 0.0), (0.0, h)):
        p = fwd.transform({'mchirp': np.array([mc * (1 + dmc)]), 'q': np.array([q * (1 + dq)])})
        n = fwd.transform({'mchirp': np.array([mc * (1 - dmc)]), 'q': np.array([q * (1 - dq)])})
        step = 2 * h * (mc if dmc else q)
        cols.append([(p['mass1'][0] - n['mass1'][0]) / step, (p['mass2'][0] - n['mass2'][0]) / step])
    det = abs(cols[0][0] * cols[1][1] - cols[0][1] * cols[1][0])
    assert abs(det - j_f[0]) <= 1e-6 * j_f[0]


@settings(max_examples=200, deadline=None)
@given(k=st.integers(7, 13), three=st.booleans(), fk=st.integers(5, 8), frac=st.floats(0.0, 1.0, **finite))
def test_fine_decomposition_tiles_any_stretch(k, three, fk, frac):
    pm = (3 << (k - 1)) if three else (1 << k)               # sub-block sizes are 2^k and 3 2^(k-1)
    fmin = (3 << (fk - 1)) if three else (1 << fk)
    if fmin > pm // 2:
        fmin = pm // 2
    while pm % fmin or (pm // fmin) & (pm // fmin - 1):      # fine sizes are pm / 2^j
        fmin //= 2
    length = int(frac * (pm - 1))
    taken, left = PlanModel.fine_decomposition(length, pm, fmin)
    pos = 0
    for off, size in taken:
        assert off == pos and size >= fmin and pm % size == 0 and off % size == 0
        pos += size
    assert pos + left == length and 0 <= left < fmin
    assert len(taken) <= int(np.log2(pm // fmin)) + 1


@settings(max_examples=100, deadline=None)
@given(W=st.integers(0, 300), world=st.integers(1, 8))
def test_round_robin_shards_cover_every_point_once(W, world):
    # WalkerSharding.my_slice / count without a process group: the arithmetic only
    seen = np.zeros(W, dtype=int)
    for rank in range(world):
        idx = np.arange(W)[slice(rank, W, world)]
        seen[idx] += 1
        assert len(idx) == max(0, -(-(W - rank) // world))
    assert np.all(seen == 1)
