"""Host-side logic that surrounds the kernel: priors, transforms, the model /
CallModel / pool protocols and the ensemble steppers (no GPU needed; the data
model is the oracle-backed test double)."""
import numpy as np
import pytest
from scipy import stats

import util
from gwin_b200 import distributions as D
from gwin_b200 import models, transforms as T
from gwin_b200.pool import BatchPool
from gwin_b200.sampler import (EmceeEnsembleSampler, EmceePTSampler,
                               default_beta_ladder, samplers)
from oracle_model import OracleGaussianNoise


def test_no_prior_and_default_stats():
    """reference test/test_models.py:30-81"""
    p = models.base._NoPrior()
    assert p.apply_boundary_conditions(a=1, b=2) == {'a': 1, 'b': 2}
    assert p() == 0.
    m = models.TestNormal('x')
    call = models.CallModel(m, 'logposterior')
    assert call.variable_params == ('x',)
    assert isinstance(call.prior_distribution, models.base._NoPrior)
    call.update(x=0.3)
    assert call.logjacobian == 0.
    assert call([0.3])[0] == m._dist.logpdf([0.3])
    assert set(['logjacobian', 'logprior', 'loglikelihood']).issubset(call.default_stats)


def test_registry_and_names():
    assert set(models.models) >= {'gaussian_noise', 'marginalized_gaussian_noise', 'test_normal'}
    assert models.models['gaussian_noise'] is models.GaussianNoise
    assert set(samplers) == {'emcee', 'emcee_pt', 'emcee_cuda', 'emcee_pt_cuda'}


def test_analytic_models_batch_equals_scalar():
    """The reference's analytic test models (gwin/models/analytic.py:88-193): the formulas as its
    docstrings state them, and batched evaluation equal to scalar calls."""
    rng = np.random.RandomState(3)
    assert {'test_eggbox', 'test_rosenbrock', 'test_volcano'} <= set(models.models)
    x = rng.uniform(-3, 3, (50, 3))
    for cls, ref in ((models.TestEggbox, lambda p: (2 + np.prod(np.cos(p / 2.))) ** 5),
                     (models.TestRosenbrock, lambda p: -sum((1 - p[i]) ** 2 + 100 * (p[i + 1] - p[i] ** 2) ** 2
                                                            for i in range(len(p) - 1)))):
        m = cls(['a', 'b', 'c'])
        call = models.CallModel(m, 'loglikelihood', return_all_stats=False)
        got = call.batch(x).values
        assert np.allclose(got, [ref(p) for p in x], rtol=1e-13)
        assert np.allclose(got, [call(p) for p in x], rtol=1e-13)
    v = models.TestVolcano(['a', 'b'])
    call = models.CallModel(v, 'loglikelihood', return_all_stats=False)
    r = np.hypot(x[:, 0], x[:, 1])
    ref = 25 * (np.exp(-r / 35) + 1 / (2 * np.sqrt(2 * np.pi)) * np.exp(-0.5 * ((r - 5) / 2) ** 2))
    assert np.allclose(call.batch(x[:, :2]).values, ref, rtol=1e-13)
    assert np.allclose([call(p) for p in x[:5, :2]], ref[:5], rtol=1e-13)
    with pytest.raises(ValueError):
        models.TestVolcano(['a', 'b', 'c'])


def test_distributions_vectorised():
    rng = np.random.RandomState(0)
    u = D.Uniform(a=(1., 3.), b=(-1., 1.))
    x = {'a': np.array([0.5, 2., 3.5]), 'b': np.array([0., 0., 0.])}
    assert np.allclose(u(**x), [-np.inf, np.log(0.25), -np.inf])
    s = D.SinAngle(inclination=None)
    th = np.linspace(0.01, np.pi - 0.01, 7)
    assert np.allclose(s.pdf(inclination=th), 0.5 * np.sin(th))
    assert np.allclose(s.apply_boundary_conditions(inclination=np.array([-0.1, np.pi + 0.2]))['inclination'],
                       [0.1, np.pi - 0.2])
    c = D.CosAngle(dec=None)
    assert np.allclose(c.pdf(dec=np.array([0.0])), 0.5)
    a = D.UniformAngle(ra=None)
    assert np.allclose(a.apply_boundary_conditions(ra=np.array([-0.5, 7.0]))['ra'],
                       [2 * np.pi - 0.5, 7.0 - 2 * np.pi])
    r = D.UniformRadius(distance=(10., 100.))
    d = r.rvs(20000, rng)['distance']
    assert stats.kstest((d ** 3 - 1e3) / (1e6 - 1e3), 'uniform').pvalue > 1e-3
    sky = D.UniformSky()
    j = D.JointDistribution(['ra', 'dec', 'a', 'b'], sky, u)
    draw = j.rvs(1000, rng)
    assert draw.fieldnames == ('ra', 'dec', 'a', 'b')
    assert np.all(np.isfinite(j(**{n: draw[n] for n in draw.fieldnames})))
    with pytest.raises(ValueError):
        D.JointDistribution(['ra'], u)
    jc = D.JointDistribution(['a', 'b'], u, constraints=[lambda p: p['a'] > 2.0])
    assert np.all(jc.rvs(200, rng)['a'] > 2.0)
    assert jc(a=np.array([1.5]), b=np.array([0.]))[0] == -np.inf


def test_transforms_roundtrip_and_jacobian():
    t = T.MchirpQToMass1Mass2()
    x = {'mchirp': np.array([1.2, 20.]), 'q': np.array([1.0, 3.0])}
    y = t.transform(x)
    z = t.inverse_transform(y)
    assert np.allclose(z['mchirp'], x['mchirp']) and np.allclose(z['q'], x['q'])
    assert np.all(y['mass1'] >= y['mass2'])
    # numerical Jacobian
    eps = 1e-6
    for i in range(2):
        def f(mc, q):
            o = t.transform({'mchirp': mc, 'q': q})
            return np.array([o['mass1'], o['mass2']])
        mc, q = x['mchirp'][i], x['q'][i]
        J = np.column_stack([(f(mc + eps, q) - f(mc - eps, q)) / (2 * eps),
                             (f(mc, q + eps) - f(mc, q - eps)) / (2 * eps)])
        assert abs(abs(np.linalg.det(J)) / t.jacobian(x)[i] - 1) < 1e-6
        assert abs(t.jacobian(x)[i] * t.inverse_jacobian(y)[i] - 1) < 1e-9
    lg = T.Logit('x', 'lx', domain=(0., 2.))
    v = lg.inverse_transform(lg.transform({'x': np.array([0.3, 1.7])}))
    assert np.allclose(v['x'], [0.3, 1.7])


def _model(cfg, **kw):
    prior = D.JointDistribution(
        ['tc', 'distance'],
        D.Uniform(tc=(cfg.inj['tc'] - 0.05, cfg.inj['tc'] + 0.05)),
        D.UniformRadius(distance=(50., 1000.)))
    static = {k: v for k, v in cfg.inj.items() if k not in ('tc', 'distance')}
    return OracleGaussianNoise(['tc', 'distance'], cfg, static_params=static, prior=prior, **kw)


def test_batch_equals_scalar_calls_and_short_circuit():
    """evaluate_batch == W x CallModel.__call__ (reference models/__init__.py:101-137),
    including the logprior = -inf short-circuit (base.py:556-568)."""
    cfg = util.c1()
    m = _model(cfg)
    rng = np.random.RandomState(3)
    pts = np.column_stack([cfg.inj['tc'] + rng.uniform(-0.06, 0.06, 12), rng.uniform(20., 1100., 12)])
    for callstat in ('logposterior', 'logplr', 'loglikelihood', 'logprior'):
        call = models.CallModel(m, callstat)
        scalar = [call(p) for p in pts]
        batch = call.batch(pts).rows()
        pool = BatchPool().map(call, pts)
        for (v0, s0), (v1, s1), (v2, s2) in zip(scalar, batch, pool):
            assert v0 == v1 == v2 or abs(v0 - v1) <= 1e-9 * abs(v0)
            for a, b in zip(s0, s1):
                assert (np.isnan(a) and np.isnan(b)) or a == b or abs(a - b) <= 1e-9 * abs(a)
    out = models.CallModel(m, 'logposterior').batch(pts)
    outside = np.isneginf(out.stats['logprior'])
    assert outside.any() and not outside.all()
    assert np.all(np.isneginf(out.values[outside]))
    assert np.all(np.isnan(out.stats['loglikelihood'][outside]))     # never evaluated
    call = models.CallModel(m, 'logposterior', return_all_stats=False)
    assert BatchPool().map(call, pts[:3]) == [call(p) for p in pts[:3]]


def test_pool_falls_back_to_plain_map_and_global_model():
    pool = BatchPool()
    assert pool.map(lambda x: 2 * x, [1, 2, 3]) == [2, 4, 6]
    assert pool.count == 1
    m = models.TestNormal(['x'])
    models._global_instance = models.CallModel(m, 'loglikelihood', return_all_stats=False)
    try:
        got = pool.map(models._call_global_model, [[0.], [1.]])
        assert np.allclose(got, [m._dist.logpdf([0.]), m._dist.logpdf([1.])])
    finally:
        models._global_instance = None


def test_pool_sees_through_sampler_library_wrappers():
    """emcee wraps the log-probability callable in ``_function_wrapper(f, args, kwargs)`` before
    mapping it; a PT sampler maps separate prior / likelihood callables
    (reference sampler/emcee.py:230-250).  Both must still reach the batched path."""
    from gwin_b200.sampler.emcee import _callloglikelihood, _callprior

    class FunctionWrapper(object):          # shape of emcee.ensemble._function_wrapper
        def __init__(self, f, args=(), kwargs=None):
            self.f, self.args, self.kwargs = f, args, kwargs or {}

        def __call__(self, x):
            return self.f(x, *self.args, **self.kwargs)

    prior = D.JointDistribution(['x'], D.Uniform(x=(-5, 5)))
    m = models.TestNormal(['x'], prior=prior)
    call = models.CallModel(m, 'logposterior')
    pts = np.linspace(-6, 6, 9)[:, None]
    pool = BatchPool()
    got = pool.map(FunctionWrapper(call), pts)
    assert pool.ncalls == 1                                   # one batched evaluation
    ref = [call(p) for p in pts]
    for (v, blob), (v0, blob0) in zip(got, ref):
        assert v == v0
        assert all(a == b or (np.isnan(a) and np.isnan(b)) for a, b in zip(blob, blob0))
    assert got[0][0] == -np.inf and np.isnan(got[0][1][2])    # likelihood never evaluated
    bound = pool.map(FunctionWrapper(call, args=('logprior',)), pts)      # extra args: plain map
    assert pool.ncalls == 1 and len(bound) == 9
    ll = pool.map(_callloglikelihood(call), pts)
    lp = pool.map(_callprior(call), pts)
    assert pool.ncalls == 3
    assert np.allclose(ll, [m._dist.logpdf(p) for p in pts])
    assert lp == [call(p, callstat='logprior', return_all_stats=False) for p in pts]


def test_ensemble_sampler_recovers_normal():
    np.random.seed(11)
    prior = D.JointDistribution(['x', 'y'], D.Uniform(x=(-20, 20), y=(-20, 20)))
    cov = [[1., .5], [.5, 2.]]
    m = models.TestNormal(['x', 'y'], mean=[1., -2.], cov=cov, prior=prior)
    s = EmceeEnsembleSampler(m, 200)
    p0 = s.set_p0()
    assert p0.shape == (200, 2)
    s.run(250)
    assert s.chain.shape == (200, 250, 2) and s.lnpost.shape == (200, 250)
    assert s.model_stats.shape == (200, 250)
    c = s.chain[:, 120::10, :].reshape(-1, 2)
    assert np.all(np.abs(c.mean(0) - [1., -2.]) < 0.1)
    assert np.allclose(np.cov(c.T), cov, atol=0.25)
    assert 0.5 < s.acceptance_fraction.mean() < 0.9
    assert stats.kstest(c[:, 0] - 1., stats.norm.cdf).pvalue > 1e-3
    s.clear_chain()
    assert s.chain.shape[1] == 0 and s.niterations == 250


def test_pt_sampler_and_ladder():
    assert np.allclose(default_beta_ladder(2, 3), [1., 1 / 7., 1 / 49.])
    np.random.seed(5)
    prior = D.JointDistribution(['x'], D.Uniform(x=(-10, 10)))
    m = models.TestNormal(['x'], mean=[0.5], cov=[0.25], prior=prior)
    pt = EmceePTSampler(m, 3, 40)
    assert pt.set_p0().shape == (3, 40, 1)
    pt.run(300)
    assert pt.chain.shape == (3, 40, 300, 1)
    cold = pt.chain[0][:, 100:, 0].ravel()
    assert abs(cold.mean() - 0.5) < 0.05 and abs(cold.std() - 0.5) < 0.06
    hot = pt.chain[2][:, 100:, 0].ravel()
    assert hot.std() > cold.std()
    ms = pt.model_stats
    assert ms['loglr'].shape == (3, 40, 300)


def test_data_model_sampler_smoke():
    """reference test/test_sampler.py: every sampler runs set_p0 -> run(4)."""
    cfg = util.c1()
    m = _model(cfg)
    np.random.seed(2)
    for s in (EmceeEnsembleSampler(m, 30, model_call=models.CallModel(m, 'logplr')),
              EmceePTSampler(m, 2, 30)):
        s.set_p0()
        s.run(4)
        assert s.niterations == 4
        assert np.all(np.isfinite(s.lnpost))


def test_checkpoint_resume_is_bit_identical(tmp_path):
    """Result file with the reference's tree (samples/, model_stats/, acceptance_fraction,
    sampler_states/) and resume: 10 iterations, write, resume in a NEW sampler for 10 more ==
    20 iterations straight (the reference restores positions + RNG state the same way,
    bin/gwin:263-285, sampler/emcee.py:160-167)."""
    from gwin_b200.io import InferenceFile, check_integrity
    prior = D.JointDistribution(['x', 'y'], D.Uniform(x=(-20, 20), y=(-20, 20)))

    def make():
        return models.TestNormal(['x', 'y'], mean=[1., -2.], cov=[[1., .5], [.5, 2.]], prior=prior)
    np.random.seed(21)
    ref = EmceeEnsembleSampler(make(), 40)
    ref.set_p0()
    ref.run(20)
    path = str(tmp_path / "run.npz")
    np.random.seed(21)
    a = EmceeEnsembleSampler(make(), 40)
    a.set_p0()
    a.run(10)
    with InferenceFile(path, 'w') as fp:
        a.write_results(fp, approximant="none")
    assert check_integrity(path)
    fp = InferenceFile(path, 'r')
    assert fp.sampler_name == 'emcee' and fp.niterations == 10 and fp.variable_params == ['x', 'y']
    assert fp['samples/x'].shape == (40, 10) and fp['model_stats/loglikelihood'].shape == (40, 10)
    assert fp['acceptance_fraction'].shape == (40,)
    assert fp.read_samples(['x'], thin_start=4, thin_interval=2)['x'].shape == (40 * 3,)
    np.random.seed(999)                       # whatever the global state is now ...
    b = EmceeEnsembleSampler(make(), 40)
    b.set_p0(samples_file=fp)
    b.set_state_from_file(fp)                 # ... the stepper continues from the saved state
    b.run(10)
    assert b.niterations == 20
    np.testing.assert_array_equal(b.chain, ref.chain[:, 10:, :])
    with InferenceFile(path, 'a') as fp2:
        b.write_results(fp2)
    fp3 = InferenceFile(path, 'r')
    assert fp3['samples/x'].shape == (40, 20) and fp3.niterations == 20
    np.testing.assert_array_equal(fp3['samples/x'], ref.chain[..., 0])
    assert check_integrity(path)
    # PT layout: ntemps x nwalkers x niterations
    np.random.seed(3)
    pt = EmceePTSampler(make(), 2, 20)
    pt.set_p0()
    pt.run(5)
    p2 = str(tmp_path / "pt.npz")
    with InferenceFile(p2, 'w') as fp:
        pt.write_results(fp)
    fp = InferenceFile(p2, 'r')
    assert fp['samples/y'].shape == (2, 20, 5) and fp.attrs['ntemps'] == 2
    assert fp['model_stats/loglr'].shape == (2, 20, 5)


def test_calibration_description_classes():
    """gwin/calibration.py: constructor, spline points, parameter naming, from_config,
    the values matrix the kernels take."""
    from gwin_b200 import calibration as cal
    cs = cal.CubicSpline("20", "1024", "5", "h1")
    assert cs.n_points == 5 and np.allclose(cs.spline_points, np.logspace(np.log10(20.), np.log10(1024.), 5))
    assert cs.param_names[:2] == ["recalib_amplitude_h1_0", "recalib_amplitude_h1_1"]
    assert cs.param_names[5] == "recalib_phase_h1_0"
    with pytest.raises(ValueError):
        cal.CubicSpline(20., 1024., 3, "h1")
    with pytest.raises(NotImplementedError):            # applied on the GPU only
        cs.map_to_adjust(np.zeros(4, complex))

    class CP(object):
        def items(self, section):
            assert section == "calibration"
            return [("h1_model", "cubic_spline"), ("h1_minimum_frequency", "20"),
                    ("h1_maximum_frequency", "512"), ("h1_n_points", "6"),
                    ("l1_model", "cubic_spline"), ("l1_minimum_frequency", "10"),
                    ("l1_maximum_frequency", "512"), ("l1_n_points", "6")]
    m = cal.Recalibrate.from_config(CP(), "H1", "calibration")
    assert isinstance(m, cal.CubicSpline) and m.ifo_name == "h1" and m.n_points == 6
    assert np.isclose(m.spline_points[0], 20.) and np.isclose(m.spline_points[-1], 512.)
    rec = {"H1": m, "L1": cal.Recalibrate.from_config(CP(), "L1", "calibration")}
    sf = cal.spline_frequencies(rec, ["H1", "L1"])
    assert sf.shape == (2, 6) and np.isclose(sf[1, 0], 10.)
    params = {nm: np.arange(3) * 0.01 for nm in rec["H1"].param_names}
    static = {nm: 0.02 for nm in rec["L1"].param_names}
    V = cal.values_matrix(rec, ["H1", "L1"], params, 3, static=static)
    assert V.shape == (24, 3) and np.all(V[:12, 2] == 0.02) and np.all(V[12:] == 0.02)
    with pytest.raises(KeyError):
        cal.values_matrix(rec, ["H1", "L1"], params, 3)
    with pytest.raises(ValueError):
        cal.spline_frequencies({"H1": m, "L1": cal.CubicSpline(20, 512, 7, "l1")}, ["H1", "L1"])


def test_parameters_the_kernel_does_not_evaluate_are_refused():
    """The reference forwards every sampled / static parameter to ``get_fd_waveform``
    (models/base_data.py:84-161); the kernel evaluates the 13 TaylorF2 parameters.  Anything
    else must raise instead of being ignored (a flat marginal, or a likelihood that differs from
    the reference's without a word)."""
    from gwin_b200 import transforms as T
    from gwin_b200.waveform import (FDomainCBCGenerator, FDomainDetFrameGenerator,
                                    NoWaveformError, check_waveform_params)
    check_waveform_params(("mass1", "tc"), {"mass2": 1.4, "f_ref": 0.0, "phase_order": -1,
                                             "approximant": "TaylorF2", "f_lower": 20.})
    check_waveform_params(("mchirp", "q", "tc"), {}, [T.MchirpQToMass1Mass2()])
    check_waveform_params(("tc", "recalib_amplitude_h1_0"), {}, None, ["recalib_amplitude_h1_0"])
    for variable, static in ((("spin1x",), {}), (("eccentricity",), {}), (("mchirp",), {}),
                             ((), {"phase_order": 4}), ((), {"f_final": 512.}), ((), {"f_ref": 20.}),
                             ((), {"spin1_a": 0.3}), ((), {"spin2y": 0.1})):
        with pytest.raises(ValueError):
            check_waveform_params(variable, static)
    with pytest.raises(ValueError):
        FDomainDetFrameGenerator(FDomainCBCGenerator, 0., detectors=["H1"], variable_args=("tc",),
                                 delta_f=0.25, f_lower=20., approximant="TaylorF2", f_final=300.)
    gen = FDomainDetFrameGenerator(FDomainCBCGenerator, 0., detectors=["H1"], variable_args=("tc",),
                                   delta_f=0.25, f_lower=20., approximant="TaylorF2", f_ref=0.)
    # zero total mass is "no waveform", not a ZeroDivisionError
    with pytest.raises(NoWaveformError):
        gen.generate(mass1=0., mass2=0., ra=0., dec=0., polarization=0., tc=1.)


class _KombineShapedSampler(object):
    """The call pattern of ``kombine.Sampler`` as the reference drives it
    (``gwin/sampler/kombine.py:81-84, 142-147``): it is given ``pool`` and ``processes = pool.count``,
    the wrapper first computes ``blob0 = [model(p0[wi, :])[1] for wi in range(nwalkers)]`` with
    scalar calls, and every iteration maps the log-posterior callable over ALL walkers' proposals,
    each call returning ``(lnpost, blob)``."""

    def __init__(self, nwalkers, ndim, lnpostfn, pool=None, processes=1):
        self.nwalkers, self.ndim, self.lnpostfn = nwalkers, ndim, lnpostfn
        self.pool, self.processes = pool, processes
        self.blobs = []

    def run_mcmc(self, niter, p0, blob0, rng):
        p, blob = np.array(p0), list(blob0)
        lnpost = np.array([self.lnpostfn(x)[0] for x in p])
        for _ in range(niter):
            prop = p + 0.3 * rng.standard_normal(p.shape)      # (kombine draws from its KDE; any proposal does here)
            results = list(self.pool.map(self.lnpostfn, prop))
            lnq = np.array([r[0] for r in results])
            acc = np.log(rng.uniform(size=self.nwalkers)) < lnq - lnpost
            p[acc], lnpost[acc] = prop[acc], lnq[acc]
            blob = [results[i][1] if acc[i] else blob[i] for i in range(self.nwalkers)]
            self.blobs.append(list(blob))
        return p, lnpost


def test_kombine_call_pattern_through_the_pool():
    prior = D.JointDistribution(['x', 'y'], D.Uniform(x=(-5, 5), y=(-5, 5)))
    m = models.TestNormal(['x', 'y'], prior=prior)
    call = models.CallModel(m, 'logposterior')            # return_all_stats: (value, blob) per point
    pool = BatchPool()
    pool.count = 4                                        # what sampler_from_cli sets from opts.nprocesses
    nwalkers = 16
    s = _KombineShapedSampler(nwalkers, 2, call, pool=pool, processes=pool.count)
    assert s.processes == 4
    rng = np.random.default_rng(5)
    p0 = rng.uniform(-1, 1, (nwalkers, 2))
    blob0 = [call(p0[wi, :])[1] for wi in range(nwalkers)]           # kombine.py:145-147
    assert all(len(b) == len(m.default_stats) for b in blob0)
    p, lnpost = s.run_mcmc(5, p0, blob0, rng)
    assert pool.ncalls == 5 and pool.npoints == 5 * nwalkers          # one batched evaluation per iteration
    # the blobs carried with the accepted positions are the scalar calls' blobs
    for wi in range(nwalkers):
        v, b = call(p[wi])
        assert v == lnpost[wi]
        assert all(x == y or (np.isnan(x) and np.isnan(y)) for x, y in zip(b, s.blobs[-1][wi]))


def test_sampler_from_cli_selects_pool_and_global_instance():
    """option_utils.sampler_from_cli (reference gwin/option_utils.py:150-185)."""
    from types import SimpleNamespace
    from gwin_b200 import option_utils
    prior = D.JointDistribution(['x'], D.Uniform(x=(-5, 5)))
    m = models.TestNormal(['x'], prior=prior)
    try:
        opts = SimpleNamespace(sampler="emcee", logpost_function="logposterior", nprocesses=1,
                               use_mpi=False, nwalkers=8)
        models._global_instance = None
        s = option_utils.sampler_from_cli(opts, m)
        assert isinstance(s, EmceeEnsembleSampler)
        assert models._global_instance is None                     # one process: the model is called directly
        assert isinstance(s.model, models.CallModel) and s.model.callstat == "logposterior"
        assert isinstance(s._pool, BatchPool) and s._pool.count == 1
        opts.nprocesses = 6
        s = option_utils.sampler_from_cli(opts, m)
        assert isinstance(models._global_instance, models.CallModel)
        assert s._model_call is models._call_global_model           # workers would call it by name
        assert s._pool.count == 6
        np.random.seed(2)
        s.set_p0()
        s.run(20)
        assert s.chain.shape == (8, 20, 1)
        opts.sampler, opts.ntemps = "emcee_pt", 2
        s = option_utils.sampler_from_cli(opts, m)
        assert isinstance(s, EmceePTSampler)
        with pytest.raises(KeyError):
            opts.sampler = "no_such_sampler"
            option_utils.sampler_from_cli(opts, m)
    finally:
        models._global_instance = None


def test_validate_checkpoint_files(tmp_path):
    """option_utils.validate_checkpoint_files (reference gwin/option_utils.py:192-283): checkpoint and
    backup end up both usable or both unusable."""
    from gwin_b200 import option_utils
    from gwin_b200.io import InferenceFile
    prior = D.JointDistribution(['x'], D.Uniform(x=(-5, 5)))
    m = models.TestNormal(['x'], prior=prior)
    np.random.seed(1)
    s = EmceeEnsembleSampler(m, 8)
    s.set_p0()
    s.run(6)
    chk, bak = str(tmp_path / "run.npz"), str(tmp_path / "run.bak.npz")
    assert not option_utils.validate_checkpoint_files(chk, bak)          # neither exists: start from scratch
    with InferenceFile(chk, 'w') as fp:
        s.write_metadata(fp)
        s.write_results(fp)
    assert option_utils.validate_checkpoint_files(chk, bak)              # the backup is created from the checkpoint
    assert InferenceFile(bak, 'r')['samples/x'].shape == (8, 6)
    with open(chk, 'wb') as f:                                           # a checkpoint cut off mid-write
        f.write(b'not an archive')
    assert option_utils.validate_checkpoint_files(chk, bak)              # ... is restored from the backup
    assert InferenceFile(chk, 'r')['samples/x'].shape == (8, 6)
    s.run(4)                                                             # the checkpoint moves on, the backup lags
    with InferenceFile(chk, 'a') as fp:
        s.write_metadata(fp)
        s.write_results(fp)
    assert InferenceFile(chk, 'r')['samples/x'].shape[-1] != InferenceFile(bak, 'r')['samples/x'].shape[-1]
    assert option_utils.validate_checkpoint_files(chk, bak)
    assert InferenceFile(bak, 'r')['samples/x'].shape == InferenceFile(chk, 'r')['samples/x'].shape
    for p in (chk, bak):
        with open(p, 'wb') as f:
            f.write(b'')
    assert not option_utils.validate_checkpoint_files(chk, bak)
