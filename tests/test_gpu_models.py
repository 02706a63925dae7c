"""The drop-in boundary on the GPU: model protocol, CallModel, BatchPool, samplers
(``-m gpu``).  The likelihood behind every call here is the CUDA library."""
import numpy as np
import pytest
from scipy import stats

import gwin_oracle as O
import util
from gwin_b200 import distributions as D
from gwin_b200 import models
from gwin_b200.pool import BatchPool
from gwin_b200.sampler import EmceeEnsembleSampler, EmceePTSampler
from gwin_b200.types import FrequencySeries
from gwin_b200.waveform import (FDomainCBCGenerator, FDomainDetFrameGenerator,
                                NoWaveformError)
from oracle_model import OracleGaussianNoise
from util import PARAM_ORDER, tol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg():
    return util.c2()


def test_scalar_protocol_matches_oracle(cfg):
    """update(**params) then the cached stat properties, as CallModel drives them."""
    model = cfg.cuda_model()
    orc = cfg.oracle()
    p = dict(cfg.inj)
    p["distance"] = 420.0
    model.update(**p)
    lr0, hd0, hh0 = orc.loglr(**p)
    assert abs(model.loglr - lr0) <= tol(lr0)
    assert abs(model.lognl - orc.lognl()) <= 1e-9 * abs(orc.lognl())
    assert abs(model.loglikelihood - (lr0 + orc.lognl())) <= 1e-9 * abs(orc.lognl())
    assert model.logprior == 0. and model.logjacobian == 0.
    assert abs(model.logplr - lr0) <= tol(lr0)
    for det in cfg.dets:
        assert abs(model.det_optimal_snrsq(det) - hh0[det]) <= 1e-9 * hh0[det]
        assert abs(model.det_cplx_loglr(det) - (hd0[det] - 0.5 * hh0[det])) <= 2e-6
        assert abs(model.det_lognl(det) - orc.det_lognl(det)) <= 1e-9 * abs(orc.det_lognl(det))
    names = model.default_stats
    assert names[:4] == ['logjacobian', 'logprior', 'loglikelihood', 'loglr']
    assert set(names[4:]) == set(['%s_cplx_loglr' % d for d in cfg.dets] +
                                 ['%s_optimal_snrsq' % d for d in cfg.dets])
    st = model.current_stats
    assert abs(st['loglr'] - lr0) <= tol(lr0)
    # the model whitened its own copy of the data, not the caller's
    assert np.allclose(model.data['H1'].numpy()[200], cfg.data['H1'][200] * np.sqrt(4 * cfg.delta_f / cfg.psd[200]))
    # no waveform -> -inf and the reference's placeholder stats (gaussian_noise.py:293-302)
    model.update(**dict(p, mass1=200., mass2=200.))
    assert model.loglr == -np.inf
    assert model.current_stats['H1_optimal_snrsq'] == 0.


def test_constructor_errors(cfg):
    gen = FDomainDetFrameGenerator(FDomainCBCGenerator, 0., detectors=['H1', 'L1'],
                                   variable_args=PARAM_ORDER, delta_f=cfg.delta_f, f_lower=20.)
    data = {d: FrequencySeries(cfg.data[d], cfg.delta_f, 0.) for d in cfg.dets}
    with pytest.raises(ValueError):        # detectors differ (gaussian_noise.py:227-233)
        models.GaussianNoise(PARAM_ORDER, data, gen, 20.)
    gen3 = FDomainDetFrameGenerator(FDomainCBCGenerator, 5., detectors=cfg.dets,
                                    variable_args=PARAM_ORDER, delta_f=cfg.delta_f, f_lower=20.)
    with pytest.raises(ValueError):        # epoch differs (:235-238)
        models.GaussianNoise(PARAM_ORDER, data, gen3, 20.)
    gen0 = FDomainDetFrameGenerator(FDomainCBCGenerator, 0., detectors=cfg.dets,
                                    variable_args=PARAM_ORDER, delta_f=cfg.delta_f, f_lower=20.)
    bad = dict(data)
    bad['V1'] = FrequencySeries(cfg.data['V1'][:-10], cfg.delta_f, 0.)
    with pytest.raises(ValueError):        # lengths differ (:240-242)
        models.GaussianNoise(PARAM_ORDER, bad, gen0, 20.)
    with pytest.raises(ValueError):        # kmax <= kmin
        models.GaussianNoise(PARAM_ORDER, data, gen0, 20., f_upper=10.)
    with pytest.raises(ValueError):
        FDomainDetFrameGenerator(FDomainCBCGenerator, 0., detectors=cfg.dets, delta_f=1., f_lower=20.,
                                 approximant="IMRPhenomPv2")
    with pytest.raises(AttributeError):    # no marginalisation flag (marginalized...:258-261)
        models.MarginalizedGaussianNoise(PARAM_ORDER, data, gen0, 20., psds=cfg.psds())
    with pytest.raises(AttributeError):    # no marg_prior (:265-268)
        models.MarginalizedGaussianNoise(PARAM_ORDER, data, gen0, 20., psds=cfg.psds(),
                                         phase_marginalization=True)
    with pytest.raises(NotImplementedError):
        models.MarginalizedGaussianNoise(PARAM_ORDER, data, gen0, 20., psds=cfg.psds(),
                                         time_marginalization=True,
                                         marg_prior=[D.Uniform(time=(0., 1.))])


def test_marginalized_model_reference_kwargs(cfg):
    model = cfg.cuda_model(cls=models.MarginalizedGaussianNoise, phase_marginalization=True,
                           marg_prior=[D.UniformAngle(phase=None)])
    orc = cfg.oracle(phase_marginalization=True)
    p = dict(cfg.inj)
    model.update(**p)
    lr0, hd0, hh0 = orc.loglr(**p)
    assert abs(model.loglr - lr0) <= tol(lr0)
    st = model.current_stats
    for det in cfg.dets:
        assert abs(st['%s_optimal_snrsq' % det] - hh0[det]) <= 1e-9 * hh0[det]
        assert abs(st['%s_matchedfilter_snrsq' % det] - hd0[det]) <= 2e-6
    assert model.phase_prior.shape == (10 ** 4,)
    # coa_phase drops out of the marginalised statistic
    model.update(**dict(p, coa_phase=2.9))
    assert abs(model.loglr - lr0) <= 1e-6


@pytest.mark.parametrize("phase", [False, True])
def test_distance_marginalized_model(cfg, phase):
    """marginalized_gaussian_noise.py:232-272, 355-389, 431-448 through the model protocol."""
    priors = [D.Uniform(distance=(50, 5000))] + ([D.UniformAngle(phase=None)] if phase else [])
    model = cfg.cuda_model(cls=models.MarginalizedGaussianNoise, distance_marginalization=True,
                           phase_marginalization=phase, marg_prior=priors)
    assert model._dist_array.shape == (10 ** 4,) and model.dist_prior[-1] == 0.
    orc = cfg.oracle(phase_marginalization=phase,
                     distance_marginalization=(model._dist_array, model._deltad * model.dist_prior))
    p = dict(cfg.inj, distance=1.0)
    model.update(**p)
    lr0, hd0, hh0 = orc.loglr(**p)
    assert abs(model.loglr - lr0) <= tol(lr0)
    assert abs(model.loglikelihood - (lr0 + orc.lognl())) <= 1e-9 * abs(orc.lognl())
    st = model.current_stats
    for det in cfg.dets:
        assert abs(st['%s_optimal_snrsq' % det] - hh0[det]) <= 1e-9 * hh0[det]
        assert abs(st['%s_matchedfilter_snrsq' % det] - hd0[det]) <= 2e-9 * hh0[det]
    # batch protocol
    rng = np.random.default_rng(5)
    P = util.draw_bbh(rng, 8, cfg)
    P[:, 4] = 1.0
    res = models.CallModel(model, "loglr", return_all_stats=False).batch(P)
    ref = [orc.loglr(**dict(zip(PARAM_ORDER, row)))[0] for row in P]
    assert np.all(np.abs(res.values - ref) <= tol(np.array(ref)))
    with pytest.raises(AttributeError):    # one prior per marginalised parameter (:364-366)
        cfg.cuda_model(cls=models.MarginalizedGaussianNoise, distance_marginalization=True,
                       phase_marginalization=True, marg_prior=[D.Uniform(distance=(50, 5000))])


def test_pool_map_equals_scalar_calls(cfg):
    prior = D.JointDistribution(
        ['tc', 'distance', 'inclination'],
        D.Uniform(tc=(cfg.inj['tc'] - 0.05, cfg.inj['tc'] + 0.05)),
        D.UniformRadius(distance=(50., 1000.)), D.SinAngle(inclination=None))
    static = {k: v for k, v in cfg.inj.items() if k not in ('tc', 'distance', 'inclination')}
    model = cfg.cuda_model(variable_params=('tc', 'distance', 'inclination'),
                           static_params=static, prior=prior)
    rng = np.random.RandomState(4)
    pts = np.column_stack([cfg.inj['tc'] + rng.uniform(-0.06, 0.06, 64), rng.uniform(20., 1100., 64),
                           rng.uniform(-0.2, 3.3, 64)])
    call = models.CallModel(model, 'logplr')
    pool = BatchPool()
    mapped = pool.map(call, pts)
    assert pool.ncalls == 1 and model.plan.last_launches >= 3
    for p, (v, blob) in zip(pts[:16], mapped[:16]):
        v0, blob0 = call(p)
        assert v == v0 or abs(v - v0) <= 1e-9 * abs(v0)
        for a, b in zip(blob, blob0):
            assert (np.isnan(a) and np.isnan(b)) or a == b or abs(a - b) <= 1e-9 * abs(a) + 1e-9
    vals = np.array([m[0] for m in mapped])
    assert np.isneginf(vals).any() and np.isfinite(vals).any()


def test_generate_matches_oracle_waveform(cfg):
    gen = FDomainDetFrameGenerator(FDomainCBCGenerator, 0., detectors=cfg.dets,
                                   variable_args=PARAM_ORDER, delta_f=cfg.delta_f, f_lower=20.)
    h = gen.generate(**cfg.inj)
    ref = O.FDomainDetFrameGenerator(0., cfg.dets, cfg.delta_f, 20.).generate(**cfg.inj)
    for d in cfg.dets:
        assert len(h[d]) == len(ref[d])
        scale = np.abs(ref[d]).max()
        assert np.max(np.abs(h[d].numpy() - ref[d])) <= 1e-8 * scale
    with pytest.raises(NoWaveformError):
        gen.generate(**dict(cfg.inj, mass1=300., mass2=300.))


def test_device_resident_call_equals_host_call(cfg):
    import torch
    model = cfg.cuda_model()
    rng = np.random.default_rng(8)
    P = util.draw_bbh(rng, 300, cfg)
    a = model.plan.loglr_host(P)
    lr, hd, hh = model.plan.loglr_device(torch.tensor(P, device="cuda"))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(lr.cpu().numpy(), a[0])
    np.testing.assert_array_equal(hh.cpu().numpy(), a[2])
    np.testing.assert_array_equal(hd.cpu().numpy()[..., 0], a[1].real)


def test_posterior_consistency_ks_vs_oracle():
    """Fixed-seed emcee-semantics run on C1 with the CUDA likelihood against the same
    sampler driven by the CPU oracle: two-sample KS p > 0.01 per parameter
    (BASELINE.json north_star; scipy.stats.ks_2samp as in gwin/burn_in.py:23,64)."""
    cfg = util.c1()
    names = ['tc', 'distance', 'inclination', 'coa_phase']
    prior = D.JointDistribution(
        names, D.Uniform(tc=(cfg.inj['tc'] - 0.02, cfg.inj['tc'] + 0.02)),
        D.UniformRadius(distance=(100., 1500.)), D.SinAngle(inclination=None),
        D.UniformAngle(coa_phase=None))
    static = {k: v for k, v in cfg.inj.items() if k not in names}
    cuda = cfg.cuda_model(variable_params=tuple(names), static_params=static, prior=prior)
    orc = OracleGaussianNoise(names, cfg, static_params=static, prior=prior, nthreads=8)
    chains = []
    for m in (cuda, orc):
        np.random.seed(20183)
        s = EmceeEnsembleSampler(m, 200, model_call=models.CallModel(m, 'logplr'))
        s.set_p0()
        s.run(400)
        chains.append(s.chain[:, 200::20, :].reshape(-1, len(names)))
    for i, n in enumerate(names):
        p = stats.ks_2samp(chains[0][:, i], chains[1][:, i]).pvalue
        assert p > 0.01, (n, p)
    # the two runs share a seed, so they should in fact follow each other closely
    assert np.mean(np.all(np.isclose(chains[0], chains[1], rtol=1e-6, atol=1e-9), axis=1)) > 0.5
    # and the sampler finds the injection
    assert abs(np.median(chains[0][:, 0]) - cfg.inj['tc']) < 2e-3


@pytest.mark.slow
def test_posterior_consistency_ks_independent_runs():
    """The north-star criterion as SURVEY.md 8d states it: C1, emcee semantics, 200 walkers,
    >= 2000 iterations after burn-in, thinned by the measured autocorrelation length; the run
    with the CUDA likelihood and the run driven by the CPU oracle use DIFFERENT seeds (independent
    chains); two-sample KS p > 0.01 per parameter.  The device-resident stepper
    (CudaEnsembleSampler, its own Philox stream) is held to the oracle-driven run the same way.

    Walkers start in a small ball around the injection (the reference's ``[initial]`` distribution,
    ``set_p0(prior=...)``) and sample tc, distance and coa_phase: started from prior draws, or with
    the inclination sampled, this posterior keeps walkers in side lobes of tc / in the face-off mode
    for good, and two independent oracle-driven runs already differ by more than KS allows
    (ACL > 600 iterations, p = 0).  Thinning is by twice the ACL: the walkers of one ensemble are
    not independent of each other at equal iteration."""
    from gwin_b200.sampler import CudaEnsembleSampler
    cfg = util.c1()
    names = ['tc', 'distance', 'coa_phase']
    prior = D.JointDistribution(
        names, D.Uniform(tc=(cfg.inj['tc'] - 0.02, cfg.inj['tc'] + 0.02)),
        D.UniformRadius(distance=(100., 1500.)), D.UniformAngle(coa_phase=None))
    init = D.JointDistribution(
        names, D.Uniform(tc=(cfg.inj['tc'] - 2e-4, cfg.inj['tc'] + 2e-4)),
        D.Uniform(distance=(250., 350.)), D.Uniform(coa_phase=(1.0, 1.2)))
    static = {k: v for k, v in cfg.inj.items() if k not in names}
    cuda = cfg.cuda_model(variable_params=tuple(names), static_params=static, prior=prior)
    orc = OracleGaussianNoise(names, cfg, static_params=static, prior=prior, nthreads=8)
    nwalkers, burn, keep = 200, 500, 2500

    def host_run(model, seed):
        np.random.seed(seed)
        s = EmceeEnsembleSampler(model, nwalkers, model_call=models.CallModel(model, 'logplr'))
        s.set_p0(prior=init)
        s.run(burn + keep)
        return s.chain[:, burn:, :]

    def device_run(seed):
        np.random.seed(seed)
        s = CudaEnsembleSampler(cuda, nwalkers, seed=seed)
        s.set_p0(prior=init)
        s.run(burn + keep)
        return s.chain[:, burn:, :]

    chains = {"oracle": host_run(orc, 202), "cuda, host stepper": host_run(cuda, 303),
              "cuda, device stepper": device_run(404)}
    thinned = {}
    for what, c in chains.items():
        acl = int(np.ceil(max(util.autocorrelation_length(c[:, :, i]) for i in range(len(names)))))
        assert acl <= keep // 25, (what, acl)          # >= 25 autocorrelation lengths per walker
        thinned[what] = c[:, ::2 * acl, :].reshape(-1, len(names))
        assert len(thinned[what]) >= 3000
    for what in ("cuda, host stepper", "cuda, device stepper"):
        for i, n in enumerate(names):
            p = stats.ks_2samp(thinned[what][:, i], thinned["oracle"][:, i]).pvalue
            assert p > 0.01, (what, n, p)
        # independent runs: no sample in common
        assert not np.any(np.isin(thinned[what][:50, 0], thinned["oracle"][:, 0]))
    assert abs(np.median(thinned["cuda, host stepper"][:, 0]) - cfg.inj['tc']) < 2e-3


def test_pt_sampler_runs_on_gpu(cfg):
    prior = D.JointDistribution(['tc', 'distance'],
                                D.Uniform(tc=(cfg.inj['tc'] - 0.05, cfg.inj['tc'] + 0.05)),
                                D.UniformRadius(distance=(50., 1000.)))
    static = {k: v for k, v in cfg.inj.items() if k not in ('tc', 'distance')}
    model = cfg.cuda_model(cls=models.MarginalizedPhaseGaussianNoise, variable_params=('tc', 'distance'),
                           static_params=static, prior=prior)
    np.random.seed(3)
    pt = EmceePTSampler(model, 4, 64)
    pt.set_p0()
    pt.run(20)
    assert pt.chain.shape == (4, 64, 20, 2)
    assert np.all(np.isfinite(pt.lnpost))
    ms = pt.model_stats
    assert np.all(np.isfinite(ms['loglr']))


def test_calibrated_model_and_generator(cfg):
    """recalib= on the generator (reference base_data.py:225-230): model protocol, batch
    protocol and generate() against the oracle's CubicSpline (calibration.py:104-173)."""
    import gwin_oracle as O
    from gwin_b200 import calibration
    n = 5
    rec = {d: calibration.CubicSpline(20., 1024., n, d.lower()) for d in cfg.dets}
    orec = {d: O.CubicSpline(20., 1024., n, d.lower()) for d in cfg.dets}
    names = [nm for d in cfg.dets for nm in rec[d].param_names]
    assert names[0] == "recalib_amplitude_h1_0" and names[n] == "recalib_phase_h1_0"
    variable = tuple(PARAM_ORDER) + tuple(names)
    gen = FDomainDetFrameGenerator(FDomainCBCGenerator, cfg.epoch, detectors=cfg.dets,
                                   variable_args=variable, delta_f=cfg.delta_f, f_lower=cfg.f_lower,
                                   recalib=rec)
    data = {d: FrequencySeries(cfg.data[d], cfg.delta_f, cfg.epoch) for d in cfg.dets}
    model = models.GaussianNoise(variable, data, gen, cfg.f_lower, psds=cfg.psds())
    orc = cfg.oracle(recalibration=orec)
    rng = np.random.default_rng(12)
    p = dict(cfg.inj)
    p.update(zip(names, rng.normal(0, 0.05, len(names))))
    model.update(**p)
    lr0, hd0, hh0 = orc.loglr(**p)
    assert abs(model.loglr - lr0) <= tol(lr0)
    # generate(): the calibrated detector-frame waveform
    wf = gen.generate(**p)
    ref = orc.generator.generate(**p)
    for det in cfg.dets:
        a, b = wf[det].numpy(), ref[det]
        assert len(a) == len(b)
        assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(b))
    # batch protocol
    W = 6
    X = np.column_stack([util.draw_bbh(rng, W, cfg), rng.normal(0, 0.05, (W, len(names)))])
    res = models.CallModel(model, "loglr", return_all_stats=False).batch(X)
    ref = [orc.loglr(**dict(zip(variable, row)))[0] for row in X]
    assert np.all(np.abs(res.values - ref) <= tol(np.array(ref)))
    # a missing calibration parameter is a KeyError (calibration.py:151-161)
    q = dict(cfg.inj)
    gen2 = FDomainDetFrameGenerator(FDomainCBCGenerator, cfg.epoch, detectors=cfg.dets,
                                    variable_args=PARAM_ORDER, delta_f=cfg.delta_f, f_lower=cfg.f_lower,
                                    recalib=rec)
    with pytest.raises(KeyError):
        gen2.generate(**q)
    with pytest.raises(ValueError):        # outside the regime the spline fit is a cubic in
        calibration.values_matrix(rec, cfg.dets, dict(zip(names, np.full(len(names), 1.5))), 1)


def test_batch_async_two_calls_in_flight(cfg):
    """CallModel.batch_async (gwin_loglr_batch_points_submit / _wait): queued evaluations give
    the numbers of the synchronous call bit for bit, two may be pending, a third is refused."""
    import gwin_b200
    names = ['tc', 'distance', 'inclination', 'mass1']
    prior = D.JointDistribution(
        names, D.Uniform(tc=(cfg.inj['tc'] - 0.05, cfg.inj['tc'] + 0.05)),
        D.UniformRadius(distance=(100., 1000.)), D.SinAngle(inclination=None), D.Uniform(mass1=(20., 40.)))
    static = {k: v for k, v in cfg.inj.items() if k not in names}
    model = cfg.cuda_model(variable_params=tuple(names), static_params=static, prior=prior)
    rng = np.random.default_rng(11)
    A = np.column_stack([cfg.inj['tc'] + rng.uniform(-0.06, 0.06, 300), rng.uniform(100., 1000., 300),
                         np.arccos(rng.uniform(-1, 1, 300)), rng.uniform(20., 40., 300)])
    B = A[::-1][:200].copy()
    Ap = gwin_b200.pinned_empty(A.shape)
    Ap[:] = A
    for stat, all_stats in (('logposterior', True), ('loglr', False), ('logplr', True)):
        call = models.CallModel(model, stat, return_all_stats=all_stats)
        ra, rb = call.batch(A), call.batch(B)
        pa = call.batch_async(Ap)                 # page-locked points: copied from directly
        pb = call.batch_async(B)                  # pageable: through the slot's staging buffer
        with pytest.raises(ValueError):
            call.batch_async(A)                   # both slots busy (GWIN_EINVAL)
        qb, qa = pb.result(), pa.result()         # any order
        np.testing.assert_array_equal(qa.values, ra.values)
        np.testing.assert_array_equal(qb.values, rb.values)
        if all_stats:
            for k in ra.stats:
                np.testing.assert_array_equal(qa.stats[k], ra.stats[k])
        assert np.isneginf(ra.values).any() or stat == 'loglr'      # some points outside the prior (skipped)
    # a longer pipeline: results stay matched to their submissions
    call = models.CallModel(model, 'loglr', return_all_stats=False)
    batches = [A[i:i + 64] for i in range(0, 256, 64)]
    ref = [call.batch(b).values for b in batches]
    pend, got = None, []
    for b in batches:
        nxt = call.batch_async(b)
        if pend is not None:
            got.append(pend.result().values)
        pend = nxt
    got.append(pend.result().values)
    for g, r in zip(got, ref):
        np.testing.assert_array_equal(g, r)
    # a handle dropped without result(): its slot is released
    call.batch_async(A)
    call.batch_async(B)
    import gc
    gc.collect()
    np.testing.assert_array_equal(call.batch_async(A).result().values, call.batch(A).values)
