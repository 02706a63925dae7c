"""Device-resident ensemble step (``gwin_ens_*`` + ``CudaEnsembleSampler`` / ``CudaPTSampler``)."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

import util
from gwin_b200 import _lib, distributions as D, models, transforms as T
from gwin_b200.sampler import (CudaEnsembleSampler, CudaPTSampler, EmceeEnsembleSampler)
from gwin_b200.sampler.cuda import prior_spec_from_model

pytestmark = pytest.mark.gpu


def _model(cfg, names, prior, **kw):
    static = {k: v for k, v in cfg.inj.items() if k not in names and k not in ('mass1', 'mass2') or
              (k in ('mass1', 'mass2') and 'mchirp' not in names and k not in names)}
    return cfg.cuda_model(variable_params=tuple(names), static_params=static, prior=prior, **kw)


def test_device_prior_matches_host_distributions():
    """gwin_ens_prior == boundary conditions + JointDistribution + parameter assembly on the host."""
    import torch
    cfg = util.c2()
    names = ['mchirp', 'q', 'distance', 'inclination', 'coa_phase', 'ra', 'dec', 'polarization', 'tc']
    prior = D.JointDistribution(
        names, D.Uniform(mchirp=(20., 30.), q=(1., 3.)), D.UniformRadius(distance=(100., 1000.)),
        D.SinAngle(inclination=None), D.UniformAngle(coa_phase=None), D.UniformSky(),
        D.UniformAngle(polarization=(0., np.pi)), D.Uniform(tc=(cfg.inj['tc'] - 0.1, cfg.inj['tc'] + 0.1)))
    model = _model(cfg, names, prior, waveform_transforms=[T.MchirpQToMass1Mass2()])
    spec = prior_spec_from_model(model)
    rng = np.random.RandomState(0)
    W = 4000
    q = np.column_stack([rng.uniform(18, 32, W), rng.uniform(0.8, 3.2, W), rng.uniform(50, 1100, W),
                         rng.uniform(-1, 4, W), rng.uniform(-7, 14, W), rng.uniform(-7, 14, W),
                         rng.uniform(-2.5, 2.5, W), rng.uniform(-4, 7, W),
                         cfg.inj['tc'] + rng.uniform(-0.12, 0.12, W)])
    qd = torch.tensor(q, device="cuda")
    logp = torch.empty(W, dtype=torch.float64, device="cuda")
    skip = torch.empty(W, dtype=torch.uint8, device="cuda")
    params = torch.empty((13, W), dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().gwin_ens_prior(C.byref(spec), W, qd.data_ptr(), logp.data_ptr(),
                                          skip.data_ptr(), params.data_ptr(), None))
    torch.cuda.synchronize()
    res = models.CallModel(model, 'logprior').batch(q)
    host = res.values
    dev = logp.cpu().numpy()
    assert np.array_equal(np.isneginf(host), np.isneginf(dev))
    ok = np.isfinite(host)
    assert ok.sum() > 100 and (~ok).sum() > 100
    assert np.allclose(host[ok], dev[ok], rtol=1e-13, atol=1e-13)
    assert np.array_equal(skip.cpu().numpy().astype(bool), ~ok)
    # the parameter rows equal what the host path hands to the kernel
    from gwin_b200.waveform import param_rows
    hp = param_rows(model._transform_params(**{n: q[:, i] for i, n in enumerate(names)}), W,
                    static=model._static_for_kernel())
    assert np.allclose(params.cpu().numpy()[:, ok], hp[:, ok], rtol=1e-13, atol=1e-13)


def test_device_sampler_recovers_the_prior_and_tracks_the_host_stepper():
    """Noise-only data and sources at 1e7 Mpc: loglr = 0, the posterior is the prior.
    (i) in a bounded 2-D problem the device step (stretch move, accept) reproduces the prior
    exactly (KS against the analytic CDFs); (ii) in 6-D with cyclic / reflected angles, where
    the stretch move itself mixes poorly within a few hundred steps, the device sampler and the
    host stepper (emcee semantics) still produce the same distributions (two-sample KS)."""
    inj = dict(util.BBH_INJ)
    inj['distance'] = 1e12                 # noise-only data
    cfg = util.Config("noise", 8, 2048., 20., ["H1", "L1", "V1"], inj)
    lo, hi = cfg.inj['tc'] - 0.05, cfg.inj['tc'] + 0.05
    names = ['distance', 'tc']
    prior = D.JointDistribution(names, D.UniformRadius(distance=(1e7, 2e7)), D.Uniform(tc=(lo, hi)))
    model = _model(cfg, names, prior)
    np.random.seed(3)
    dev = CudaEnsembleSampler(model, 512, seed=42)
    dev.set_p0()
    dev.run(300)
    s = dev.samples
    sel = (slice(None), slice(50, None, 25))
    u = {'distance': (np.asarray(s['distance'])[sel].ravel() ** 3 - 1e21) / 7e21,
         'tc': (np.asarray(s['tc'])[sel].ravel() - lo) / (hi - lo)}
    for n, v in u.items():
        assert np.all((v >= 0) & (v <= 1)), n
        assert stats.kstest(v, 'uniform').pvalue > 1e-3, (n, stats.kstest(v, 'uniform'))
    assert np.all(np.abs(dev.lnpost[:, -1] - np.asarray(
        models.CallModel(model, 'logprior').batch(dev.chain[:, -1, :]).values)) < 0.05)
    assert 0.5 < dev.acceptance_fraction.mean() < 0.75

    names = ['distance', 'inclination', 'coa_phase', 'ra', 'dec', 'tc']
    prior = D.JointDistribution(
        names, D.UniformRadius(distance=(1e7, 2e7)), D.SinAngle(inclination=None),
        D.UniformAngle(coa_phase=None), D.UniformSky(), D.Uniform(tc=(lo, hi)))
    model = _model(cfg, names, prior)
    np.random.seed(4)
    host = EmceeEnsembleSampler(model, 512, model_call=models.CallModel(model, 'logplr'))
    host.set_p0()
    host.run(300)
    np.random.seed(5)
    dev = CudaEnsembleSampler(model, 512, seed=43)
    dev.set_p0()
    dev.run(300)
    assert abs(dev.acceptance_fraction.mean() - host.acceptance_fraction.mean()) < 0.03
    a, b = host.samples, dev.samples
    for n in names:
        x = np.asarray(a[n])[:, 100::50].ravel()
        y = np.asarray(b[n])[:, 100::50].ravel()
        assert stats.ks_2samp(x, y).pvalue > 1e-3, n
    # boundary conditions were applied when reading the samples
    assert np.all((np.asarray(b['coa_phase']) >= 0) & (np.asarray(b['coa_phase']) < 2 * np.pi))
    assert np.all((np.asarray(b['inclination']) >= 0) & (np.asarray(b['inclination']) <= np.pi))


def test_device_ensemble_sampler_matches_host_sampler_statistically():
    """Same target, two implementations of the emcee step (host numpy RNG vs device Philox) on
    a weak signal (broad, unimodal posterior): per-parameter two-sample KS p > 0.01."""
    inj = dict(util.BBH_INJ)
    inj['distance'] = 2500.
    cfg = util.Config("weak", 8, 2048., 20., ["H1", "L1"], inj)
    names = ['distance', 'inclination']
    prior = D.JointDistribution(names, D.UniformRadius(distance=(500., 8000.)),
                                D.SinAngle(inclination=None))
    model = _model(cfg, names, prior)
    np.random.seed(77)
    host = EmceeEnsembleSampler(model, 256, model_call=models.CallModel(model, 'logplr'))
    host.set_p0()
    host.run(600)
    np.random.seed(78)
    dev = CudaEnsembleSampler(model, 256, seed=1234)
    dev.set_p0()
    dev.run(600)
    assert dev.chain.shape == (256, 600, 2) and dev.lnpost.shape == (256, 600)
    assert abs(dev.acceptance_fraction.mean() - host.acceptance_fraction.mean()) < 0.05
    a = host.samples
    b = dev.samples
    for n in names:
        x = np.asarray(a[n])[:, 200::40].ravel()
        y = np.asarray(b[n])[:, 200::40].ravel()
        assert stats.ks_2samp(x, y).pvalue > 0.01, (n, x.mean(), y.mean())
    # the chain's lnpost is the model's logplr at the stored positions
    pts = dev.chain[:8, -1, :]
    ref = models.CallModel(model, 'logplr').batch(pts).values
    assert np.allclose(dev.lnpost[:8, -1], ref, rtol=1e-10, atol=1e-8)
    # same seed -> same chain (counter-based RNG)
    np.random.seed(78)
    dev2 = CudaEnsembleSampler(model, 256, seed=1234)
    dev2.set_p0()
    dev2.run(20)
    np.testing.assert_array_equal(dev2.chain, dev.chain[:, :20, :])


def test_device_pt_sampler():
    cfg = util.c2()
    names = ['tc', 'distance']
    prior = D.JointDistribution(names, D.Uniform(tc=(cfg.inj['tc'] - 0.05, cfg.inj['tc'] + 0.05)),
                                D.UniformRadius(distance=(50., 2000.)))
    model = _model(cfg, names, prior, cls=models.MarginalizedPhaseGaussianNoise)
    np.random.seed(5)
    pt = CudaPTSampler(model, 4, 64, seed=99)
    assert pt.set_p0().shape == (4, 64, 2)
    pt.run(60)
    assert pt.chain.shape == (4, 64, 60, 2)
    assert np.all(np.isfinite(pt.lnpost))
    ms = pt.model_stats
    # logl history is loglr at the stored positions; lnprob = beta*loglr + logprior
    pts = pt.chain[1, :5, -1, :]
    res = models.CallModel(model, 'logplr').batch(pts)
    assert np.allclose(ms['loglr'][1, :5, -1], res.stats['loglr'], rtol=1e-10, atol=1e-8)
    assert np.allclose(ms['logprior'][1, :5, -1], res.stats['logprior'], rtol=1e-10, atol=1e-8)
    sw = pt.tswap_acceptance_fraction
    assert np.all(sw >= 0) and np.all(sw <= 1) and sw.sum() > 0
    # hotter chains explore more
    assert pt.chain[3, :, 30:, 1].std() > pt.chain[0, :, 30:, 1].std()
    class Unknown(D.Uniform):             # a distribution the device prior kernel does not know
        pass
    with pytest.raises(NotImplementedError):
        CudaPTSampler(_model(cfg, names, D.JointDistribution(names, Unknown(tc=(0., 1.)), D.Uniform(distance=(1, 2)))), 2, 8)


def test_device_sampler_checkpoint_resume(tmp_path):
    """Counter-based RNG: seed + iteration counter + positions resume a run bit for bit."""
    from gwin_b200.io import InferenceFile
    cfg = util.c2()
    names = ['tc', 'distance']
    prior = D.JointDistribution(names, D.Uniform(tc=(cfg.inj['tc'] - 0.05, cfg.inj['tc'] + 0.05)),
                                D.UniformRadius(distance=(50., 2000.)))
    model = _model(cfg, names, prior)
    np.random.seed(8)
    ref = CudaPTSampler(model, 3, 32, seed=7)
    ref.set_p0()
    ref.run(12)
    np.random.seed(8)
    a = CudaPTSampler(model, 3, 32, seed=7)
    a.set_p0()
    a.run(6)
    path = str(tmp_path / "pt.npz")
    with InferenceFile(path, 'w') as fp:
        a.write_results(fp)
    fp = InferenceFile(path, 'r')
    assert fp['samples/tc'].shape == (3, 32, 6) and fp.attrs['sampler'] == 'emcee_pt_cuda'
    b = CudaPTSampler(model, 3, 32, seed=12345)
    b.set_p0(samples_file=fp)
    b.set_state_from_file(fp)
    b.run(6)
    assert b.niterations == 12
    np.testing.assert_array_equal(b.chain, ref.chain[:, :, 6:, :])


@pytest.mark.parametrize("ntemps", [1, 3])
def test_graph_replay_equals_host_launched_iterations(ntemps):
    """One iteration recorded as a CUDA graph and replayed gives the chains of the
    kernel-by-kernel loop bit for bit (the iteration number is a device counter in both),
    also when a run is split into several run() calls."""
    cfg = util.c2()
    names = ['tc', 'distance']
    prior = D.JointDistribution(names, D.Uniform(tc=(cfg.inj['tc'] - 0.05, cfg.inj['tc'] + 0.05)),
                                D.UniformRadius(distance=(50., 2000.)))
    model = _model(cfg, names, prior)

    def make():
        np.random.seed(21)
        s = CudaPTSampler(model, ntemps, 32, seed=5) if ntemps > 1 else CudaEnsembleSampler(model, 32, seed=5)
        s.set_p0()
        return s
    eager = make()
    eager.run(15, use_graph=False)
    graph = make()
    graph.run(9, use_graph=True)
    graph.run(6, use_graph=True)
    np.testing.assert_array_equal(graph.chain, eager.chain)
    np.testing.assert_array_equal(graph.lnpost, eager.lnpost)
    np.testing.assert_array_equal(graph.acceptance_fraction, eager.acceptance_fraction)
    assert graph.likelihood_evals == eager.likelihood_evals
    if ntemps > 1:
        np.testing.assert_array_equal(graph.tswap_acceptance_fraction, eager.tswap_acceptance_fraction)


def test_device_sampler_with_calibration_parameters():
    """A calibrated model (recalib_* parameters with Gaussian priors, gwin/calibration.py:104-173,
    enabled at bin/gwin:199-204) in the device-resident stepper: 2 + 3 detectors x 2 x 4 = 26 sampled
    dimensions; the prior kernel fills the calibration matrix, the likelihood takes the calibrated
    path.  The chain's lnpost is the model's logplr at the stored positions, the host stepper with
    the same model lands on the same distance, a static recalib_* value is honoured."""
    from gwin_b200 import calibration
    from gwin_b200.types import FrequencySeries
    from gwin_b200.waveform import FDomainCBCGenerator, FDomainDetFrameGenerator
    cfg = util.c2()
    n = 4
    rec = {d: calibration.CubicSpline(20., 1024., n, d.lower()) for d in cfg.dets}
    cal_names = [nm for d in cfg.dets for nm in rec[d].param_names]
    fixed_cal = cal_names[-1]                      # one calibration value is static, the others sampled
    names = ['distance', 'tc'] + cal_names[:-1]
    dists = [D.UniformRadius(distance=(100., 1000.)),
             D.Uniform(tc=(cfg.inj['tc'] - 0.01, cfg.inj['tc'] + 0.01))]
    for nm in cal_names[:-1]:
        dists.append(D.Gaussian(**{nm: (-0.3, 0.3), nm + "_mean": 0.0, nm + "_var": 0.03 ** 2}))
    prior = D.JointDistribution(names, *dists)
    static = {k: v for k, v in cfg.inj.items() if k not in names}
    static[fixed_cal] = 0.02
    gen = FDomainDetFrameGenerator(FDomainCBCGenerator, cfg.epoch, detectors=cfg.dets,
                                   variable_args=tuple(names), delta_f=cfg.delta_f, f_lower=cfg.f_lower,
                                   recalib=rec)
    data = {d: FrequencySeries(cfg.data[d], cfg.delta_f, cfg.epoch) for d in cfg.dets}
    model = models.GaussianNoise(tuple(names), data, gen, cfg.f_lower, psds=cfg.psds(),
                                 static_params=static, prior=prior)
    spec = prior_spec_from_model(model)
    assert spec.ndim == 25 and spec.type[2] == 5 and spec.target[2] == 200 and abs(spec.sigma[2] - 0.03) < 1e-15
    np.random.seed(5)
    dev = CudaEnsembleSampler(model, 128, seed=99)
    dev.set_p0()
    dev.run(150)
    assert dev.chain.shape == (128, 150, 25)
    assert 0.05 < dev.acceptance_fraction.mean() < 0.8
    call = models.CallModel(model, 'logplr')
    for it in (0, 75, 149):
        pts = dev.chain[:16, it, :]
        ref = call.batch(pts).values
        assert np.allclose(dev.lnpost[:16, it], ref, rtol=1e-10, atol=1e-7), it
    # the calibration parameters move (they are sampled), within their truncated priors
    s = dev.samples
    x = np.asarray(s[cal_names[0]])
    assert x.std() > 1e-3 and np.all(np.abs(x) <= 0.3)
    # the static value matters: a different one gives a different likelihood at the same points
    static2 = dict(static)
    static2[fixed_cal] = -0.2
    gen2 = FDomainDetFrameGenerator(FDomainCBCGenerator, cfg.epoch, detectors=cfg.dets,
                                    variable_args=tuple(names), delta_f=cfg.delta_f, f_lower=cfg.f_lower,
                                    recalib=rec)
    model2 = models.GaussianNoise(tuple(names), data, gen2, cfg.f_lower, psds=cfg.psds(),
                                  static_params=static2, prior=prior)
    dev2 = CudaEnsembleSampler(model2, 128, seed=99)
    np.random.seed(5)
    dev2.set_p0()
    dev2.run(2)
    assert np.max(np.abs(dev2.lnpost[:, 0] - dev.lnpost[:, 0])) > 1e-3
    pts = dev2.chain[:8, 1, :]
    assert np.allclose(dev2.lnpost[:8, 1], models.CallModel(model2, 'logplr').batch(pts).values, rtol=1e-10, atol=1e-7)
    # host stepper, same model: distance agrees statistically
    np.random.seed(6)
    host = EmceeEnsembleSampler(model, 128, model_call=models.CallModel(model, 'logplr'))
    host.set_p0()
    host.run(150)
    # (150 iterations in 25 dimensions are not a converged run: the two steppers are held to each
    #  other's distance estimate, not to a KS test)
    a = np.asarray(host.samples['distance'])[:, 75::15].ravel()
    b = np.asarray(s['distance'])[:, 75::15].ravel()
    assert abs(np.median(a) - np.median(b)) < 0.05 * np.median(a)
    assert abs(host.acceptance_fraction.mean() - dev.acceptance_fraction.mean()) < 0.1


def test_device_prior_with_sampling_transform_and_jacobian():
    """Sampling transforms on the device (gwin/models/base.py:109-220): the walkers move in
    (mchirp, q), the prior is uniform in (mass1, mass2); the device prior kernel applies the inverse
    transform, the priors of the masses and log |d(mass1, mass2)/d(mchirp, q)| exactly as
    ``_transform_params`` + ``logprior`` (with ``_logjacobian``) do on the host; a short device run
    reproduces the model's logplr at its stored positions and hands back samples in the masses."""
    import torch
    cfg = util.c2()
    variable = ['mass1', 'mass2', 'distance', 'tc']
    prior = D.JointDistribution(
        variable, D.Uniform(mass1=(25., 40.)), D.Uniform(mass2=(20., 35.)),
        D.UniformRadius(distance=(100., 1000.)), D.Uniform(tc=(cfg.inj['tc'] - 0.1, cfg.inj['tc'] + 0.1)))
    st = models.SamplingTransforms(variable, ['mchirp', 'q'], ['mass1', 'mass2'], [T.Mass1Mass2ToMchirpQ()])
    static = {k: v for k, v in cfg.inj.items() if k not in variable}
    model = cfg.cuda_model(variable_params=tuple(variable), static_params=static, prior=prior,
                           sampling_transforms=st)
    assert tuple(model.sampling_params) == ('distance', 'tc', 'mchirp', 'q')
    spec = prior_spec_from_model(model)
    assert spec.derived_masses == 1 and spec.type[2] == _lib.PRIOR_NONE and spec.target[3] == _lib.TARGET_Q
    rng = np.random.RandomState(1)
    W = 4000
    q = np.column_stack([rng.uniform(50, 1100, W), cfg.inj['tc'] + rng.uniform(-0.12, 0.12, W),
                         rng.uniform(18., 34., W), rng.uniform(0.6, 2.2, W)])
    qd = torch.tensor(q, device="cuda")
    logp = torch.empty(W, dtype=torch.float64, device="cuda")
    skip = torch.empty(W, dtype=torch.uint8, device="cuda")
    params = torch.empty((13, W), dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().gwin_ens_prior(C.byref(spec), W, qd.data_ptr(), logp.data_ptr(),
                                          skip.data_ptr(), params.data_ptr(), None))
    torch.cuda.synchronize()
    host = models.CallModel(model, 'logprior').batch(q).values          # logprior includes the log-Jacobian
    dev = logp.cpu().numpy()
    assert np.array_equal(np.isneginf(host), np.isneginf(dev))
    ok = np.isfinite(host)
    assert ok.sum() > 100 and (~ok).sum() > 100
    assert np.allclose(host[ok], dev[ok], rtol=1e-13, atol=1e-13)
    assert np.ptp(host[ok]) > 0.5                                       # the Jacobian varies over the box
    from gwin_b200.waveform import param_rows
    hp = param_rows(model._transform_params(**{n: q[:, i] for i, n in enumerate(model.sampling_params)}), W,
                    static=model._static_for_kernel())
    assert np.allclose(params.cpu().numpy()[:, ok], hp[:, ok], rtol=1e-13, atol=1e-13)
    np.random.seed(8)
    dev_s = CudaEnsembleSampler(model, 64, seed=5)
    dev_s.set_p0()
    dev_s.run(30)
    pts = dev_s.chain[:8, -1, :]
    ref = models.CallModel(model, 'logplr').batch(pts).values
    assert np.allclose(dev_s.lnpost[:8, -1], ref, rtol=1e-10, atol=1e-8)
    s = dev_s.samples
    assert 'mass1' in s.dtype.names and np.all(np.asarray(s['mass1']) >= 25. - 1e-9)
