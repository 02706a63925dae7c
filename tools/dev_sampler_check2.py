import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from scipy import stats
import util
from gwin_b200 import distributions as D, models
from gwin_b200.sampler import CudaEnsembleSampler, EmceeEnsembleSampler
inj = dict(util.BBH_INJ); inj['distance'] = 1e12
cfg = util.Config("noise", 8, 2048., 20., ["H1", "L1", "V1"], inj)
for names, prior in (
    (['distance', 'tc'], D.JointDistribution(['distance', 'tc'], D.UniformRadius(distance=(1e7, 2e7)), D.Uniform(tc=(6.05, 6.15)))),
    (['distance', 'inclination', 'coa_phase', 'ra', 'dec', 'tc'], D.JointDistribution(['distance', 'inclination', 'coa_phase', 'ra', 'dec', 'tc'], D.UniformRadius(distance=(1e7, 2e7)), D.SinAngle(inclination=None), D.UniformAngle(coa_phase=None), D.UniformSky(), D.Uniform(tc=(6.05, 6.15)))),
):
    static = {k: v for k, v in cfg.inj.items() if k not in names}
    model = cfg.cuda_model(variable_params=tuple(names), static_params=static, prior=prior)
    for cls in ('host', 'dev'):
        np.random.seed(3)
        s = EmceeEnsembleSampler(model, 512, model_call=models.CallModel(model, 'logplr')) if cls == 'host' else CudaEnsembleSampler(model, 512, seed=42)
        s.set_p0(); s.run(400)
        d = np.asarray(s.samples['distance'])
        for lo in (0, 100, 300):
            u = (d[:, lo::20].ravel() ** 3 - 1e21) / 7e21
            print(len(names), cls, 'from', lo, 'acc %.3f' % s.acceptance_fraction.mean(), 'KS', '%.3f p=%.2g' % tuple(stats.kstest(u, 'uniform')[:2]), 'hist', np.round(np.histogram(u, bins=5, range=(0, 1))[0] / len(u), 3))
        print('   lnpost range', s.lnpost[:, -1].min(), s.lnpost[:, -1].max())
