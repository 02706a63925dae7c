import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from scipy import stats
import util
from gwin_b200 import distributions as D, models
from gwin_b200.sampler import CudaEnsembleSampler, EmceeEnsembleSampler, CudaPTSampler, EmceePTSampler
cfg = util.c1()
for names, prior in (
    (['tc', 'distance'], D.JointDistribution(['tc', 'distance'], D.Uniform(tc=(cfg.inj['tc'] - 0.02, cfg.inj['tc'] + 0.02)), D.UniformRadius(distance=(100., 1500.)))),
    (['tc', 'distance', 'inclination', 'coa_phase'], D.JointDistribution(['tc', 'distance', 'inclination', 'coa_phase'], D.Uniform(tc=(cfg.inj['tc'] - 0.02, cfg.inj['tc'] + 0.02)), D.UniformRadius(distance=(100., 1500.)), D.SinAngle(inclination=None), D.UniformAngle(coa_phase=None))),
):
    static = {k: v for k, v in cfg.inj.items() if k not in names}
    model = cfg.cuda_model(variable_params=tuple(names), static_params=static, prior=prior)
    nw, nit = 256, 1500
    np.random.seed(77)
    host = EmceeEnsembleSampler(model, nw, model_call=models.CallModel(model, 'logplr'))
    p0 = host.set_p0(); host.run(nit)
    np.random.seed(77)
    dev = CudaEnsembleSampler(model, nw, seed=1234)
    dev.set_p0(); dev.run(nit)
    print(names, 'acc host %.3f dev %.3f' % (host.acceptance_fraction.mean(), dev.acceptance_fraction.mean()))
    for i, n in enumerate(names):
        for lo in (nit // 4, nit // 2):
            x = host.chain[:, lo::25, i].ravel(); y = dev.chain[:, lo::25, i].ravel()
            print('  %-12s from %4d: host %.6f +- %.6f  dev %.6f +- %.6f  KS p %.3g' % (n, lo, x.mean(), x.std(), y.mean(), y.std(), stats.ks_2samp(x, y).pvalue))
    print('  lnpost max host %.3f dev %.3f' % (host.lnpost[:, -1].max(), dev.lnpost[:, -1].max()))
