"""Test double: a data model with the *gwin_b200 model API* whose likelihood is
evaluated by the CPU oracle (C restatement).  Lets the host logic -- CallModel,
BatchPool, samplers, walker sharding -- run in the ``-m "not gpu"`` suite and
provides the oracle-driven arm of the posterior-consistency test.  Test
infrastructure only."""
from collections import OrderedDict

import numpy as np

from c_oracle import COracle
from gwin_b200.models.base import BaseModel
from gwin_b200.waveform import param_matrix


class OracleGaussianNoise(BaseModel):
    name = "oracle_gaussian_noise"

    def __init__(self, variable_params, cfg, marg=False, nthreads=4, **kwargs):
        super(OracleGaussianNoise, self).__init__(variable_params, **kwargs)
        self._co = COracle(cfg.data, cfg.delta_f, cfg.epoch, cfg.f_lower, psds=cfg.psds())
        self._dets = list(cfg.dets)
        self._mode = int(marg)
        self._nthreads = nthreads
        self.lognl = self._co.lognl()

    @property
    def _extra_stats(self):
        return ['loglr'] + ['{}_cplx_loglr'.format(d) for d in self._dets] + \
               ['{}_optimal_snrsq'.format(d) for d in self._dets]

    def _kernel(self, params, W):
        P = param_matrix(params, W, static=self._static_params)
        return self._co.batch_loglr(P, mode=self._mode, nthreads=self._nthreads)

    def _loglikelihood(self):
        lr, hd, hh = self._kernel(self.current_params, 1)
        self._current_stats.loglr = float(lr[0])
        for i, d in enumerate(self._dets):
            setattr(self._current_stats, '{}_cplx_loglr'.format(d), complex(hd[0, i] - 0.5 * hh[0, i]))
            setattr(self._current_stats, '{}_optimal_snrsq'.format(d), float(hh[0, i]))
        return float(lr[0]) + self.lognl

    @property
    def loglr(self):
        self.loglikelihood
        return self._current_stats.loglr

    @property
    def logplr(self):
        logp = self.logprior
        if logp == -np.inf:
            return logp
        return logp + self.loglr

    def _batch_callstat(self, callstat, params, W, skip, logp, stats):
        lr, hd, hh = self._kernel(params, W)
        if skip is not None:
            lr = np.where(skip, np.nan, lr)
        logl = lr + self.lognl
        extra = OrderedDict(loglr=lr, loglikelihood=logl)
        for i, d in enumerate(self._dets):
            extra['{}_cplx_loglr'.format(d)] = hd[:, i] - 0.5 * hh[:, i]
            extra['{}_optimal_snrsq'.format(d)] = hh[:, i]
        if skip is not None:
            for k in extra:
                extra[k] = np.where(skip, np.nan, extra[k])
        if callstat == 'loglr':
            return lr, extra
        if callstat == 'loglikelihood':
            return logl, extra
        base = lr if callstat == 'logplr' else logl
        return np.where(skip, -np.inf, logp + np.where(skip, 0., base)), extra
