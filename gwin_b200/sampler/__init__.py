"""Sampler wrappers (namespace mirrors ``gwin.sampler``, registry as in
``gwin/sampler/__init__.py:25-30``).  ``kombine`` and the toy ``mcmc`` sampler of
the reference are not re-implemented; a kombine
user plugs ``gwin_b200.pool.BatchPool`` into ``kombine.Sampler(pool=...)``."""
from .cuda import CudaEnsembleSampler, CudaPTSampler
from .emcee import EmceeEnsembleSampler, EmceePTSampler
from .ensemble import EnsembleStepper, PTStepper, default_beta_ladder

samplers = {cls.name: cls for cls in (EmceeEnsembleSampler, EmceePTSampler,
                                      CudaEnsembleSampler, CudaPTSampler)}
