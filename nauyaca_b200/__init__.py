"""nauyaca_b200 -- B200-native batched TTV likelihood engine behind nauyaca's API.

One hot path only: flat planetary parameter vector -> mid-transit ephemerides ->
chi-square -> log-likelihood (reference nauyaca/utils.py:16-447), evaluated for whole
batches on the GPU through the C ABI in include/ttvb.h. No CPU fallback.
"""
from ._lib import (MODE_CUBE, MODE_PHYSICAL, MODE_PHYSICAL_FULL, EVENT_CAP, TTVBError, build)
from .engine import BatchEngine, SpecSystem, system_spec
from . import utils
from .pool import GpuPool, vectorized_chi2
from .dist import ShardedEvaluator, shard_bounds
from . import ptsampler, mcmc, optimizers, walkers
from .walkers import init_walkers

__all__ = ["BatchEngine", "SpecSystem", "system_spec", "utils", "GpuPool", "vectorized_chi2", "ShardedEvaluator", "shard_bounds", "ptsampler", "mcmc", "optimizers", "walkers", "init_walkers", "build", "TTVBError", "MODE_CUBE", "MODE_PHYSICAL",
           "MODE_PHYSICAL_FULL", "EVENT_CAP"]
