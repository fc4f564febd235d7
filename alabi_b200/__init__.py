"""alabi_b200 — B200-native GP surrogate hot path behind alabi's API.

The package mirrors the reference's module layout for the hot path only
(``kernels`` + ``gp`` = george, ``ensemble`` = emcee, ``core`` = SurrogateModel,
``utility`` / ``gp_utils`` / ``mcmc_utils`` / ``benchmarks`` = alabi helpers); all
arithmetic lives in ``libalabi_b200.so`` (hand-written sm_100a CUDA, see
``csrc/``) reached through ctypes (``_lib``).
"""
from . import _lib, kernels
from .gp import GP, LinAlgError
from .ensemble import EnsembleSampler, SurrogateLogProb
from .core import SurrogateModel, CachedSurrogateLikelihood
from . import utility, gp_utils, mcmc_utils, benchmarks, parallel, nested, cache_utils
from .cache_utils import load_model_cache

__version__ = "0.1.0"
__all__ = ["GP", "kernels", "LinAlgError", "EnsembleSampler", "SurrogateLogProb", "SurrogateModel",
           "CachedSurrogateLikelihood", "utility", "gp_utils", "mcmc_utils", "benchmarks", "parallel", "nested",
           "cache_utils", "load_model_cache"]
