"""stan4bart_b200 -- B200-native implementation of stan4bart's BART <-> Stan Gibbs hot path.

Only the hot path lives here: CUDA kernels + C ABI under csrc/,
and the Python mirror of the reference's `.Call` interface in sampler.py.
"""
from .structs import BartConfig, CommonControl, GlmmData, StanControl, StanData, bart_config, stan_control  # noqa: F401

__all__ = ["BartConfig", "CommonControl", "GlmmData", "StanControl", "StanData", "bart_config", "stan_control"]
