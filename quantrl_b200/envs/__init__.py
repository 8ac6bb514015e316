from .base import BaseVecEnv, PartCacheWriter  # noqa: F401
from .stochastic import LinearVecEnv  # noqa: F401
from .deterministic import LinearizedHOVecEnv  # noqa: F401
