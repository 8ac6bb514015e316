"""quantrl_b200 — B200-native hot path of QuantRL behind the reference's plugin surface (backend ``'b200'``).

Importing the package never compiles anything; the C-ABI library must have been built in-tree
(``python -m quantrl_b200.build``).  Every compute entry point raises if it is missing — there is no fallback.
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401
from .build import build_extension  # noqa: F401


def __getattr__(name):
    # torch-dependent parts are imported lazily so that `import quantrl_b200` stays cheap for build()/tests
    if name in ("B200Backend", "get_backend_instance"):
        from . import backends
        return getattr(backends, name)
    if name in ("BaseVecEnv", "LinearVecEnv", "LinearizedHOVecEnv"):
        from . import envs
        return getattr(envs, name)
    if name in ("B200IVPSolver", "get_IVP_solver"):
        from . import solvers
        return getattr(solvers, name)
    if name == "register":
        from .plugin import register
        return register
    raise AttributeError(name)
