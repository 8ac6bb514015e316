"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference (read-only at /root/reference).

The reference (Sampreet/quantrl v0.0.8) is pure Python but imports ``gymnasium``, ``stable_baselines3`` and
``matplotlib`` at module import time (quantrl/envs/base.py:15-19, quantrl/plotters.py) and uses the NumPy-1.x
aliases ``np.float_``/``np.complex_`` (quantrl/io.py:99).  None of those packages is installed in this image, so
this module injects inert stand-ins *only for the symbols the reference touches at import/constructor time* and
then imports the reference from ``/root/reference``.  No reference arithmetic is replaced: every number produced
through this shim comes from the reference's own code + NumPy/SciPy.

It is used by ``oracle/make_golden.py`` (run in the build container, where /root/reference exists) to create the
fixtures under ``tests/golden/``, and — through the staged copy ``oracle/_ref/`` that ``oracle/stage_ref.py`` makes
(git-ignored, built like a compiled reference would be, shipped to the GPU box) — by the ``-m gpu`` test that runs the
reference's OWN env class on the ``'b200'`` backend and by ``bench.py --impl reference`` (the reference's own
``LinearEnv`` timed on the host cores).  Nothing in the product package imports it.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("QRL_REFERENCE_ROOT", "/root/reference")
# the staged copy of the reference package (oracle/stage_ref.py -> oracle/_ref/, git-ignored, travels to the GPU box)
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "quantrl"))


def staged_available() -> bool:
    return os.path.isdir(os.path.join(STAGED_ROOT, "quantrl"))


def import_reference(root=None):
    """Import the reference from ``root`` (default: /root/reference if present, else the staged copy oracle/_ref)."""
    global REFERENCE_ROOT
    if root is None:
        root = REFERENCE_ROOT if available() else STAGED_ROOT
    REFERENCE_ROOT = root
    return load()


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load():
    """Import and return the reference ``quantrl`` package (raises if /root/reference is absent)."""
    if not available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    import numpy as np

    if not hasattr(np, "float_"):
        np.float_ = np.float64
    if not hasattr(np, "complex_"):
        np.complex_ = np.complex128

    if "gymnasium" not in sys.modules:
        class Env:  # gymnasium.Env: the reference only inherits from it
            def __init__(self, *a, **k):
                pass

        class Box:
            def __init__(self, low, high, shape=None, dtype=None):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        class MultiDiscrete:
            def __init__(self, nvec):
                self.nvec = nvec

        gym = _stub("gymnasium", Env=Env)
        gym.spaces = _stub("gymnasium.spaces", Box=Box, MultiDiscrete=MultiDiscrete)

    if "stable_baselines3" not in sys.modules:
        class VecEnv:  # SB3 VecEnv: constructor only records the three arguments
            def __init__(self, num_envs, observation_space, action_space):
                self.num_envs = num_envs
                self.observation_space = observation_space
                self.action_space = action_space

        sb3 = _stub("stable_baselines3")
        common = _stub("stable_baselines3.common")
        env_util = _stub("stable_baselines3.common.env_util", is_wrapped=lambda env, cls: False)
        vec_env = _stub("stable_baselines3.common.vec_env", VecEnv=VecEnv)
        callbacks = _stub("stable_baselines3.common.callbacks", BaseCallback=object)
        results_plotter = _stub("stable_baselines3.common.results_plotter", load_results=None, ts2xy=None)
        sb3.common = common
        common.env_util, common.vec_env = env_util, vec_env
        common.callbacks, common.results_plotter = callbacks, results_plotter

    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot", rcParams={})
        mpl.gridspec = _stub("matplotlib.gridspec", GridSpec=object)
        mpl.colors = _stub("matplotlib.colors")
        mpl.ticker = _stub("matplotlib.ticker")

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import quantrl  # noqa: F401
    import quantrl.envs.deterministic  # noqa: F401
    import quantrl.envs.stochastic  # noqa: F401
    import quantrl.solvers.numpy  # noqa: F401
    return quantrl
