"""CPU comparator for BASELINE config 2 (SURVEY.md §8d "CPU baseline beside it"): the UNMODIFIED reference
``LinearizedHOVecEnv`` (NumPy backend + SciPy integrators, whole batch flattened into one ODE with one shared step
controller, quantrl/solvers/numpy.py:87-178) on the optomechanical system at E envs, K sub-steps per action, timed on
ONE core (the reference path is single-threaded).  Needs /root/reference, so it runs in the build container only — it
is not part of bench.py, whose CPU legs must also run on the GPU box.

    OMP_NUM_THREADS=1 python tools/ref_cpu_config2.py [--envs 1024] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402

import ref_shim  # noqa: E402

assert ref_shim.available(), "the reference tree is not present"
ref_shim.load()
import make_golden as mg  # noqa: E402
import systems  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--interval", type=int, default=100)
    ap.add_argument("--methods", default="vode,dopri5,dop853")
    ap.add_argument("--config1", action="store_true",
                    help="BASELINE config 1 instead: ONE reference LinearizedHOEnv, 1000 sub-steps, K = 10, evolve()")
    a = ap.parse_args()
    if a.config1:
        return config1(a)
    E, K, ssz = a.envs, a.interval, 0.01
    S = K * (a.steps + 1)
    p = systems.optomech_default_params(E, seed=2)
    y0 = systems.optomech_y0(p)
    actions = np.random.default_rng(3).uniform(-1, 1, (a.steps + 1, E, 1))
    for method in a.methods.split(","):
        env = mg.make_vec_env("optomech", p, 2, 4, y0, K, S, ssz, method, 1e-9, 1e-6)   # the reference's default tolerances
        env.reset()
        env.step_async(actions[0])
        env.step_wait()                       # warm-up interval
        t0 = time.perf_counter()
        for i in range(1, a.steps + 1):
            env.step_async(actions[i])
            env.step_wait()
        dt = (time.perf_counter() - t0) / a.steps
        print(json.dumps({"config": f"reference LinearizedHOVecEnv numpy/{method} optomech E={E} K={K} atol=1e-9 rtol=1e-6",
                          "cores": 1, "ms_per_action_step": round(dt * 1e3, 2),
                          "env_substeps_per_s": float(f"{E * K / dt:.4g}"), "host": "build container (no GPU)"}), flush=True)


def config1(a):
    """single deterministic optomechanical env (m = 2, q = 4, d = 20), t_norm_max = 10 (S = 1000),
    K = 10, actions = action_maximums (``evolve``), the reference's single-env class on the NumPy/SciPy backend."""
    from quantrl.envs.deterministic import LinearizedHOEnv
    p = systems.optomech_default_params(1, seed=2, jitter=False)
    y0 = systems.optomech_y0(p)
    K, S, ssz = 10, 1000, 0.01
    for method in a.methods.split(","):
        env = mg.make_vec_env("optomech", p, 2, 4, y0, K, S - K, ssz, method, 1e-9, 1e-6, cls=LinearizedHOEnv)
        assert env.shape_T[0] == S + 1, env.shape_T
        t0 = time.perf_counter()
        env.evolve(close=False)
        dt = time.perf_counter() - t0
        print(json.dumps({"config": f"reference LinearizedHOEnv numpy/{method} optomech E=1 S={S} K={K} (evolve, config 1)",
                          "cores": 1, "s_per_episode": round(dt, 3), "env_substeps_per_s": float(f"{S / dt:.4g}"),
                          "host": "build container (no GPU)"}), flush=True)


if __name__ == "__main__":
    main()
