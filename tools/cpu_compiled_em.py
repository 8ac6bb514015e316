"""Compiled, threaded CPU comparator for BASELINE config 3 (NOT the reference arm: the reference executes a Python loop
per env and sub-step, which is what bench.py's ``cpu_baseline`` / ``--impl reference`` time).  This answers the other
question — what does the same work cost on the host cores when it is written in C: oracle/em_port.c (plain C, OpenMP
over environments, xoshiro256++ + libm Box–Muller for the 2n normals per sub-step, States / Observations / Reward rows
kept like the reference does).

    python tools/cpu_compiled_em.py [--envs 4096] [--n 4] [--interval 100] [--steps 50] [--threads N]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--n", type=int, default=4)
    ap.add_argument("--interval", type=int, default=100)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--lyap", action="store_true", help="BASELINE config 2 instead: optomech modes + Lyapunov RK4 "
                    "(oracle/lyap_port.c), default --envs 1024")
    a = ap.parse_args()
    os.environ["OMP_NUM_THREADS"] = str(a.threads)
    import numpy as np
    import em_port
    import systems as sy
    if a.lyap:
        E = 1024 if a.envs == 4096 else a.envs
        p = sy.optomech_default_params(E, seed=2)
        y = sy.optomech_y0(p)
        u = np.random.default_rng(3).uniform(-1, 1, E)
        for _ in range(2):
            y = em_port.lyap_rk4_interval(p, u, y, a.interval, 0.01)[-1]
        t0 = time.perf_counter()
        for _ in range(a.steps):
            y = em_port.lyap_rk4_interval(p, u, y, a.interval, 0.01)[-1]
        dt = (time.perf_counter() - t0) / a.steps
        print(json.dumps({"config": f"oracle/lyap_port.c (gcc -O3 -march=x86-64-v3, OpenMP) optomech RK4 E={E} K={a.interval} h=t_ssz, States rows kept",
                          "threads": a.threads, "ms_per_action_step": round(dt * 1e3, 3),
                          "env_substeps_per_s": float(f"{E * a.interval / dt:.4g}"), "host": f"{os.cpu_count()} logical cores"}))
        return
    A0, Au = sy.affine_matrices(a.n, 1, seed=0)
    u = np.random.default_rng(1).uniform(-1, 1, (a.envs, 1))
    st = em_port.RngStepper(a.envs, a.n, a.interval, 0.01, sy.affine_A(A0, Au, u), np.full(a.n, 0.05), np.full(a.n, 0.01),
                            np.ones(a.n), seed=1)
    for _ in range(3):
        st.step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        st.step()
    dt = (time.perf_counter() - t0) / a.steps
    print(json.dumps({"config": f"oracle/em_port.c (gcc -O3 -march=x86-64-v3, OpenMP) E={a.envs} n={a.n} K={a.interval}, trajectory on",
                      "threads": a.threads, "ms_per_action_step": round(dt * 1e3, 3),
                      "env_substeps_per_s": float(f"{a.envs * a.interval / dt:.4g}"),
                      "host": f"{os.cpu_count()} logical cores"}))


if __name__ == "__main__":
    main()
