"""Scratch timing of qrl_ode_step (CUDA-graph replay, pure device time)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import systems as sy
from quantrl_b200 import _native as N

def bench(system, E, m, q, K=100, method="rk4", reps=10, traj=True, atol=1e-9, rtol=1e-6, flags=0, reward_kind=0):
    dev = "cuda"
    d = 2 * m + q * q
    if system == "optomech":
        p = sy.optomech_default_params(E, seed=2); y0 = sy.optomech_y0(p)
    else:
        p = sy.kerr_default_params(E, m, seed=4); y0 = sy.kerr_y0(p, m); y0[:, :m] = 0.3
    P = torch.as_tensor(p, device=dev); Y0 = torch.as_tensor(y0, device=dev)
    u = torch.as_tensor(np.random.default_rng(3).uniform(-1, 1, (E, 1)), device=dev)
    S = torch.empty((K + 1, E, d), dtype=torch.float64, device=dev); O = torch.empty_like(S)
    yl = Y0.clone(); ol = torch.empty_like(Y0)
    hs = torch.zeros((E,), dtype=torch.float64, device=dev); ns = torch.zeros((E, 2), dtype=torch.int32, device=dev)
    mm = torch.tensor([float("inf"), float("-inf")], dtype=torch.float64, device=dev); nan = torch.zeros((1,), dtype=torch.int32, device=dev)
    a = N.OdeArgs()
    a.system = {"optomech": N.SYS_OPTOMECH, "kerr_chain": N.SYS_KERR_CHAIN}[system]
    a.num_modes, a.num_quads, a.n_envs, a.K = m, q, E, K
    a.method = N.ODE_METHODS[method][0]
    a.substeps, a.t0, a.t_ssz, a.atol, a.rtol = 1, 0.0, 0.01, atol, rtol
    a.params, a.params_stride_e, a.actions, a.n_actions = N.ptr(P), P.shape[1], N.ptr(u), 1
    a.y0, a.y_last, a.obs_last = N.ptr(yl), N.ptr(yl), N.ptr(ol)
    if traj:
        a.states, a.observations = N.ptr(S), N.ptr(O)
    a.h_state, a.n_steps, a.obs_minmax, a.nan_flag = N.ptr(hs), N.ptr(ns), N.ptr(mm), N.ptr(nan)
    a.flags = flags
    if reward_kind:
        R = torch.empty((K + 1, E), dtype=torch.float64, device=dev); rl = torch.empty((E,), dtype=torch.float64, device=dev)
        a.reward_kind, a.reward, a.reward_last = reward_kind, N.ptr(R), N.ptr(rl)
    for _ in range(3):
        N.ode_step(a)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        N.ode_step(a)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for i in range(reps):
            a.t0 = i * K * 0.01
            N.ode_step(a)
    g.replay(); torch.cuda.synchronize()
    ns.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sub = E * K / (ms * 1e-3)
    steps = ns.sum().item() / reps if method != "rk4" else E * K
    flops = (4 * (4 * q ** 3 + 40) + 13 * d) if method == "rk4" else 0
    return dict(system=system, E=E, q=q, method=method, traj=traj, flags=flags, reward_kind=reward_kind, ms=round(ms, 4), substeps_per_s=f"{sub:.3e}",
                rk_steps_per_interval=steps / E, gflops=round(sub * flops / 1e9, 1), hbm_GBs=round(sub * 8 * (2 * d + 1) / 1e9, 1) if traj else 0)

if __name__ == "__main__":
    for E in (256, 1024, 2048, 4096, 8192):
        for flags in (1, 2):
            print(json.dumps(bench("optomech", E, 2, 4, flags=flags)))
