"""Scratch timing of qrl_em_step at BASELINE config sizes (CUDA events on the launch stream)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from quantrl_b200 import _native as N

def bench(E, n, K, philox=True, traj=True, reps=50, flags=0, graph=True, pack=0, rotate=False):
    dev = "cuda"
    rng = np.random.default_rng(0)
    A = torch.as_tensor(0.3 * rng.standard_normal((E, n, n)), device=dev)
    nz = torch.full((n,), 0.05, dtype=torch.float64, device=dev)
    std = torch.full((n,), 0.01, dtype=torch.float64, device=dev)
    x = torch.ones((E, n), dtype=torch.float64, device=dev)
    nbuf = reps if rotate else 1
    S = torch.empty((nbuf, K + 1, E, n), dtype=torch.float64, device=dev)
    O = torch.empty((nbuf, K + 1, E, n), dtype=torch.float64, device=dev)
    R = torch.empty((K + 1, E), dtype=torch.float64, device=dev)
    rl = torch.empty((E,), dtype=torch.float64, device=dev); rc = torch.zeros((E,), dtype=torch.float64, device=dev)
    xl = torch.empty((E, n), dtype=torch.float64, device=dev); ol = torch.empty((E, n), dtype=torch.float64, device=dev)
    w = torch.ones((n,), dtype=torch.float64, device=dev)
    mm = torch.tensor([float("inf"), float("-inf")], dtype=torch.float64, device=dev)
    nan = torch.zeros((1,), dtype=torch.int32, device=dev)
    a = N.EmArgs()
    a.n, a.n_envs, a.K, a.t_ssz = n, E, K, 0.01
    a.A, a.A_stride_e, a.nz = N.ptr(A), n * n, N.ptr(nz)
    if philox:
        a.rng_mode, a.seed, a.obs_std = N.RNG_PHILOX, 1, N.ptr(std)
    else:
        W = torch.as_tensor(0.1 * rng.standard_normal((K, E, n)), device=dev)
        ON = torch.as_tensor(0.01 * rng.standard_normal((K + 1, E, n)), device=dev)
        a.rng_mode, a.W, a.obs_noise = N.RNG_FED, N.ptr(W), N.ptr(ON)
    a.x0, a.x_last, a.obs_last = N.ptr(x), N.ptr(xl), N.ptr(ol)
    if traj:
        a.states, a.observations, a.reward = N.ptr(S), N.ptr(O), N.ptr(R)
    a.reward_kind, a.reward_w, a.reward_last, a.rewards_cum = N.REWARD_NEG_WSQ_OBS, N.ptr(w), N.ptr(rl), N.ptr(rc)
    a.obs_minmax, a.nan_flag = N.ptr(mm), N.ptr(nan)
    a.flags = flags
    if pack:
        S1 = reps * K + 1
        cache = torch.zeros((E, S1, n + 4), dtype=torch.float64, device=dev)
        Tt = torch.arange(S1, dtype=torch.float64, device=dev) * 0.01
        acts = torch.zeros((E, 1), dtype=torch.float64, device=dev)
        idxs = torch.arange(n + 3, dtype=torch.int32, device=dev) if pack == 1 else torch.tensor([n + 2], dtype=torch.int32, device=dev)
        nsel = idxs.numel()
        pitch = nsel + (nsel & 1)
        cache = cache[:, :, :pitch].contiguous()
        a.packed_stride_e, a.packed_row_pitch, a.actions, a.n_actions, a.sel_idxs, a.n_sel = S1 * pitch, pitch, N.ptr(acts), 1, N.ptr(idxs), nsel
    for _ in range(5):
        N.em_step(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        # back-to-back launches replayed from a CUDA graph: pure device time, no CPU launch-rate floor
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            N.em_step(a)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            for i in range(reps):
                a.t_idx = i * K
                if pack:
                    a.packed, a.T_step = N.ptr(cache) + 8 * i * K * pitch, N.ptr(Tt) + 8 * i * K
                if rotate and traj:
                    a.states, a.observations = N.ptr(S) + 8 * i * (K + 1) * E * n, N.ptr(O) + 8 * i * (K + 1) * E * n
                N.em_step(a)
        g.replay(); torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    else:
        e0.record()
        for i in range(reps):
            a.t_idx = i * K
            if pack:
                a.packed, a.T_step = N.ptr(cache) + 8 * i * K * pitch, N.ptr(Tt) + 8 * i * K
            N.em_step(a)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sub = E * K / (ms * 1e-3)
    return dict(E=E, n=n, K=K, philox=philox, traj=traj, flags=flags, ms=round(ms, 4), substeps_per_s=f"{sub:.3e}",
                hbm_GBs=round(sub * 8 * (2 * n + 1) / 1e9, 1) if traj else 0)

if __name__ == "__main__":
    print("fp64 peak TFLOP/s:", N.measure_fp64_peak(8192))
    for E, n in ((4096, 4), (8192, 4), (65536, 4), (65536, 8), (262144, 4)):
        for philox in (True, False):
            for traj in (True, False):
                print(json.dumps(bench(E, n, 100, philox, traj)))
