"""Helpers for the GPU parity tests: thin wrappers that call the C ABI on torch CUDA tensors."""
import numpy as np
import torch

from quantrl_b200 import _native as N


def dev(x, device="cuda"):
    if isinstance(x, torch.Tensor):
        return x.to(device).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), device=device)


def em_step(x0, A, nz, t_ssz, K, W=None, obs_noise=None, philox=None, obs_std=None, reward_kind=N.REWARD_NONE,
            reward_w=None, rewards_cum=None, traj=True, flags=0, pack=None):
    """Run qrl_em_step once.  x0 (E,n) numpy/torch; A (n,n)|(E,n,n)|(K,E,n,n); nz (n,)|(E,n)|(K,E,n).
    philox = dict(seed, episode, t_idx, env_offset).  Returns dict of torch tensors."""
    x0 = dev(x0).double()
    E, n = x0.shape
    A, nz = dev(A).double(), dev(nz).double()
    a = N.EmArgs()
    a.n, a.n_envs, a.K, a.t_ssz = n, E, K, t_ssz
    a.A, a.nz = N.ptr(A), N.ptr(nz)
    a.A_stride_k = E * n * n if A.dim() == 4 else 0
    a.A_stride_e = n * n if A.dim() >= 3 else 0
    a.nz_stride_k = E * n if nz.dim() == 3 else 0
    a.nz_stride_e = n if nz.dim() >= 2 else 0
    keep = [x0, A, nz]
    if philox is None:
        a.rng_mode = N.RNG_FED
        if W is not None:
            W = dev(W).double(); keep.append(W); a.W = N.ptr(W)
        if obs_noise is not None:
            obs_noise = dev(obs_noise).double(); keep.append(obs_noise); a.obs_noise = N.ptr(obs_noise)
    else:
        a.rng_mode = N.RNG_PHILOX
        a.seed, a.episode, a.t_idx, a.env_offset = philox["seed"], philox.get("episode", 0), philox.get("t_idx", 0), \
            philox.get("env_offset", 0)
        if obs_std is not None:
            obs_std = dev(obs_std).double(); keep.append(obs_std)
            a.obs_std, a.obs_std_stride_e = N.ptr(obs_std), (n if obs_std.dim() == 2 else 0)
    out = {
        "states": torch.full((K + 1, E, n), float("nan"), dtype=torch.float64, device="cuda") if traj else None,
        "observations": torch.full((K + 1, E, n), float("nan"), dtype=torch.float64, device="cuda") if traj else None,
        "x_last": torch.empty((E, n), dtype=torch.float64, device="cuda"),
        "obs_last": torch.empty((E, n), dtype=torch.float64, device="cuda"),
        "reward": torch.zeros((K + 1, E), dtype=torch.float64, device="cuda"),
        "reward_last": torch.zeros((E,), dtype=torch.float64, device="cuda"),
        "rewards_cum": torch.zeros((E,), dtype=torch.float64, device="cuda") if rewards_cum is None else dev(rewards_cum).double(),
        "minmax": torch.tensor([float("inf"), float("-inf")], dtype=torch.float64, device="cuda"),
        "nan": torch.zeros((1,), dtype=torch.int32, device="cuda"),
    }
    a.x0 = N.ptr(x0)
    a.states, a.observations = N.ptr(out["states"]), N.ptr(out["observations"])
    a.x_last, a.obs_last = N.ptr(out["x_last"]), N.ptr(out["obs_last"])
    a.reward_kind = reward_kind
    if reward_kind != N.REWARD_NONE:
        reward_w = dev(reward_w).double(); keep.append(reward_w)
        a.reward_w = N.ptr(reward_w)
        a.reward, a.reward_last, a.rewards_cum = N.ptr(out["reward"]), N.ptr(out["reward_last"]), N.ptr(out["rewards_cum"])
    a.obs_minmax, a.nan_flag = N.ptr(out["minmax"]), N.ptr(out["nan"])
    a.flags = flags
    if pack is not None:   # dict(T_step (K+1,), actions (E,na), idxs list, rows_total, row0)
        T_step, actions = dev(pack["T_step"]).double(), dev(pack["actions"]).double().reshape(E, -1).contiguous()
        idxs = torch.as_tensor(np.asarray(pack["idxs"], dtype=np.int32), device="cuda")
        rows_total, row0 = pack.get("rows_total", K + 1), pack.get("row0", 0)
        pitch = pack.get("pitch", len(pack["idxs"]))
        out["packed"] = torch.full((E, rows_total, pitch), float("nan"), dtype=torch.float64, device="cuda")
        keep += [T_step, actions, idxs]
        a.packed = N.ptr(out["packed"]) + 8 * row0 * pitch
        a.packed_stride_e, a.packed_row_pitch = rows_total * pitch, pitch
        a.T_step, a.actions, a.n_actions = N.ptr(T_step), N.ptr(actions), actions.shape[1]
        a.sel_idxs, a.n_sel = N.ptr(idxs), len(pack["idxs"])
    N.em_step(a)
    torch.cuda.synchronize()
    return out


def ode_step(system, m, q, y0, K, t0, t_ssz, params=None, actions=None, A=None, D=None, method="rk4", substeps=1,
             atol=1e-9, rtol=1e-6, obs_noise=None, h_state=None, philox=None, obs_std=None, flags=0, reward_kind=0,
             rewards_cum=None):
    """Run qrl_ode_step once; returns dict of torch tensors."""
    y0 = dev(y0).double()
    E, d = y0.shape
    a = N.OdeArgs()
    a.system = {"optomech": N.SYS_OPTOMECH, "kerr_chain": N.SYS_KERR_CHAIN, "const_AD": N.SYS_CONST_AD}[system]
    a.num_modes, a.num_quads, a.n_envs, a.K = m, q, E, K
    a.method = N.ODE_METHODS[method][0]
    a.substeps, a.t0, a.t_ssz, a.atol, a.rtol = substeps, t0, t_ssz, atol, rtol
    keep = [y0]
    if params is not None:
        params = dev(params).double(); keep.append(params)
        a.params, a.params_stride_e = N.ptr(params), (params.shape[-1] if params.dim() == 2 else 0)
    if actions is not None:
        actions = dev(actions).double().reshape(E, -1).contiguous(); keep.append(actions)
        a.actions, a.n_actions = N.ptr(actions), actions.shape[1]
    if A is not None:
        A, D = dev(A).double(), dev(D).double(); keep += [A, D]
        a.A, a.A_stride_e, a.D, a.D_stride_e = N.ptr(A), (q * q if A.dim() == 3 else 0), N.ptr(D), (q * q if D.dim() == 3 else 0)
    out = {
        "states": torch.full((K + 1, E, d), float("nan"), dtype=torch.float64, device="cuda"),
        "observations": torch.full((K + 1, E, d), float("nan"), dtype=torch.float64, device="cuda"),
        "y_last": torch.empty((E, d), dtype=torch.float64, device="cuda"),
        "obs_last": torch.empty((E, d), dtype=torch.float64, device="cuda"),
        "h_state": torch.zeros((E,), dtype=torch.float64, device="cuda") if h_state is None else dev(h_state).double(),
        "n_steps": torch.zeros((E, 2), dtype=torch.int32, device="cuda"),
        "minmax": torch.tensor([float("inf"), float("-inf")], dtype=torch.float64, device="cuda"),
        "nan": torch.zeros((1,), dtype=torch.int32, device="cuda"),
    }
    if obs_noise is not None:
        obs_noise = dev(obs_noise).double(); keep.append(obs_noise); a.obs_noise = N.ptr(obs_noise)
    if philox is not None:
        obs_std = dev(obs_std).double(); keep.append(obs_std)
        a.obs_std, a.obs_std_stride_e = N.ptr(obs_std), (d if obs_std.dim() == 2 else 0)
        a.seed, a.episode, a.t_idx, a.env_offset = philox["seed"], philox.get("episode", 0), philox.get("t_idx", 0), philox.get("env_offset", 0)
    a.y0, a.states, a.observations = N.ptr(y0), N.ptr(out["states"]), N.ptr(out["observations"])
    a.y_last, a.obs_last = N.ptr(out["y_last"]), N.ptr(out["obs_last"])
    a.h_state, a.n_steps = N.ptr(out["h_state"]), N.ptr(out["n_steps"])
    a.obs_minmax, a.nan_flag = N.ptr(out["minmax"]), N.ptr(out["nan"])
    a.flags = flags
    if reward_kind:
        out["reward"] = torch.full((K + 1, E), float("nan"), dtype=torch.float64, device="cuda")
        out["reward_last"] = torch.zeros((E,), dtype=torch.float64, device="cuda")
        out["rewards_cum"] = torch.zeros((E,), dtype=torch.float64, device="cuda") if rewards_cum is None else dev(rewards_cum).double()
        a.reward_kind, a.reward = reward_kind, N.ptr(out["reward"])
        a.reward_last, a.rewards_cum = N.ptr(out["reward_last"]), N.ptr(out["rewards_cum"])
    N.ode_step(a)
    torch.cuda.synchronize()
    return out


def ode_smoke():
    """One RK4 Lyapunov interval on cuda:0 vs oracle/lyap_oracle.py (used by __graft_entry__.smoke)."""
    import lyap_oracle as lo
    import systems as sy
    E, K, dt = 48, 10, 0.01
    p = sy.optomech_default_params(E, seed=2)
    y0 = sy.optomech_y0(p)
    u = np.random.default_rng(0).uniform(-1, 1, E)
    func = lo.make_func(2, 4, lambda t, mo, uu: sy.optomech_mode_rates(p, t, mo, uu), lambda t, mo, uu: sy.optomech_A(p, t, mo, uu),
                        lambda t, mo, uu: sy.optomech_D(p, t, mo, uu), E)
    ref = lo.rk4_interval(func, np.arange(K + 1) * dt, y0, u)
    out = ode_step("optomech", 2, 4, y0, K, 0.0, dt, params=p, actions=u)
    np.testing.assert_allclose(out["states"].cpu().numpy(), ref, rtol=1e-11, atol=1e-13)
