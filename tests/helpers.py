"""Helpers for the GPU parity tests: thin wrappers that call the C ABI on torch CUDA tensors."""
import numpy as np
import torch

from quantrl_b200 import _native as N


def dev(x, device="cuda"):
    return torch.as_tensor(np.ascontiguousarray(x), device=device)


def em_step(x0, A, nz, t_ssz, K, W=None, obs_noise=None, philox=None, obs_std=None, reward_kind=N.REWARD_NONE,
            reward_w=None, rewards_cum=None, traj=True):
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
    N.em_step(a)
    torch.cuda.synchronize()
    return out
