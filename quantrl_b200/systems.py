"""Shipped system families: ready-made environments whose coefficient hooks are batched torch on the device
(stochastic) or CUDA device functors (deterministic, see ``csrc/systems.cuh``).

The reference ships no concrete system (its hooks raise ``NotImplementedError``, quantrl/envs/stochastic.py:244-284,
quantrl/envs/deterministic.py:748-818); these are the "user subclasses" of BASELINE.json's configs, with the formulas
of ``oracle/systems.py``.
"""
import numpy as np
import torch

from .envs.stochastic import LinearVecEnv


def affine_matrices(n, n_act=1, seed=0):
    """Config 3/4 drift: A0 = J - 0.05 I with J antisymmetric, A_a = 0.1 N(0,1)."""
    rng = np.random.default_rng(seed)
    G = 0.5 * rng.standard_normal((n, n))
    A0 = (G - G.T) / 2.0 * 2.0 ** 0.5 - 0.05 * np.eye(n)
    Au = 0.1 * rng.standard_normal((n_act, n, n))
    return A0, Au


class AffineLinearVecEnv(LinearVecEnv):
    """Stochastic linear envs with action-affine drift ``A(u) = A0 + sum_a u_a A_a`` and constant noise prefixes.

    ``A0`` is ``(n,n)`` (shared) or ``(n_envs,n,n)`` (per-env parameters), ``Au`` is ``(n_actions,n,n)``.
    Reward: ``-sum_c w_c obs_c^2`` evaluated in the kernel epilogue.
    """

    def __init__(self, n_envs, n, A0, Au, noise_prefixes, x0, t_norm_max=100.0, t_norm_ssz=0.01, t_norm_mul=1.0,
                 action_interval=100, action_maximums=None, reward_weights=None, data_idxs=None, name="affine_linear",
                 desc="stochastic linear environments with action-affine drift", **kwargs):
        Au = np.asarray(Au, dtype=np.float64)
        n_act = Au.shape[0]
        kwargs.setdefault("reward_functor", "neg_wsq_obs")
        kwargs.setdefault("cache_all_data", False)
        super().__init__(name=name, desc=desc, params={}, t_norm_max=t_norm_max, t_norm_ssz=t_norm_ssz,
                         t_norm_mul=t_norm_mul, n_envs=n_envs, n_observations=n, n_properties=0, n_actions=n_act,
                         action_maximums=[1.0] * n_act if action_maximums is None else action_maximums,
                         action_interval=action_interval,
                         data_idxs=list(range(1 + n_act + n + 1)) if data_idxs is None else data_idxs,
                         reward_weights=reward_weights, **kwargs)
        dev = self.device
        self._A0 = torch.as_tensor(np.asarray(A0, dtype=np.float64), device=dev)
        self._Au = torch.as_tensor(Au, device=dev).reshape(n_act, n * n).contiguous()
        self._nz = torch.as_tensor(np.broadcast_to(np.asarray(noise_prefixes, dtype=np.float64), (n,)).copy(), device=dev)
        self._x0 = torch.as_tensor(np.broadcast_to(np.asarray(x0, dtype=np.float64), (n_envs, n)).copy(), device=dev)
        self._Au3 = self._Au.view(n_act, n, n)
        self._A = torch.empty((n_envs, n, n), dtype=torch.float64, device=dev)
        self._A0f = self._A0.reshape(-1, n * n).expand(n_envs, n * n) if self._A0.dim() == 2 else self._A0.reshape(n_envs, n * n)

    def reset_states(self):
        return self._x0

    def _affine_drift(self):
        return self._A0, self._Au3

    def get_A(self, t_idx, args):
        n = self.n_observations
        # A0 + sum_a u_a A_a -> (E,n,n): small elementwise torch ops, once per action interval
        u = args[0]
        A = self._A.view(self.n_envs, n * n)
        torch.addcmul(self._A0f, u[:, 0:1], self._Au[0:1], out=A)          # A0 + u_0 A_0 in one launch
        for k in range(1, self.n_actions):
            A.addcmul_(u[:, k:k + 1], self._Au[k:k + 1])
        return self._A

    def get_noise_prefixes(self, t_idx, args):
        return self._nz
