"""BASELINE config 2 through the public env: LinearizedHOVecEnv (optomech functor, RK4, 100 sub-steps per action,
E = 1024, FP64), host actions in, host observations / rewards out, every step.  Prints env-sub-steps/s."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import systems as sy
from quantrl_b200.envs.deterministic import LinearizedHOVecEnv


def run(E=1024, K=100, method="rk4", steps=90, reward=None, return_numpy=True):
    p = sy.optomech_default_params(E, seed=2)
    y0 = sy.optomech_y0(p)

    class Env(LinearizedHOVecEnv):
        def reset_states(self):
            return torch.as_tensor(y0, device="cuda")

        def get_reward(self):
            V = self.Observations[..., 4:].reshape(*self.Observations.shape[:-1], 4, 4)
            return -torch.diagonal(V, dim1=-2, dim2=-1).sum(-1)

    S = 100 * K
    env = Env(name="optomech", desc="config 2", params={}, num_modes=2, num_quads=4, t_norm_max=S * 0.01 + 0.005,
              t_norm_ssz=0.01, t_norm_mul=1.0, n_envs=E, n_observations=20, n_properties=0, n_actions=1,
              action_maximums=[1.0], action_interval=K, data_idxs=[-1], cache_all_data=False, dir_prefix="/tmp/qrl_lho",
              ode_method=method, system="optomech", system_params=p, return_numpy=return_numpy,
              **({"reward_functor": reward} if reward else {}))
    env.reset()
    acts = [np.random.default_rng(3 + i).uniform(-1, 1, (E, 1)) for i in range(8)]
    for i in range(5):
        env.step(acts[i % 8])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        obs, rew, done, _ = env.step(acts[i % 8])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    env.close()
    return {"config": f"LinearizedHOVecEnv optomech E={E} K={K} {method} reward={reward or 'torch hook'}",
            "us_per_step": round(dt / steps * 1e6, 1), "env_substeps_per_s": f"{E * K * steps / dt:.3e}"}


if __name__ == "__main__":
    for kw in (dict(), dict(reward="entan_ln_obs"), dict(method="dopri5", reward="entan_ln_obs"), dict(E=65536, reward="entan_ln_obs")):
        try:
            print(json.dumps(run(**kw)))
        except Exception as exc:  # noqa: BLE001
            print("failed", kw, repr(exc))
