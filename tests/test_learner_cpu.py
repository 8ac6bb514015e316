"""CPU: the learner glue (``quantrl_b200.learner.PPO``) on a stand-in env that speaks the VecEnv protocol — the torch
PPO itself has no CUDA dependency; the GPU test drives it with the real fused env (tests/test_learner_gpu.py)."""
import torch

from quantrl_b200.learner import PPO


class _DampEnv:
    """x <- 0.8 x + 0.5 u + 0.02 N(0,1); reward -|x|^2 - 0.01 u^2: the best constant-gain policy pushes x to zero."""

    def __init__(self, n_envs=256, steps=8):
        self.n_envs, self.n_observations, self.n_actions = n_envs, 2, 2
        self.action_steps, self.action_space_range, self.device = steps, [-1.0, 1.0], torch.device("cpu")
        self.gen = torch.Generator().manual_seed(0)
        self.resets = 0

    def reset(self):
        self.x = 1.5 * (torch.rand((self.n_envs, 2), generator=self.gen, dtype=torch.float64) - 0.5) + 1.0
        self.t = 0
        self.resets += 1
        return self.x.clone()

    def step(self, a):
        assert a.dtype == torch.float64 and float(a.abs().max()) <= 1.0        # clamped to the action space
        self.x = 0.8 * self.x + 0.5 * a + 0.02 * torch.randn((self.n_envs, 2), generator=self.gen, dtype=torch.float64)
        rew = -(self.x ** 2).sum(1) - 0.01 * (a ** 2).sum(1)
        self.t += 1
        done = self.t >= self.action_steps
        obs = self.reset() if done else self.x.clone()                            # auto-reset (base.py:1277-1293)
        return obs, rew, [done] * self.n_envs, [{}] * self.n_envs


def test_ppo_improves_the_return_of_a_damping_task():
    env = _DampEnv()
    ppo = PPO(env, lr=3e-3, seed=0, hidden=32)
    hist = ppo.learn(12)
    assert ppo.n_steps == env.action_steps and env.resets == 13              # one episode per rollout + the first reset
    assert all(h["mean_episode_return"] == h["mean_episode_return"] for h in hist)
    assert hist[-1]["mean_episode_return"] > hist[0]["mean_episode_return"] + 1.0, [h["mean_episode_return"] for h in hist]
