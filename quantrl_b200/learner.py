"""Minimal on-device PPO for BASELINE config 5 (the learner side of the SB3 ``VecEnv`` surface).

Stable-Baselines3 is what the reference's users drive the envs with (``BaseSB3Env`` implements its ``VecEnv`` protocol,
quantrl/envs/base.py:999-1492; ``SaveOnBestMeanRewardCallback`` in quantrl/utils.py).  SB3 is not installable here, so
this module provides the smallest learner that exercises the same protocol — ``reset()``, ``step(actions)`` returning
``(obs, reward, dones, infos)`` with auto-reset — but keeps observations, rewards and the rollout buffer on the GPU
(``env.step`` with ``return_numpy=False``): the per-step H2D/D2H of the reference (base.py:1245, 1357) disappears.
With ``torch.distributed`` initialised it is the "learner rank" of ``quantrl_b200.distributed``.

This is glue, not the accelerated path: the policy/value MLP and the PPO update are plain torch.
"""
import math

import torch
from torch import nn


class ActorCritic(nn.Module):
    def __init__(self, n_obs, n_act, hidden=64):
        super().__init__()
        self.pi = nn.Sequential(nn.Linear(n_obs, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(), nn.Linear(hidden, n_act))
        self.v = nn.Sequential(nn.Linear(n_obs, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(), nn.Linear(hidden, 1))
        self.log_std = nn.Parameter(torch.full((n_act,), -0.5))

    def dist(self, obs):
        return torch.distributions.Normal(self.pi(obs), self.log_std.exp())

    def value(self, obs):
        return self.v(obs).squeeze(-1)


class PPO:
    """Clipped-surrogate PPO with GAE over a vectorized env that follows the SB3 ``VecEnv`` protocol."""

    def __init__(self, env, n_steps=None, gamma=0.99, lam=0.95, clip=0.2, lr=3e-4, epochs=4, minibatches=4, hidden=64,
                 ent_coef=0.0, vf_coef=0.5, obs_scale=1.0, reward_scale=1.0, seed=0):
        self.env = env
        self.device = env.device
        self.n_envs, self.n_obs, self.n_act = env.n_envs, env.n_observations, env.n_actions
        self.n_steps = n_steps if n_steps is not None else env.action_steps   # rollouts aligned with episodes
        self.gamma, self.lam, self.clip, self.epochs, self.minibatches = gamma, lam, clip, epochs, minibatches
        self.ent_coef, self.vf_coef, self.obs_scale, self.reward_scale = ent_coef, vf_coef, obs_scale, reward_scale
        torch.manual_seed(seed)
        self.ac = ActorCritic(self.n_obs, self.n_act, hidden).to(self.device)
        self.opt = torch.optim.Adam(self.ac.parameters(), lr=lr)
        lo, hi = env.action_space_range
        self.act_lo, self.act_hi = float(lo), float(hi)
        self.obs = None
        self.history = []

    def _prep(self, obs):
        return (obs.to(torch.float32) * self.obs_scale)

    @torch.no_grad()
    def collect(self):
        T, E = self.n_steps, self.n_envs
        f32 = dict(dtype=torch.float32, device=self.device)
        buf = dict(obs=torch.empty((T, E, self.n_obs), **f32), act=torch.empty((T, E, self.n_act), **f32),
                   logp=torch.empty((T, E), **f32), val=torch.empty((T + 1, E), **f32), rew=torch.empty((T, E), **f32),
                   done=torch.empty((T, E), **f32))
        if self.obs is None:
            self.obs = self.env.reset()
        ep_ret = torch.zeros((E,), **f32)
        finished = []
        for t in range(T):
            o = self._prep(self.obs)
            d = self.ac.dist(o)
            a = d.sample()
            buf["obs"][t], buf["act"][t], buf["logp"][t], buf["val"][t] = o, a, d.log_prob(a).sum(-1), self.ac.value(o)
            obs, rew, dones, _ = self.env.step(a.clamp(self.act_lo, self.act_hi).to(torch.float64))
            r = rew.to(torch.float32) * self.reward_scale
            buf["rew"][t] = r
            ep_ret += r
            done = bool(dones[0])      # lock-stepped episodes: all envs finish together (base.py:1293)
            buf["done"][t] = float(done)
            if done:
                finished.append(ep_ret.mean().item())
                ep_ret.zero_()
            self.obs = obs
        buf["val"][T] = self.ac.value(self._prep(self.obs))
        adv = torch.zeros((T, E), **f32)
        last = torch.zeros((E,), **f32)
        for t in reversed(range(T)):
            nonterminal = 1.0 - buf["done"][t]
            delta = buf["rew"][t] + self.gamma * buf["val"][t + 1] * nonterminal - buf["val"][t]
            last = delta + self.gamma * self.lam * nonterminal * last
            adv[t] = last
        buf["adv"], buf["ret"] = adv, adv + buf["val"][:T]
        return buf, finished

    def update(self, buf):
        T, E = self.n_steps, self.n_envs
        flat = {k: buf[k][:T].reshape(T * E, *buf[k].shape[2:]) for k in ("obs", "act", "logp", "adv", "ret")}
        adv = flat["adv"]
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        n = T * E
        stats = {}
        for _ in range(self.epochs):
            perm = torch.randperm(n, device=self.device)
            for idx in perm.chunk(self.minibatches):
                d = self.ac.dist(flat["obs"][idx])
                logp = d.log_prob(flat["act"][idx]).sum(-1)
                ratio = (logp - flat["logp"][idx]).exp()
                a = adv[idx]
                pg = -torch.min(ratio * a, ratio.clamp(1 - self.clip, 1 + self.clip) * a).mean()
                vf = 0.5 * (self.ac.value(flat["obs"][idx]) - flat["ret"][idx]).pow(2).mean()
                ent = d.entropy().sum(-1).mean()
                loss = pg + self.vf_coef * vf - self.ent_coef * ent
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                nn.utils.clip_grad_norm_(self.ac.parameters(), 0.5)
                self.opt.step()
                stats = dict(pg=pg.item(), vf=vf.item(), ent=ent.item())
        return stats

    def learn(self, iterations, log=None):
        for it in range(iterations):
            buf, finished = self.collect()
            stats = self.update(buf)
            mean_ret = sum(finished) / len(finished) if finished else float("nan")
            mean_rew = buf["rew"].mean().item()
            self.history.append(dict(iteration=it, mean_step_reward=mean_rew, mean_episode_return=mean_ret, **stats))
            if log is not None:
                log(self.history[-1])
        return self.history
