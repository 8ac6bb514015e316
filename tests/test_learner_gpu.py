"""GPU: the learner glue of config 5 — a minimal PPO drives the VecEnv protocol with device-resident rollouts and
improves the return of the shipped stochastic control task."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("graph", [False, True])
def test_ppo_improves_return_on_affine_linear_env(native, cuda, graph):
    from quantrl_b200.learner import PPO
    from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices
    A0, Au = affine_matrices(4, 1, seed=0)
    env = AffineLinearVecEnv(n_envs=1024, n=4, A0=A0, Au=3.0 * Au, noise_prefixes=0.05, x0=1.0, t_norm_max=4.005,
                             t_norm_ssz=0.01, action_interval=50, observation_stds=[0.01] * 4, seed=1, data_idxs=[-1],
                             use_cuda_graph=graph, dir_prefix="/tmp/qrl_test")
    ppo = PPO(env, lr=1e-3, seed=0)
    hist = ppo.learn(8)
    assert len(hist) == 8 and all(torch.isfinite(torch.tensor(h["mean_step_reward"])) for h in hist)
    assert hist[-1]["mean_episode_return"] > hist[0]["mean_episode_return"] + 0.5, hist
    assert env.batch_idx >= 8          # one episode per rollout (rollouts are aligned with episodes)
    env.close()
