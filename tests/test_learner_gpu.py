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


def test_config5_size_ppo_on_8192_envs_single_gpu(native, cuda):
    """BASELINE config 5 at its size on one GPU: PPO on 8192 stochastic envs (n = 4, K = 100, S = 1000), the single-kernel
    step (in-kernel affine drift, fused reward, re-enqueued argument block), rollouts on the device.  The step protocol of
    the reference (quantrl/envs/base.py:1233-1293) is what the learner drives: auto-reset at the episode end, lock-stepped
    dones; the return improves."""
    from quantrl_b200.learner import PPO
    from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices
    A0, Au = affine_matrices(4, 1, seed=0)
    env = AffineLinearVecEnv(n_envs=8192, n=4, A0=A0, Au=3.0 * Au, noise_prefixes=0.05, x0=1.0, t_norm_max=10.005,
                             t_norm_ssz=0.01, action_interval=100, observation_stds=[0.01] * 4, seed=1, data_idxs=[-1],
                             use_cuda_graph=True, fused_coefficients=True, dir_prefix="/tmp/qrl_test")
    ppo = PPO(env, lr=1e-3, seed=0)
    hist = ppo.learn(6)
    assert ppo.n_steps == 10 and env.batch_idx >= 6 and len(env.data_rewards) >= 6
    assert all(torch.isfinite(torch.tensor(h["mean_step_reward"])) for h in hist)
    assert hist[-1]["mean_episode_return"] > hist[0]["mean_episode_return"], hist
    assert env.nonfinite_detected is False
    env.close()
