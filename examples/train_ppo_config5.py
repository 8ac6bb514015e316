"""BASELINE config 5 (scaled by flags): PPO on vectorized stochastic linear environments, everything on the GPU.

    python examples/train_ppo_config5.py --envs 8192 --iterations 20

The env is the shipped ``AffineLinearVecEnv`` (drift ``A0 + u A1``, reward ``-|obs|^2``): the agent learns the constant
action that damps the oscillator fastest.  Rollouts never leave the device (``env.step`` returns CUDA tensors).
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from quantrl_b200.learner import PPO  # noqa: E402
from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--iterations", type=int, default=20)
    ap.add_argument("--episode-substeps", type=int, default=1000)
    ap.add_argument("--action-interval", type=int, default=100)
    ap.add_argument("--graph", action="store_true", help="replay one CUDA graph per action interval")
    args = ap.parse_args()
    A0, Au = affine_matrices(4, 1, seed=0)
    env = AffineLinearVecEnv(n_envs=args.envs, n=4, A0=A0, Au=3.0 * Au, noise_prefixes=0.05, x0=1.0,
                             t_norm_max=args.episode_substeps * 0.01 + 0.005, t_norm_ssz=0.01,
                             action_interval=args.action_interval, observation_stds=[0.01] * 4, seed=1,
                             data_idxs=[-1], use_cuda_graph=args.graph, dir_prefix="/tmp/qrl_ppo")
    ppo = PPO(env, lr=1e-3, epochs=4, minibatches=4, reward_scale=1.0)
    t0 = time.perf_counter()
    ppo.learn(args.iterations, log=lambda h: print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in h.items()}))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sub = args.iterations * ppo.n_steps * args.envs * args.action_interval
    print(f"{sub / dt:.3e} env-sub-steps/s including the PPO updates ({dt:.2f} s)")
    env.close()


if __name__ == "__main__":
    main()
