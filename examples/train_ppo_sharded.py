"""BASELINE config 5 across the GPUs of one box: environments sharded over the ranks, PPO on the learner rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \\
        examples/train_ppo_sharded.py --envs 8192 --iterations 20

Every rank integrates its block of environments with the fused kernel (no data-path collective); per action step the
learner scatters the actions and gathers ``[obs | reward]`` (``quantrl_b200.distributed.ShardedVecEnv``, NCCL).  The
protocol is covered on the CPU by tests/test_distributed_cpu.py (gloo, world size 2); this script has not been timed on
a multi-GPU box yet (round 2).
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from quantrl_b200.distributed import ShardedVecEnv, make_sharded_env  # noqa: E402
from quantrl_b200.learner import PPO  # noqa: E402
from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192, help="environments in total (sharded over the ranks)")
    ap.add_argument("--iterations", type=int, default=20)
    ap.add_argument("--episode-substeps", type=int, default=1000)
    ap.add_argument("--action-interval", type=int, default=100)
    args = ap.parse_args()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl")
    A0, Au = affine_matrices(4, 1, seed=0)
    local = make_sharded_env(AffineLinearVecEnv, args.envs, n=4, A0=A0, Au=3.0 * Au, noise_prefixes=0.05, x0=1.0,
                             t_norm_max=args.episode_substeps * 0.01 + 0.005, t_norm_ssz=0.01,
                             action_interval=args.action_interval, observation_stds=[0.01] * 4, seed=1, data_idxs=[-1],
                             dir_prefix=f"/tmp/qrl_ppo_rank{dist.get_rank()}")
    env = ShardedVecEnv(local, args.envs)
    if env.is_learner:
        ppo = PPO(env, lr=1e-3, epochs=4, minibatches=4)
        t0 = time.perf_counter()
        ppo.learn(args.iterations, log=lambda h: print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in h.items()}))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        sub = args.iterations * ppo.n_steps * args.envs * args.action_interval
        print(f"{sub / dt:.3e} env-sub-steps/s on {dist.get_world_size()} GPUs including the PPO updates ({dt:.2f} s)")
        env.shutdown()
    else:
        env.serve()
    env.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
