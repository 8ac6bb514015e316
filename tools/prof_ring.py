"""Profile driver (used under ncu): the config-3 interval launch with its trajectory blocks rotating over a ring of 8 sets
(> L2), i.e. every launch writes lines that are not L2 resident.  argv: [n_launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from quantrl_b200 import _native as N
torch.cuda.set_device(0)
env = bench.make_env(0, return_numpy=False, use_cuda_graph=False)
acts = torch.rand((8, env.n_envs, 1), device="cuda", dtype=torch.float64) * 2 - 1
ms = bench.kernel_time(torch, N, env, acts, env.action_interval, n_launch=int(sys.argv[1]) if len(sys.argv) > 1 else 40, ring=8)
print(f"ring of 8: {ms * 1e3:.2f} us per launch")
