"""Host-side profile of env.step (e2e path) — where do the microseconds outside the GPU go?"""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.argv = ["bench.py"]
import bench
env = bench.make_env(0, return_numpy=True)
E = env.n_envs
acts = [np.random.default_rng(i).uniform(-1, 1, (E, 1)) for i in range(8)]
env.reset()
for i in range(30):
    env.step(acts[i % 8])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(300):
    env.step(acts[i % 8])
torch.cuda.synchronize()
print("us/step", (time.perf_counter() - t0) / 300 * 1e6)
pr = cProfile.Profile(); pr.enable()
for i in range(300):
    env.step(acts[i % 8])
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
