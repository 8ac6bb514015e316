"""Where the host time of ``env.step(host actions)`` goes (cProfile) and how the step splits into host-before-launch,
device and host-after-sync time.  Run on a GPU box: python tools/profile_host_step.py"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench

env = bench.make_env(0, return_numpy=True)
acts = [np.random.default_rng(5).uniform(-1, 1, (bench.E_PER_GPU, 1)) for _ in range(8)]
env.reset()
for i in range(30):
    env.step(acts[i % 8])
n = 60   # stay inside one episode (100 action steps)
t0 = time.perf_counter()
for i in range(n):
    env.step(acts[i % 8])
print(f"step: {(time.perf_counter() - t0) / n * 1e6:.1f} us")
# split: time until the launch returns, the sync, the rest
ta = tb = tc = 0.0
for i in range(n - 30):
    t0 = time.perf_counter(); env.step_async(acts[i % 8]); t1 = time.perf_counter()
    out = env.step_wait(); t2 = time.perf_counter()
    ta += t1 - t0; tb += t2 - t1
print(f"step_async {ta / (n - 30) * 1e6:.1f} us, step_wait {tb / (n - 30) * 1e6:.1f} us")
env.reset()
pr = cProfile.Profile()
pr.enable()
for i in range(90):
    env.step(acts[i % 8])
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
