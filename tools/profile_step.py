"""Host-side profile of env.step (e2e path) — where do the microseconds outside the GPU go?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.argv = ["bench.py"]
import bench
for zc in (True, False):
    env = bench.make_env(0, return_numpy=True, host_zero_copy=zc)
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
    print(f"zero_copy={zc} (active: {bool(env._zc)}): us/step {(time.perf_counter() - t0) / 300 * 1e6:.1f}")
    # segments
    seg = {"async": 0.0, "update+replay": 0.0, "sync": 0.0, "rest": 0.0}
    for i in range(300):
        t0 = time.perf_counter(); env.step_async(acts[i % 8]); t1 = time.perf_counter()
        env.update(reset_minmax=False); env.update_data(); t2 = time.perf_counter()
        env._replay_stream.synchronize(); t3 = time.perf_counter()
        env._host_graph_done = False
        host = env._h_out; tr = env.check_truncation(host)
        no = env.n_observations
        r = env._h_out_np[E * no:E * (no + 1)].copy(); o = env._h_out_np[:E * no].reshape(E, no).copy(); t4 = time.perf_counter()
        seg["async"] += t1 - t0; seg["update+replay"] += t2 - t1; seg["sync"] += t3 - t2; seg["rest"] += t4 - t3
    print("  segments us/step:", {k: round(v / 300 * 1e6, 1) for k, v in seg.items()})
    env.close()
