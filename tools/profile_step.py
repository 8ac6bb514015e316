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
    n = 400
    t0 = time.perf_counter()
    for i in range(n):
        env.step(acts[i % 8])
    torch.cuda.synchronize()
    full = (time.perf_counter() - t0) / n * 1e6
    # bare replay + sync of the same captured host graph (advances the device-side time index; episode bookkeeping is
    # irrelevant for timing as long as we stay inside the cache: reset first)
    env.reset()
    g, st = env._graph_host, env._replay_stream
    t0 = time.perf_counter()
    for i in range(90):
        g.replay(); st.synchronize()
    bare = (time.perf_counter() - t0) / 90 * 1e6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.reset()
    e0.record()
    for i in range(90):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    dev = e0.elapsed_time(e1) / 90 * 1e3
    print(f"zero_copy={zc} (active: {bool(env._zc)}): env.step {full:.1f} us | graph.replay()+sync {bare:.1f} us | back-to-back replays (device time) {dev:.1f} us")
    env.close()
