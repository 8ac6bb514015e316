"""Development aid: us per ``step(host actions) -> host obs, reward`` at BASELINE config 3 for the flavours of the host
round trip (set through the environment before the env is built):

  default                    lean path: replay of the captured one-kernel graph + stream synchronisation, bookkeeping inline
  QRL_HOST_STEP_NATIVE=1     ONE native call: launch + spin on the completion word (qrl_em_step_host, ABI 5)
  QRL_HOST_STEP_NATIVE=sync  ONE native call: launch + cudaStreamSynchronize
  QRL_NO_FAST_HOST_STEP=1    the general step_async / step_wait path (graph replay inside update())

and, for orientation, the floor of the same launch: enqueue + synchronise from C with nothing else around it.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def loop(env, acts, n):
    for i in range(n):
        env.step(acts[i % 8])


def main():
    steps = int(os.environ.get("STEPS", "90"))
    for n_envs, n in ((4096, 4), (8192, 4), (8192, 8)):
        env = bench.make_env(0, return_numpy=True, host_zero_copy=True, host_output_views=True, n_envs=n_envs, n=n)
        acts = [np.random.default_rng(5).uniform(-1, 1, (n_envs, 1)) for _ in range(8)]
        env.reset()
        loop(env, acts, 8)
        best = 1e9
        for _ in range(5):
            if env.t_idx + (steps + 1) * env.action_interval >= env.shape_T[0]:
                env.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            loop(env, acts, steps)
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / steps * 1e6)
        fast = env.__dict__.get("_fast_host")
        kind = "graph replay" if not fast else ("lean, graph replay" if not fast["native"] else ("native call, spin" if fast["blocks"][0]["done"] is not None else "native call, stream sync"))
        line = f"E={n_envs} n={n}: step() {best:7.2f} us  [{kind}]"
        if fast and fast["native"]:
            # floor: the same argument block enqueued and waited for from C, nothing else on the host
            b = fast["blocks"][env._h_sel]
            a = b["a"]
            t = 0
            fn, ref, st, done = fast["fn"], b["ref"], fast["stream"], b["done"]
            fl = 1e9
            for _ in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(steps):
                    env._fast_seq = a.done_value = env._fast_seq + 1
                    fn(ref, st, done, 0)
                torch.cuda.synchronize()
                fl = min(fl, (time.perf_counter() - t0) / steps * 1e6)
            line += f"   bare native call {fl:7.2f} us"
        else:
            # floor of the graph flavour: replay + stream synchronisation, nothing else on the host
            g = env._graph_host if env._h_sel == 0 else env._graph_host1
            stream = env._replay_stream
            fl = 1e9
            for _ in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(steps):
                    g.replay()
                    stream.synchronize()
                torch.cuda.synchronize()
                fl = min(fl, (time.perf_counter() - t0) / steps * 1e6)
            line += f"   bare replay + sync {fl:7.2f} us"
            import ctypes as C
            from quantrl_b200 import _native as N
            a, keep = env.interval_args(env.action_interval, env._x_last, True, graph_mode="host")
            fn, ref, st = N.lib().qrl_em_step_host, C.byref(a), C.c_void_p(stream.cuda_stream)
            fl = 1e9
            for _ in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(steps):
                    fn(ref, st, None, 0)
                torch.cuda.synchronize()
                fl = min(fl, (time.perf_counter() - t0) / steps * 1e6)
            line += f"   bare launch + sync (same args) {fl:7.2f} us"
        print(line, flush=True)
        env.close()


if __name__ == "__main__":
    main()
