"""How much does L2 residency of the trajectory rows matter to the benchmarked step?  (NOT run yet —
written after round 1's GPU budget was spent.)

bench.py's device-resident loop rewrites the same 33 MB of States / Observations / Reward rows every step, so they stay in
the 126 MB L2 (DRAM traffic 0.9 MB per launch).  This tool re-enqueues the argument block of one config-3 interval
(the same block bench.py times for its roofline object) in three ways and prints the kernel time of each:

  in-place   the rows of every launch land on the same addresses (what bench.py measures),
  ring       the rows rotate over R sets of trajectory buffers, R x 33 MB > L2: every launch writes lines that are not in
             L2 and evicts dirty ones to HBM ("working set larger than L2"),
  flush      in-place, but a buffer larger than L2 is overwritten between launches; each launch is timed on its own with
             CUDA events so that the flush stays outside the timed regions.

    python tools/bench_l2_variants.py [--ring 8] [--launches 200]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from quantrl_b200 import _native as N  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ring", type=int, default=8)
    ap.add_argument("--launches", type=int, default=200)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    env = bench.make_env(0, return_numpy=False, use_cuda_graph=False)
    env.reset()
    acts = torch.rand((env.n_envs, 1), device="cuda", dtype=torch.float64) * 2 - 1
    for _ in range(3):
        env.step_device(acts)
    env._set_actions(acts)
    K, E, n = env.action_interval, env.n_envs, env.n_observations
    base, keep = env.interval_args(K)
    # trajectory buffer sets of the ring (the first one is the env's own)
    sets = [(base.states, base.observations, base.reward, None)]
    for _ in range(a.ring - 1):
        st = torch.empty((K + 1, E, n), dtype=torch.float64, device="cuda")
        ob = torch.empty((K + 1, E, n), dtype=torch.float64, device="cuda")
        rw = torch.empty((K + 1, E), dtype=torch.float64, device="cuda")
        sets.append((st.data_ptr(), ob.data_ptr(), rw.data_ptr(), (st, ob, rw)))
    blocks = []
    for s_ptr, o_ptr, r_ptr, _ in sets:
        b = N.EmArgs.from_buffer_copy(base)     # a by-value copy of the argument block
        b.states, b.observations, b.reward = s_ptr, o_ptr, r_ptr
        blocks.append(b)

    def timed(seq):
        for b in seq[:10]:
            N.em_step(b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in seq:
            N.em_step(b)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / len(seq)

    res = {"launches": a.launches, "ring_sets": a.ring, "ring_mb": round(a.ring * (2 * (K + 1) * E * n + (K + 1) * E) * 8 / 2 ** 20, 1)}
    res["in_place_us"] = timed([blocks[0]] * a.launches)
    res["ring_us"] = timed([blocks[i % a.ring] for i in range(a.launches)])
    flush = torch.empty((256 * 2 ** 20 // 8,), dtype=torch.float64, device="cuda")
    per = []
    for i in range(min(a.launches, 100)):
        flush.fill_(float(i))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.em_step(blocks[0])
        e1.record()
        torch.cuda.synchronize()
        per.append(e0.elapsed_time(e1) * 1e3)
    per.sort()
    res["flush_median_us"] = per[len(per) // 2]
    res["note"] = "single launches after a flush also pay the launch latency that back-to-back launches hide"
    print(json.dumps(res))


if __name__ == "__main__":
    main()
