#!/usr/bin/env python
"""bench.py — env-sub-steps/s of the fused B200 hot path on BASELINE config 3 (contract: see the task statement).

A "step" is one action interval (K = 100 Euler–Maruyama sub-steps) of E = 4096 stochastic linear environments per
GPU (n = 4, in-kernel Philox noise, measurement-noise observations, fused reward), with the reference's trajectory
semantics kept: the (K+1,E,n) States/Observations blocks, the (K+1,E) Reward block and the packed cache rows
[t, actions, obs, cumulative reward] of BaseSB3Env.update_data, the latter written straight into a device-resident
(E, shape_T, n_data) episode cache (2.3 GB > L2, so consecutive steps never rewrite cached lines).

  value  : whole-job env-sub-steps/s, inputs resident in HBM (`step_device`), CUDA events, max over ranks
  e2e    : the same metric through the public VecEnv API with HOST actions in pinned memory and HOST obs/reward
           out (H2D + D2H + sync inside the timed region, every step)
  --impl reference : the reference algorithm's CPU path (oracle port of LinearEnv, one Python loop per env and
           sub-step exactly like quantrl/envs/stochastic.py + backends/numpy.py:178-186) on all host cores
"""
import argparse
import json
import os
import sys
import threading
import time

# The contract is ONE JSON line on stdout.  Native libraries write banners to fd 1 (e.g. "NCCL version ..." when the
# box exports NCCL_DEBUG=VERSION), so fd 1 is pointed at stderr for the whole run and the JSON line goes to a private
# duplicate of the original stdout.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

E_PER_GPU, N_STATE, K_SUB, S_EPISODE, T_SSZ = 4096, 4, 100, 10000, 0.01
METRIC, UNIT = "env_substeps_per_s", "env-substeps/s"
WORKLOAD = ("BASELINE config 3: 4096 stochastic linear envs per GPU (n=4, A=A0+u*A1), Euler-Maruyama, K=100 sub-steps "
            "per action, S=10000 sub-steps per episode, FP64, Wiener + measurement noise")


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: oracle port of the reference's LinearEnv on host cores
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    """Each worker owns `n_env` single envs (the reference's stochastic env is single-env) and steps them in lock-step
    for `warmup + steps` action intervals; returns the wall time of the timed part."""
    wid, n_env, steps, warmup = job
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import em_oracle as eo
    import systems as sy
    A0, Au = sy.affine_matrices(N_STATE, 1, seed=0)
    nz = np.full(N_STATE, 0.05)
    rng = np.random.default_rng(1000 + wid)
    total = (warmup + steps) * K_SUB
    x = [np.ones(N_STATE) for _ in range(n_env)]
    # the reference pre-draws all increments at reset (stochastic.py:162-170); drawn here outside the timed part
    Ws = [np.sqrt(T_SSZ) * rng.standard_normal((total, N_STATE)) for _ in range(n_env)]
    ON = [0.01 * rng.standard_normal((total + 1, N_STATE)) for _ in range(n_env)]
    w = np.ones(N_STATE)
    t_start = None
    for a in range(warmup + steps):
        if a == warmup:
            t_start = time.perf_counter()
        for e in range(n_env):
            u = rng.uniform(-1, 1, 1)
            get_A = lambda t, args: A0 + args[0][0] * Au[0]  # noqa: E731
            Y = eo.linear_env_interval(x[e], a * K_SUB, K_SUB, get_A, lambda t, args: nz, (u, None, None), Ws[e], T_SSZ)
            O = Y + ON[e][a * K_SUB:(a + 1) * K_SUB + 1]
            _ = -np.sum(w * O * O, axis=-1)
            x[e] = Y[-1]
    return time.perf_counter() - t_start


def run_cpu_port(steps, warmup, budget_s=20.0):
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_env_step_s = K_SUB * 15e-6                      # ~15 us of Python per sub-step per env
    n_env = int(max(1, min(64, budget_s / max(1e-9, (steps + warmup) * per_env_step_s))))
    with mp.get_context("spawn").Pool(cores) as pool:
        times = pool.map(_cpu_worker, [(w, n_env, steps, warmup) for w in range(cores)])
    wall = max(times)
    value = cores * n_env * steps * K_SUB / wall
    sample = (f"{cores} processes x {n_env} single envs x {steps} action intervals x {K_SUB} sub-steps "
              f"(oracle port of LinearEnv: Python loop per env per sub-step)")
    return value, cores, sample, wall


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, cores, sample, wall = run_cpu_port(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": E_PER_GPU, "n": N_STATE, "action_interval": K_SUB},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampler (NVML) during the timed region
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index, period=0.01):
        self.samples, self.reasons, self._stop, self.max_mhz = [], set(), threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None
        self.period = period
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self.thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv is not None:
            self.thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def make_env(rank, return_numpy, use_cuda_graph=True, host_zero_copy=True, host_output_views=False):
    from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices
    A0, Au = affine_matrices(N_STATE, 1, seed=0)
    return AffineLinearVecEnv(
        n_envs=E_PER_GPU, n=N_STATE, A0=A0, Au=Au, noise_prefixes=0.05, x0=1.0, t_norm_max=S_EPISODE * T_SSZ + T_SSZ / 2,
        t_norm_ssz=T_SSZ, action_interval=K_SUB, observation_stds=[0.01] * N_STATE, seed=1, rng="philox",
        env_offset=rank * E_PER_GPU, return_numpy=return_numpy, use_cuda_graph=use_cuda_graph, fused_coefficients=True, data_idxs=[-1],
        host_zero_copy=host_zero_copy, host_output_views=host_output_views, dir_prefix="/tmp/qrl_bench")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl", "none"],
                    help="multi-GPU: obs|reward reach the learner through peer stores fused into the kernel epilogue "
                         "(p2p, default; falls back to nccl if symmetric memory is unavailable), an NCCL all_gather, or "
                         "none: independent shards without the per-step learner exchange (the upper bound of weak scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from quantrl_b200 import _native as N

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: quantrl_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N.lib()

    env = make_env(rank, return_numpy=False)
    E, n, K = env.n_envs, env.n_observations, env.action_interval
    steps_per_episode = env.action_steps
    gen = torch.Generator(device=dev).manual_seed(3 + rank)
    n_act_bufs = 64
    actions_dev = torch.rand((n_act_bufs, E, 1), generator=gen, device=dev, dtype=torch.float64) * 2 - 1
    gathered = [torch.empty((world, E, n + 1), dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
    local_or = torch.empty((E, n + 1), dtype=torch.float64, device=dev)
    peer, gather_mode = None, "none"
    if world > 1:
        gather_mode = args.gather
        if args.gather == "p2p":
            try:
                from quantrl_b200.distributed import PeerGather
                peer = PeerGather(world * E, n, dev, dst=0)
                peer.attach(env)
            except Exception as exc:  # noqa: BLE001
                if rank == 0:
                    print(f"[bench] symmetric memory unavailable ({exc!r}); using the NCCL all_gather", file=sys.stderr)
                peer, gather_mode = None, "nccl"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_device(n_steps):
        """n_steps action intervals, everything resident in HBM; multi-GPU: obs|reward all-gathered every step."""
        pending = []
        for i in range(n_steps):
            if env.t_idx + 1 >= env.shape_T[0] or env.batch_idx < 0:
                env.reset()
            obs, rew, _ = env.step_device(actions_dev[i % n_act_bufs])
            if peer is not None:
                peer.fence()     # the kernel epilogue already stored this rank's slice into the learner's buffer
            elif world > 1 and gather_mode != "none":
                local_or[:, :n].copy_(obs)
                local_or[:, n].copy_(rew)
                pending.append(dist.all_gather_into_tensor(gathered[i % 2].view(-1), local_or.view(-1), async_op=True))
                if len(pending) > 1:
                    pending.pop(0).wait()
        for p in pending:
            p.wait()

    # ---- value: device-resident ----
    run_device(args.warmup)
    barrier()
    launches0, replays0 = N.launch_count(), env.graph_replays
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        t_host0 = time.perf_counter()
        run_device(args.steps)
        host_enqueue_us = (time.perf_counter() - t_host0) * 1e6 / args.steps   # host time to ENQUEUE one step (no sync)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    print(f"[bench] rank {rank}: host enqueue {host_enqueue_us:.1f} us/step, device {ms * 1e3 / args.steps:.1f} us/step",
          file=sys.stderr)
    if peer is not None:
        # not timed: the learner's peer-written buffer must equal an NCCL all_gather of every rank's local last rows
        ref = torch.empty((world, E, n), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(ref.view(-1), env.Observations[-1].contiguous().view(-1))
        torch.cuda.synchronize()
        if rank == 0:
            assert torch.equal(peer.obs_all.view(world, E, n), ref), "peer-stored observations differ from the NCCL gather"
    # kernels of libquantrl_b200.so launched in the timed region: eager launches + graph replays x kernels per graph
    launches = (N.launch_count() - launches0) + (env.graph_replays - replays0) * env.graph_native_launches

    # dominant kernel (the fused integration launch) timed live with CUDA events on its stream: the argument block of
    # one interval of an identical env is enqueued back to back (launch cost < kernel time, so the device never idles;
    # timing single launches of the eager env instead measures a cold, down-clocked GPU between Python steps)
    env_k = make_env(rank, return_numpy=False, use_cuda_graph=False)
    env_k.reset()
    for i in range(3):
        env_k.step_device(actions_dev[i % n_act_bufs])
    env_k._set_actions(actions_dev[3])
    k_args, k_keep = env_k.interval_args(K)
    n_k = 100
    for _ in range(10):
        N.em_step(k_args)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(n_k):
        N.em_step(k_args)
    k1.record()
    torch.cuda.synchronize()
    em_ms = k0.elapsed_time(k1) / n_k
    del env_k
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * E * K * args.steps / (ms_max * 1e-3)

    # ---- e2e: public VecEnv API, host actions in, host obs/reward out, every step ----
    # (multi-GPU: the learner gather below reads the device-side last rows, so the results take the copy-node route)
    env2 = make_env(rank, return_numpy=True, host_zero_copy=(world == 1), host_output_views=True)
    host_actions = [np.random.default_rng(5 + rank).uniform(-1, 1, (E, 1)) for _ in range(8)]
    env2.reset()
    e2e_steps = max(10, min(args.steps, 300))
    for i in range(args.warmup):
        env2.step(host_actions[i % 8])
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        obs_h, rew_h, done, _ = env2.step(host_actions[i % 8])
        if world > 1:  # learner-rank gather of the host-visible result
            local_or[:, :n].copy_(env2._obs_last)
            local_or[:, n].copy_(env2._reward_last)
            dist.all_gather_into_tensor(gathered[0].view(-1), local_or.view(-1))
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)   # wall clock: every step already ends with a host sync
    barrier()
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * E * K * e2e_steps / (float(t.item()) * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        hbm_peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured)") if "hbm_gbs" in peaks else \
            (6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)")
        # States + Observations + Reward rows + the selected cache column (cumulative reward, data_idxs=[-1])
        bytes_per_substep = 8 * (2 * n + 1) + 8
        alg_bytes = bytes_per_substep * E * K
        achieved = alg_bytes / (em_ms * 1e-3) / 1e9
        fp64_peak = N.measure_fp64_peak(8192)
        traffic, traffic_src = None, None
        try:   # DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture (profiles/)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))   # written by tools/summarize_profiles.py
            traffic, traffic_src = tj["traffic_bytes_per_launch"], tj["capture"]
        except Exception:  # noqa: BLE001
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": E, "n": n, "action_interval": K,
                       "rng": "philox4x32-10 in-kernel", "trajectory": "States+Observations+Reward blocks kept on device; data cache column data_idxs=[-1] (cumulative reward) "
                                     "written by the same launch into a device-resident (E, shape_T, 1) episode cache",
                       "l2": "inputs per step are < 1 MB (no reuse to exploit); the 33 MB of trajectory rows per step are rewritten "
                             "every step and stay L2-resident (126 MB L2), the episode cache (655 MB) streams to HBM", "parallelism": f"env-sharded dp{world}",
                       "collective": {"none": "none (independent shards, no learner exchange)" if world > 1 else "none",
                                      "nccl": "all_gather(obs|reward) per step (NCCL, async)",
                                      "p2p": "obs|reward stored into the learner rank's buffer over NVLink by the kernel "
                                             "epilogue; the per-step cross-GPU barrier is fused into the same kernel's last block "
                                             "(no collective or barrier launch)"}[gather_mode]},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": E * 1 * 8, "d2h_bytes_per_step": (E * (n + 1) + 4) * 8,
                    "steps": e2e_steps, "api": "AffineLinearVecEnv.step(host actions) -> host obs, reward",
                    "transport": ("zero-copy: the step kernel reads the actions from and writes obs|reward|min,max to pinned host "
                                  "memory over PCIe inside the launch; step() returns NumPy views of that (double buffered) "
                                  "pinned memory, valid until the end of the next step" if getattr(env2, "_zc", None) else
                                  "H2D / D2H copy nodes from / to pinned host memory around the step kernel")},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "em_tile_kernel<4> (qrl_em_step)", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "algorithmic_bytes_per_substep": bytes_per_substep, "kernel_ms": em_ms,
                         # kernel_ms: launches serialised by the stream (no programmatic dependent launch); in the timed loop
                         # the next kernel's prologue overlaps the previous kernel's tail, so a step can be shorter
                         "kernel_share_of_step": min(1.0, em_ms / (ms_max / args.steps)),
                         "fp64_peak_tflops_measured": fp64_peak,
                         "fp64_algorithmic_frac": (2 * n * n + 5 * n) * E * K / (em_ms * 1e-3) / 1e12 / fp64_peak},
        }
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample, _ = run_cpu_port(steps=20, warmup=2, budget_s=15.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
