"""Where does the multi-GPU step tail go?  Per-rank device time per step of the config-3 loop (torchrun, N ranks):
  none      independent shards, no exchange
  push      copy engines: events + cudaMemcpyAsync of every rank's rows into the learner's block (what bench.py runs)
  publish   every rank's step kernel stores its epoch flag (ring + one NVLink store), nobody gathers, no flow control
  pull      publish + learner-side qrl_peer_pull per step on a side stream + credit flow control"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from quantrl_b200.distributed import PeerGather
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
gen = torch.Generator(device=dev).manual_seed(3 + rank)
acts = torch.rand((64, bench.E_PER_GPU, 1), generator=gen, device=dev, dtype=torch.float64) * 2 - 1

def run(mode):
    env = bench.make_env(rank, return_numpy=False)
    peer = None
    if mode != "none":
        peer = PeerGather(world * env.n_envs, env.n_observations, dev, dst=0, mode="push" if mode == "push" else "pull")
        peer.attach(env)
        if mode == "publish":
            orig = peer.next_launch
            peer.next_launch = lambda: (lambda r: (r[0], r[1], r[2], 0))(orig())
    def loop(n):
        for i in range(n):
            if env.t_idx + 1 >= env.shape_T[0] or env.batch_idx < 0:
                env.reset()
            env.step_device(acts[i % 64])
            if peer is not None and mode in ("pull", "push"):
                peer.gather()
        if peer is not None and mode in ("pull", "push"):
            peer.wait()
    loop(20)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loop(steps); e1.record()
    torch.cuda.synchronize(); dist.barrier()
    us = e0.elapsed_time(e1) * 1e3 / steps
    env.close()
    return us

for mode in ("none", "push", "publish", "pull", "none"):
    us = run(mode)
    t = torch.tensor([us], device=dev, dtype=torch.float64)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        print(f"{mode:10s} us/step per rank: " + " ".join(f"{float(x):.1f}" for x in allt), flush=True)
dist.destroy_process_group()
