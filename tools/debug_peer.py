"""2-rank smoke of the publish / pull exchange with stage prints (torchrun --nproc-per-node 2 tools/debug_peer.py)."""
import faulthandler, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(60, exit=True)
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from quantrl_b200.distributed import PeerGather, make_sharded_env
from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices
def log(*a): print(f"[r{rank}]", *a, flush=True)
E, n, K = 601, 4, 10
A0, Au = affine_matrices(n, 1, seed=0)
kw = dict(n=n, A0=A0, Au=Au, noise_prefixes=0.05, x0=1.0, t_norm_max=0.505, t_norm_ssz=0.01, action_interval=K, observation_stds=[0.01] * n, seed=4, dir_prefix="/tmp/qrl_test")
mode = sys.argv[1] if len(sys.argv) > 1 else "eager"
ekw = {} if mode == "eager" else dict(use_cuda_graph=True, fused_coefficients=True, return_numpy=(mode == "host"))
env = make_sharded_env(AffineLinearVecEnv, E, **ekw, **kw)
log("env built", mode)
peer = PeerGather(E, n, dev, dst=0)
log("peer built")
peer.attach(env); env.reset()
full = AffineLinearVecEnv(n_envs=E, **kw) if rank == 0 else None
if full: full.reset()
rng = np.random.default_rng(0)
for step in range(8):
    acts = torch.as_tensor(rng.uniform(-1, 1, (E, 1)), device=dev)
    a_loc = acts[peer.offset:peer.offset + peer.count].contiguous()
    if mode == "host":
        env.step(a_loc.cpu().numpy())
    else:
        env.step_device(a_loc)
        if env.t_idx + 1 >= env.shape_T[0]: env.reset()
    log("stepped", step, "epoch", peer.epoch)
    if rank == 0:
        peer.pull(); peer.wait()
        o, r, term = full.step_device(acts)
        torch.cuda.synchronize()
        log("pulled", step, "max diff", float((peer.obs_all - o).abs().max()), float((peer.reward_all - r).abs().max()), "status", int(peer._status.item()))
        if term: full.reset()
torch.cuda.synchronize(); dist.barrier()
log("flags", peer.flags.tolist(), "nan", env._nan.tolist())
peer.check()
dist.destroy_process_group()
log("done")
