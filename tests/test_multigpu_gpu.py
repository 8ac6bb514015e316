"""GPU, needs >= 2 devices (skipped on a single-GPU box): environment sharding is invisible in the results and the
peer-store gather (obs|reward written by the kernel epilogue into the learner's buffer over NVLink) equals the NCCL
gather.  One process per GPU — never two ranks on one device (device-side barriers would wait on one another)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (one rank per device)")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from quantrl_b200.distributed import LearnerLink, PeerGather, make_sharded_env
    from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        E, n, K = 600, 4, 10
        A0, Au = affine_matrices(n, 1, seed=0)
        kw = dict(n=n, A0=A0, Au=Au, noise_prefixes=0.05, x0=1.0, t_norm_max=0.505, t_norm_ssz=0.01, action_interval=K,
                  observation_stds=[0.01] * n, seed=4, dir_prefix="/tmp/qrl_test")
        env = make_sharded_env(AffineLinearVecEnv, E, **kw)
        peer = PeerGather(E, n, dev, dst=0)
        peer.attach(env)
        link = LearnerLink(E, n, 1, dev, dst=0)
        env.reset()
        actions = torch.as_tensor(np.random.default_rng(0).uniform(-1, 1, (E, 1)), device=dev)
        full = AffineLinearVecEnv(n_envs=E, **kw) if rank == 0 else None   # the unsharded twin on the learner
        if full is not None:
            full.reset()
        # the same shard once more with the step captured in a CUDA graph: the cross-GPU barrier is then fused into the
        # kernel's last block (qrl_em_args.peer_flags) and peer.fence() launches nothing
        env_g = make_sharded_env(AffineLinearVecEnv, E, use_cuda_graph=True, fused_coefficients=True, **kw)
        peer_g = PeerGather(E, n, dev, dst=0)
        peer_g.attach(env_g)
        env_g.reset()
        for step in range(4):
            a_loc = link.scatter_actions(actions if rank == 0 else None)
            env.step_device(a_loc)
            peer.fence()
            env_g.step_device(a_loc)
            assert env_g.last_step_peer_fenced
            peer_g.fence()
            res = link.gather_step(env.Observations[-1], env.Reward[-1])
            if rank == 0:
                o_full, r_full, _ = full.step_device(actions)
                torch.cuda.synchronize()
                assert torch.equal(peer.obs_all, res[0]) and torch.equal(peer.reward_all, res[1])
                assert torch.equal(peer.obs_all, o_full) and torch.equal(peer.reward_all, r_full)
                torch.testing.assert_close(peer_g.obs_all, o_full, rtol=1e-12, atol=1e-14)   # in-kernel affine drift
                torch.testing.assert_close(peer_g.reward_all, r_full, rtol=1e-12, atol=1e-14)
            assert int(env_g._nan[0].item()) == 0      # no barrier time-out
        torch.cuda.synchronize()
        assert peer_g.flags.tolist()[:world] == [4] * world and int(peer_g.epoch.item()) == 4
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_envs_and_peer_gather_match_single_gpu(native):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, "ok"), (1, "ok")], res
