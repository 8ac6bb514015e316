"""GPU, needs >= 2 devices (skipped on a single-GPU box): environment sharding is invisible in the results and the
publish / pull gather (every rank keeps its last rows in a ring of its own memory and publishes an epoch flag; the
learner's qrl_peer_pull copies the slices over NVLink) equals the NCCL gather and the unsharded env, bit for bit.
One process per GPU."""
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


def _worker(rank, world, port, out, mode):
    import torch.distributed as dist
    from quantrl_b200.distributed import LearnerLink, PeerGather, make_sharded_env
    from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        E, n, K = 601, 4, 10                                      # uneven shards: 301 + 300
        A0, Au = affine_matrices(n, 1, seed=0)
        kw = dict(n=n, A0=A0, Au=Au, noise_prefixes=0.05, x0=1.0, t_norm_max=0.505, t_norm_ssz=0.01, action_interval=K,
                  observation_stds=[0.01] * n, seed=4, dir_prefix="/tmp/qrl_test")
        link = LearnerLink(E, n, 1, dev, dst=0)
        # three flavours of the same shard, each publishing through its own ring: eager launches (torch hooks), the
        # re-enqueued single-kernel step (device-resident loop), and the host round trip of step()
        env = make_sharded_env(AffineLinearVecEnv, E, **kw)
        env_d = make_sharded_env(AffineLinearVecEnv, E, use_cuda_graph=True, fused_coefficients=True, **kw)
        env_h = make_sharded_env(AffineLinearVecEnv, E, use_cuda_graph=True, fused_coefficients=True, return_numpy=True, **kw)
        # ... and the host round trip whose learner step also waits for every rank's rows of the same step (wait_in_step: in
        # push mode the launches store their last rows straight into the learner's block and only the epoch word is copied)
        env_w = make_sharded_env(AffineLinearVecEnv, E, use_cuda_graph=True, fused_coefficients=True, return_numpy=True, **kw)
        peers = [PeerGather(E, n, dev, dst=0, mode=mode) for _ in range(4)]
        for e_, p_ in zip((env, env_d, env_h), peers):
            p_.attach(e_)
            e_.reset()
        peers[3].attach(env_w, wait_in_step=True)
        env_w.reset()
        full = AffineLinearVecEnv(n_envs=E, **kw) if rank == 0 else None   # the unsharded twin on the learner
        if full is not None:
            full.reset()
        rng = np.random.default_rng(0)
        for step in range(8):                                     # more steps than ring slices, across an episode end
            actions = torch.as_tensor(rng.uniform(-1, 1, (E, 1)), device=dev)
            a_loc = link.scatter_actions(actions if rank == 0 else None)
            o_loc, r_loc, _ = env.step_device(a_loc)
            o_d, r_d, _ = env_d.step_device(a_loc.clone())
            o_h, r_h, done_h, _ = env_h.step(a_loc.cpu().numpy())
            torch.testing.assert_close(o_d, o_loc, rtol=1e-12, atol=1e-14)                      # in-kernel affine drift
            np.testing.assert_allclose(o_h, o_loc.cpu().numpy() if not done_h[0] else o_h, rtol=1e-12, atol=1e-14)
            o_w, r_w, done_w, _ = env_w.step(a_loc.cpu().numpy())      # (the learner's call returns with everybody's rows in place)
            if rank == 0:
                ow_all, rw_all = peers[3].obs_all.clone(), peers[3].reward_all.clone()
            np.testing.assert_array_equal(o_w, o_h)
            res = link.gather_step(env.Observations[-1], env.Reward[-1])
            if rank == 0:
                for p_ in peers[:3]:
                    p_.gather()
                    p_.wait()
                o_full, r_full, term = full.step_device(actions)
                torch.cuda.synchronize()
                assert torch.equal(peers[0].obs_all, res[0]) and torch.equal(peers[0].reward_all, res[1])
                assert torch.equal(peers[0].obs_all, o_full) and torch.equal(peers[0].reward_all, r_full)   # bit for bit
                for p_ in peers[1:3]:
                    torch.testing.assert_close(p_.obs_all, o_full, rtol=1e-12, atol=1e-14)
                    torch.testing.assert_close(p_.reward_all, r_full, rtol=1e-12, atol=1e-14)
                torch.testing.assert_close(ow_all, o_full, rtol=1e-12, atol=1e-14)
                torch.testing.assert_close(rw_all, r_full, rtol=1e-12, atol=1e-14)
                if term:
                    full.reset()
            if env.t_idx + 1 >= env.shape_T[0]:
                env.reset(); env_d.reset()
            assert int(env_d._nan[0].item()) == 0 and int(env._nan[0].item()) == 0      # no flow-control time-out
        for p_ in peers:
            p_.check()
        torch.cuda.synchronize()
        dist.barrier()
        assert peers[1].epoch == 8 and (rank != 0 or peers[1].flags.tolist()[:world] == [8] * world)
        assert mode == "push" or int(peers[1].flags[world].item()) == 8                # pull mode: every rank got its credits
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, repr(e) + traceback.format_exc()[-800:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["push", "pull"])
def test_sharded_envs_and_peer_gather_match_single_gpu(native, mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, mode)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, "ok"), (1, "ok")], res
