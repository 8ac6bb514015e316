"""CPU, world_size 2 over gloo: the host-side sharding / learner-link logic of the multi-GPU path (the kernels
themselves are covered by the gpu tests; sharding invariance of the Philox stream by
test_em_gpu.py::test_philox_sharding_invariance_full_size)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quantrl_b200.distributed import LearnerLink, all_shard_bounds, shard_bounds


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 4096, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            b = all_shard_bounds(n, w)
            assert b[0][0] == 0 and sum(c for _, c in b) == n
            assert all(b[i][0] + b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in b) - min(c for _, c in b) <= 1
    assert shard_bounds(10, 0, 4) == (0, 3) and shard_bounds(10, 3, 4) == (8, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, E, n_obs, n_act, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        link = LearnerLink(E, n_obs, n_act, device="cpu", dst=0)
        rng = np.random.default_rng(0)
        actions = torch.as_tensor(rng.uniform(-1, 1, (E, n_act)))       # same on every rank, only rank 0's is used
        for step in range(3):
            loc = link.scatter_actions(actions if rank == 0 else None)
            off, cnt = link.offset, link.count
            assert torch.equal(loc, actions[off:off + cnt])
            # fake env: obs = global env index + 0.1 * action sum + step, reward = -global index
            gidx = torch.arange(off, off + cnt, dtype=torch.float64)
            obs = gidx[:, None] + 0.1 * loc.sum(1, keepdim=True) + step + torch.zeros((cnt, n_obs), dtype=torch.float64)
            res = link.gather_step(obs, -gidx, truncated_local=(rank == 1 and step == 2))
            if rank == 0:
                o, r, trunc = res
                exp = torch.arange(E, dtype=torch.float64)[:, None] + 0.1 * actions.sum(1, keepdim=True) + step
                assert torch.allclose(o, exp.expand(E, n_obs)) and torch.equal(r, -torch.arange(E, dtype=torch.float64))
                assert trunc == (step == 2)          # any rank's truncation ends the episode for all (base.py:485-499)
            else:
                assert res is None
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("E", [10, 7])
def test_learner_link_two_ranks_gloo(E):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, E, 4, 1, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, "ok"), (1, "ok")], res


def _oracle_shard_worker(rank, world, port, E, n, K, out):
    """CPU model of the sharded stochastic step: every rank integrates ITS block of envs with the oracle, drawing the
    noise of the global env indices (``env_offset``), and the learner gathers the last observations / rewards."""
    import em_oracle as eo
    import systems as sy
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        link = LearnerLink(E, n, 1, device="cpu", dst=0)
        A0, Au = sy.affine_matrices(n, 1, seed=0)
        dt, seed = 0.01, 1234
        actions = torch.as_tensor(np.random.default_rng(5).uniform(-1, 1, (E, 1)))

        def interval(x0, u, env0, count, t_idx):
            z_w, z_o = eo.philox_noise_block(seed, 0, t_idx, K, env0, count, n)
            Y = eo.em_interval_vec(x0, sy.affine_A(A0, Au, u), np.full((count, n), 0.05), np.sqrt(dt) * z_w, dt)
            obs = Y + 0.01 * z_o
            return Y[-1], obs[-1], -np.sum(obs[-1] ** 2, -1)

        x = np.ones((link.count, n))
        x_full = np.ones((E, n))
        for step in range(2):
            u = link.scatter_actions(actions if rank == 0 else None).numpy()
            x, obs, rew = interval(x, u, link.offset, link.count, step * K)
            res = link.gather_step(torch.as_tensor(obs), torch.as_tensor(rew))
            if rank == 0:   # the unsharded batch, bit for bit (results invariant to the number of ranks)
                x_full, obs_full, rew_full = interval(x_full, actions.numpy(), 0, E, step * K)
                assert np.array_equal(res[0].numpy(), obs_full) and np.array_equal(res[1].numpy(), rew_full) and not res[2]
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_oracle_step_equals_the_unsharded_batch_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_oracle_shard_worker, args=(r, 2, port, 9, 4, 6, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, "ok"), (1, "ok")], res


def test_make_sharded_env_slices_per_env_arguments():
    from quantrl_b200.distributed import make_sharded_env

    class FakeEnv:
        def __init__(self, **kw):
            self.kw = kw

    params = np.arange(10 * 3).reshape(10, 3)
    for rank, (off, cnt) in enumerate(all_shard_bounds(10, 4)):
        env = make_sharded_env(FakeEnv, 10, rank=rank, world_size=4, system_params=params, x0=None, shared=np.ones(3),
                               _per_env=("system_params", "x0"))
        assert env.kw["n_envs"] == cnt and env.kw["env_offset"] == off and env.kw["x0"] is None
        assert np.array_equal(env.kw["system_params"], params[off:off + cnt]) and env.kw["shared"].shape == (3,)
        assert "_per_env" not in env.kw
