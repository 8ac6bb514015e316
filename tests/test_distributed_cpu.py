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

def _run_ranks(target, world, *args, timeout=180):
    """Spawn ``world`` ranks of ``target(rank, world, port, *args, out)``; returns their sorted results.  Ranks that are
    still alive afterwards (stuck in a collective because a peer failed) are terminated, whatever happened."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, *args, out)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        return sorted(out.get(timeout=timeout) for _ in procs)
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.terminate()


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
    res = _run_ranks(_worker, 2, E, 4, 1)
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
    res = _run_ranks(_oracle_shard_worker, 2, 9, 4, 6)
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


class _FakeShard:
    """CPU stand-in for one rank's env shard (the protocol ShardedVecEnv needs: ``reset``, ``step_device``,
    ``truncation_pending``): x <- 0.9 x + 0.1 u + 0.05 * global_index, reward = -|x|^2, 4 action steps per episode,
    truncation when any |x| exceeds ``limit``."""

    def __init__(self, n_envs, env_offset, limit=1e9):
        self.n_envs, self.env_offset, self.limit = n_envs, env_offset, limit
        self.n_observations, self.n_actions, self.action_steps, self.action_space_range = 3, 1, 4, [-1.0, 1.0]
        self.device = torch.device("cpu")
        self.gidx = torch.arange(env_offset, env_offset + n_envs, dtype=torch.float64)[:, None]
        self.resets = 0

    def reset(self):
        self.x = torch.ones((self.n_envs, 3), dtype=torch.float64) * (1.0 + 0.001 * self.gidx)
        self.t = 0
        self.resets += 1
        return self.x.clone()

    def step_device(self, a):
        self.x = 0.9 * self.x + 0.1 * a.reshape(self.n_envs, 1) + 0.05 * self.gidx
        self.t += 1
        return self.x.clone(), -(self.x ** 2).sum(1), self.t >= self.action_steps

    def truncation_pending(self):
        return bool((self.x.abs() > self.limit).any())

    def close(self, **kw):
        pass


def _sharded_env_worker(rank, world, port, E, out):
    from quantrl_b200.distributed import ShardedVecEnv, shard_bounds
    from quantrl_b200.learner import PPO
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) lock-step protocol: the sharded batch == one unsharded env, incl. auto-reset and GLOBAL truncation
        for limit, ends in ((1e9, [3]), (1.25, None)):
            off, cnt = shard_bounds(E, rank, world)
            env = ShardedVecEnv(_FakeShard(cnt, off, limit), E)
            if rank != 0:
                env.serve()
                assert env.env.resets >= 2
                continue
            whole = _FakeShard(E, 0, limit)
            obs = env.reset()
            assert torch.equal(obs, whole.reset())
            rng = np.random.default_rng(1)
            done_steps = []
            for step in range(6):
                u = torch.as_tensor(rng.uniform(-1, 1, (E, 1)))
                o, r, dones, infos = env.step(u)
                wo, wr, wterm = whole.step_device(u)
                wdone = wterm or whole.truncation_pending()
                if wdone:
                    wo = whole.reset()
                    done_steps.append(step)
                assert torch.equal(o, wo) and torch.equal(r, wr) and dones == [wdone] * E and infos == [{}] * E
            if ends is not None:
                assert done_steps == ends
            else:   # only the LAST env (on rank 1) leaves the range: its truncation ends the episode for every rank
                assert done_steps and done_steps[0] < 3
            env.shutdown()
        # (2) the learner glue on top: PPO drives the global batch on rank 0 while rank 1 serves its shard
        off, cnt = shard_bounds(E, rank, world)
        env = ShardedVecEnv(_FakeShard(cnt, off), E)
        if rank == 0:
            ppo = PPO(env, seed=0, hidden=16, epochs=1, minibatches=1)
            hist = ppo.learn(2)
            assert len(hist) == 2 and all(np.isfinite(h["mean_step_reward"]) for h in hist)
            assert env.episodes == 2 and ppo.n_steps == 4
            env.shutdown()
        else:
            env.serve()
            assert env.episodes == 2
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_vec_env_serves_the_learner_gloo():
    """ShardedVecEnv (BASELINE config 5's learner protocol) over gloo, world size 2, uneven shards (4 + 3 envs)."""
    res = _run_ranks(_sharded_env_worker, 2, 7)
    assert res == [(0, "ok"), (1, "ok")], res
