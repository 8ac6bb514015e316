"""Environment sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

Environments are independent units (no cross-env term in the reference's RHS / EM update: quantrl/envs/
deterministic.py:653-746, quantrl/envs/stochastic.py:196-242), so the hot path shards with NO data-path collective:
rank ``r`` owns the contiguous block ``[offset_r, offset_r + count_r)`` of the global batch and passes ``env_offset``
to the kernels, whose Philox counters use the GLOBAL env index — results do not depend on the number of ranks.

The only exchange is the learner protocol of one action step:

    learner (rank ``dst``)          workers
    scatter  actions (E, n_act)  -> local slices
                                    env.step_device(local actions)          # fused kernel, no collective
    gather   [obs | reward | done-flag]  <- (E_loc, n_obs + 2) per rank      # NCCL over NVLink / gloo on CPU

The truncation flag rides in the same buffer (column ``n_obs + 1``): the reference ends the episode for ALL envs
when any observation leaves the range (quantrl/envs/base.py:485-499, 1271-1293), so the learner ORs the flags.
"""
import os
import torch
import torch.distributed as dist


def shard_bounds(n_envs_total: int, rank: int, world_size: int):
    """Contiguous block of rank ``rank``: (offset, count).  The first ``n % world`` ranks get one extra env."""
    assert 0 <= rank < world_size and n_envs_total >= 0
    base, extra = divmod(n_envs_total, world_size)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def all_shard_bounds(n_envs_total: int, world_size: int):
    return [shard_bounds(n_envs_total, r, world_size) for r in range(world_size)]


class LearnerLink:
    """Scatter actions from / gather (obs, reward, truncation flag) to the learner rank, once per action step."""

    def __init__(self, n_envs_total, n_observations, n_actions, device, dst=0, group=None, dtype=torch.float64):
        assert dist.is_initialized(), "torch.distributed is not initialised"
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dst, self.n_obs, self.n_act, self.E = dst, n_observations, n_actions, n_envs_total
        self.bounds = all_shard_bounds(n_envs_total, self.world)
        self.offset, self.count = self.bounds[self.rank]
        self.device, self.dtype = torch.device(device), dtype
        width = n_observations + 2
        self._local = torch.zeros((self.count, width), dtype=dtype, device=self.device)
        self._local_actions = torch.zeros((self.count, n_actions), dtype=dtype, device=self.device)
        self.is_learner = self.rank == dst
        # uneven shards are padded to the largest block so that one all_gather_into_tensor / gather call suffices
        self._max = max(c for _, c in self.bounds)
        self._send = torch.zeros((self._max, width), dtype=dtype, device=self.device)
        self._recv = torch.zeros((self.world, self._max, width), dtype=dtype, device=self.device) if self.is_learner else None
        self._act_pad = torch.zeros((self.world, self._max, n_actions), dtype=dtype, device=self.device) if self.is_learner else None
        self._act_recv = torch.zeros((self._max, n_actions), dtype=dtype, device=self.device)

    def scatter_actions(self, actions_global=None):
        """Learner passes (E, n_act); every rank returns its (E_loc, n_act) slice."""
        if self.world == 1:
            return actions_global.to(self.device, self.dtype).reshape(self.E, self.n_act)
        scatter_list = None
        if self.is_learner:
            a = actions_global.to(self.device, self.dtype).reshape(self.E, self.n_act)
            for r, (off, cnt) in enumerate(self.bounds):
                self._act_pad[r, :cnt] = a[off:off + cnt]
            scatter_list = list(self._act_pad.unbind(0))
        dist.scatter(self._act_recv, scatter_list, src=self.dst, group=self.group)
        self._local_actions.copy_(self._act_recv[:self.count])
        return self._local_actions

    def gather_step(self, obs_local, reward_local, truncated_local=False):
        """Every rank passes its (E_loc, n_obs) observations, (E_loc,) rewards and truncation flag.  The learner gets
        ``(obs (E, n_obs), reward (E,), truncated: bool)`` in global env order, the other ranks get ``None``."""
        n = self.n_obs
        self._send[:self.count, :n] = obs_local
        self._send[:self.count, n] = reward_local
        self._send[:self.count, n + 1] = float(bool(truncated_local))
        if self.world == 1:
            return obs_local, reward_local, bool(truncated_local)
        gather_list = list(self._recv.unbind(0)) if self.is_learner else None
        dist.gather(self._send, gather_list, dst=self.dst, group=self.group)
        if not self.is_learner:
            return None
        parts = [self._recv[r, :cnt] for r, (_, cnt) in enumerate(self.bounds)]
        full = torch.cat(parts, dim=0)
        return full[:, :n], full[:, n], bool((full[:, n + 1] > 0).any().item()) if full.shape[0] else False


def make_sharded_env(env_cls, n_envs_total, rank=None, world_size=None, **env_kwargs):
    """Build this rank's shard of a ``n_envs_total`` batch: ``n_envs`` and ``env_offset`` are filled in, per-env array
    arguments named in ``env_kwargs['_per_env']`` are sliced to the local block."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    offset, count = shard_bounds(n_envs_total, rank, world_size)
    per_env = env_kwargs.pop("_per_env", ())
    for key in per_env:
        if env_kwargs.get(key) is not None:
            env_kwargs[key] = env_kwargs[key][offset:offset + count]
    return env_cls(n_envs=count, env_offset=offset, n_envs_total=n_envs_total, **env_kwargs)


class ShardedVecEnv:
    """The learner rank's view of ONE lock-stepped batch whose environments live on all ranks (BASELINE config 5:
    "observations/rewards all-gathered to the learner rank").

    Wraps this rank's shard (an env of ``quantrl_b200.envs`` built by ``make_sharded_env``) and speaks the same
    ``reset() / step(actions) -> (obs, reward, dones, infos)`` protocol as a single env (quantrl/envs/base.py:1185-1293),
    for the GLOBAL batch: on the learner ``step`` takes ``(n_envs_total, n_act)`` actions and returns global
    observations / rewards; the other ranks call ``serve()``, which mirrors the learner's calls until ``shutdown()``.

    Per action step: scatter actions -> ``step_device`` on every shard (the fused kernel, no collective) -> one 1-word
    MAX all-reduce of the truncation flag (the reference ends the episode for ALL envs when any observation leaves the
    range, base.py:485-499, 1271-1293; termination is the same on every rank by construction) -> gather
    ``[obs | reward]`` to the learner.  When the episode ends every shard resets and the reset observation is what
    the learner receives (auto-reset, no ``terminal_observation``, base.py:1277-1293)."""

    _STOP = float("nan")

    def __init__(self, local_env, n_envs_total, dst=0, group=None, link=None):
        self.env = local_env
        self.device = torch.device(local_env.device)
        self.n_envs, self.n_observations, self.n_actions = int(n_envs_total), local_env.n_observations, local_env.n_actions
        self.action_steps, self.action_space_range = local_env.action_steps, local_env.action_space_range
        self.link = link if link is not None else LearnerLink(n_envs_total, self.n_observations, self.n_actions,
                                                              device=self.device, dst=dst, group=group)
        assert self.link.count == local_env.n_envs, "the local env does not have this rank's share of the batch"
        self.is_learner = self.link.is_learner
        self._flag = torch.zeros((1,), dtype=torch.float64, device=self.device)
        self._zero_reward = torch.zeros((local_env.n_envs,), dtype=torch.float64, device=self.device)
        self.episodes = 0

    def _gather(self, obs, rew):
        res = self.link.gather_step(obs, rew, False)
        return None if res is None else (res[0].clone(), res[1].clone())

    def reset(self, seed=None, options=None):
        obs = self.env.reset()
        res = self._gather(obs, self._zero_reward)
        return None if res is None else res[0]

    def step(self, actions=None):
        """Learner: ``actions (n_envs_total, n_act)`` -> ``(obs, reward, dones, infos)`` of the global batch.
        Other ranks (through ``serve``): returns ``None``, or ``False`` once the learner has shut the loop down."""
        a = self.link.scatter_actions(actions)
        if self.link.world > 1 and bool(torch.isnan(a).any()):
            return False
        obs, rew, terminated = self.env.step_device(a)
        self._flag[0] = 1.0 if self.env.truncation_pending() else 0.0
        if self.link.world > 1:
            dist.all_reduce(self._flag, op=dist.ReduceOp.MAX, group=self.link.group)
        done = bool(terminated) or bool(self._flag[0] > 0)
        rew = rew.clone()
        if done:
            self.episodes += 1
            obs = self.env.reset()
        res = self._gather(obs, rew)
        if res is None:
            return None
        return res[0], res[1], [done] * self.n_envs, [{}] * self.n_envs

    def serve(self):
        """Non-learner ranks: follow the learner (its first ``reset`` and every ``step``) until it calls ``shutdown``."""
        assert not self.is_learner
        self.reset()
        while self.step(None) is not False:
            pass

    def shutdown(self):
        """Learner: release the ranks blocked in ``serve``."""
        assert self.is_learner
        if self.link.world > 1:
            self.link.scatter_actions(torch.full((self.n_envs, self.n_actions), self._STOP, dtype=torch.float64,
                                                 device=self.device))

    def close(self, **kw):
        self.env.close(**kw)


class PeerGather:
    """Gather-to-learner WITHOUT a collective launch and without a barrier (the one exchange step of the sharded path).

    Every rank keeps the last rows of an action step — ``[obs (E_loc, n) | reward (E_loc,)]`` — in its OWN memory, in a
    ring of ``SLOTS`` slices of torch symmetric memory; they reach the learner's gathered block
    ``obs_all (E, n) | reward_all (E,)`` in one of two ways:

    * ``mode='push'`` (default): copy engines.  After every integration launch the rank enqueues, on a side stream that
      waits for that launch through an event, a peer-to-peer ``cudaMemcpyAsync`` of its slice into the learner's block
      and an 8-byte epoch word into the learner's flag array (``qrl_peer_push``).  The step kernel is untouched, no SM is
      involved, the transfer of step ``i`` overlaps step ``i+1``.  The learner's block is double buffered by epoch parity.
    * ``mode='pull'``: kernels.  The step kernel's last block stores the epoch word itself (``QRL_EM_PEER_PUBLISH``) and the
      learner runs ``qrl_peer_pull`` on a side stream, which polls the flags and copies the slices with 16-byte peer loads;
      credits (``pulled``) let a rank reuse a slice.  (Kept as the measured comparison: a polling gather block has to be
      co-resident with the next step's CTAs, which costs the learner a few microseconds per step.)

    ``wait()`` on the learner makes the current stream wait until every rank's rows of the latest epoch have arrived.
    The NCCL path (``LearnerLink.gather_step``) stays as the fallback and as the measured comparison.
    """

    SLOTS = 32          # ring depth = how many steps the host may run ahead of the copies (jitter tolerance: 0.7 ms at config 3), 32 x 164 KB

    def __init__(self, n_envs_total, n_observations, device, dst=0, group=None, mode="push"):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _native as N
        assert dist.is_initialized() and mode in ("push", "pull")
        self._N, self.mode = N, mode
        self.group = group if group is not None else dist.group.WORLD
        gname = self.group.group_name if hasattr(self.group, "group_name") else self.group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.dst, self.E, self.n_obs = dst, n_envs_total, n_observations
        self.device = torch.device(device)
        self.bounds = all_shard_bounds(n_envs_total, self.world)
        self.offset, self.count = self.bounds[self.rank]
        self.is_learner = self.rank == dst
        n = n_observations
        words = max(c for _, c in self.bounds) * (n + 1)
        self.slot_words = words + (words & 1)                       # even: every slice starts 16-byte aligned
        self.ring = symm_mem.empty(self.SLOTS * self.slot_words, dtype=torch.float64, device=device)
        self.ring.zero_()
        self.ring_hdl = symm_mem.rendezvous(self.ring, gname)
        # flags[0:world] = epoch published by rank p (only the learner's copy is written), flags[world] = `pulled` credit
        self.flags = symm_mem.empty(self.world + 2, dtype=torch.int64, device=device)
        self.flags.zero_()
        self.flags_hdl = symm_mem.rendezvous(self.flags, gname)
        self.flag_ptrs = torch.tensor([int(p) for p in self.flags_hdl.buffer_ptrs], dtype=torch.int64, device=device)
        self.pulled_ptr = self.flags.data_ptr() + 8 * self.world
        self.epoch = 0                                              # launches published by this rank
        self._direct = False                                        # set by attach(..., wait_in_step=True)
        self._slot_views = []
        for s in range(self.SLOTS):
            base = s * self.slot_words
            self._slot_views.append((self.ring[base:base + self.count * n].view(self.count, n),
                                     self.ring[base + self.count * n:base + self.count * (n + 1)]))
        self._status = torch.zeros((1,), dtype=torch.int32, device=device)
        self.side = torch.cuda.Stream(device=device)
        self._side_raw = C.c_void_p(self.side.cuda_stream)
        gwords = n_envs_total * (n + 1)
        self._gwords = gwords + (gwords & 1)
        if mode == "push":
            # the learner's gathered block, double buffered by epoch parity (symmetric: every rank maps the learner's copy)
            self.gathered = symm_mem.empty(2 * self._gwords, dtype=torch.float64, device=device)
            self.gathered.zero_()
            self.gathered_hdl = symm_mem.rendezvous(self.gathered, gname)
            self._g_dst = int(self.gathered_hdl.buffer_ptrs[dst])
            self._flag_dst = int(self.flags_hdl.buffer_ptrs[dst]) + 8 * self.rank
            self._epoch_words = torch.zeros((self.SLOTS,), dtype=torch.int64).pin_memory()
            self._epoch_np = self._epoch_words.numpy()              # (a NumPy store costs a fraction of a tensor item store)
            self._ev_step = [torch.cuda.Event() for _ in range(self.SLOTS)]
            self._ev_copy = [torch.cuda.Event() for _ in range(self.SLOTS)]
            for ev in self._ev_step + self._ev_copy:
                ev.record()                                         # materialise the cudaEvent_t handles
            self._ctx = N.peer_ctx_create(self._side_raw)          # worker thread: the copies' driver calls leave this thread
            self._slot_seq = [0] * self.SLOTS                       # sequence number of the job that last used a ring slice
            self._push = []
            for s in range(self.SLOTS):
                a = N.PeerPushArgs()
                obs_v, rew_v = self._slot_views[s]
                a.src, a.bytes = obs_v.data_ptr(), 8 * self.count * n
                a.src2, a.bytes2 = rew_v.data_ptr(), 8 * self.count
                a.flag_src = self._epoch_words.data_ptr() + 8 * s
                a.flag_dst = self._flag_dst
                a.step_done, a.copy_done = self._ev_step[s].cuda_event, self._ev_copy[s].cuda_event
                # destination of slot s: SLOTS is even, so the parity of an epoch is the parity of its slot
                base = self._g_dst + 8 * (s % 2) * self._gwords
                a.dst = base + 8 * self.offset * n
                a.dst2 = base + 8 * (n_envs_total * n + self.offset)
                if self.is_learner:
                    a.bytes = a.bytes2 = 0      # rows already in place; only the epoch word follows the launch
                self._push.append(a)
            assert self.SLOTS % 2 == 0
            # what next_launch returns for the learner, per epoch parity (its launch writes straight into the gathered block)
            self._learner_out = [(self.gathered.data_ptr() + 8 * par * self._gwords + 8 * self.offset * n,
                                  self.gathered.data_ptr() + 8 * par * self._gwords + 8 * (n_envs_total * n + self.offset))
                                 for par in (0, 1)]
            self._slot_ptrs = [(o.data_ptr(), r.data_ptr()) for o, r in self._slot_views]
        elif self.is_learner:
            i64 = dict(dtype=torch.int64, device=device)
            self.gathered = torch.zeros((self._gwords,), dtype=torch.float64, device=device)
            self._ring_ptrs = torch.tensor([int(p) for p in self.ring_hdl.buffer_ptrs], **i64)
            self._strides = torch.full((self.world,), self.slot_words, **i64)
            self._offsets = torch.tensor([o for o, _ in self.bounds], **i64)
            self._counts = torch.tensor([c for _, c in self.bounds], **i64)
            self._pulled_ptrs = torch.tensor([int(p) + 8 * self.world for p in self.flags_hdl.buffer_ptrs], **i64)
            self._pull_epoch = torch.zeros((1,), **i64)
            self._finish = torch.zeros((1,), dtype=torch.int32, device=device)
            a = N.PeerPullArgs()
            a.world, a.n_obs, a.out_slots, a.blocks_per_rank = self.world, n, self.SLOTS, 8
            a.flags, a.peer_ring, a.slot_stride = self.flags.data_ptr(), N.ptr(self._ring_ptrs), N.ptr(self._strides)
            a.env_offset, a.env_count = N.ptr(self._offsets), N.ptr(self._counts)
            a.obs_all, a.reward_all = self.gathered.data_ptr(), self.gathered.data_ptr() + 8 * n_envs_total * n
            a.pull_epoch = N.ptr(self._pull_epoch)
            a.peer_pulled, a.finish_counter, a.status = N.ptr(self._pulled_ptrs), N.ptr(self._finish), N.ptr(self._status)
            self._pull_args = a
        torch.cuda.synchronize(device)
        self.ring_hdl.barrier(channel=0)      # everybody's ring, flags and gathered block are zeroed before the first epoch

    # ---- the learner's view of the latest epoch -----------------------------------------------------------------------
    def _block(self):
        base = (self.epoch % 2) * self._gwords if self.mode == "push" else 0
        return self.gathered[base:base + self._gwords]

    @property
    def obs_all(self):
        return self._block()[:self.E * self.n_obs].view(self.E, self.n_obs)

    @property
    def reward_all(self):
        return self._block()[self.E * self.n_obs:self.E * (self.n_obs + 1)]

    def attach(self, env, wait_in_step=False):
        """Route the env's step outputs through the ring (before its first step).  ``wait_in_step``: on the learner,
        ``env.step()`` also enqueues the gather wait behind its launch, so that the host synchronisation that ends the step
        covers "every rank's rows of this step have arrived" (one sync instead of two)."""
        assert env.n_envs == self.count and env.n_observations == self.n_obs
        env.attach_peer(self)
        # step(host actions) with the learner waiting for the rows of the SAME step: the transfer is latency critical, so the
        # rank enqueues it itself right behind its launch (the worker thread polls with 20 us sleeps; its driver calls are
        # hidden behind the running step kernel here anyway)
        self._direct = bool(wait_in_step) and self.mode == "push"
        if self._direct and not self.is_learner:
            # ... and the rows do not take the copy engine at all: the launch's second copy of the last rows goes straight
            # into this rank's slice of the LEARNER's gathered block (peer-mapped stores over NVLink, complete when the kernel
            # is), so only the 8-byte epoch word follows the launch — one copy-engine operation between "kernel done" and "the
            # learner sees the step" instead of three (measured at 2 GPUs: the learner waited ~10 us per step for the chain)
            n = self.n_obs
            self._learner_out = [(self._g_dst + 8 * par * self._gwords + 8 * self.offset * n,
                                  self._g_dst + 8 * par * self._gwords + 8 * (self.E * n + self.offset)) for par in (0, 1)]
            for a in self._push:
                a.bytes = a.bytes2 = 0
        if wait_in_step and self.is_learner:
            def hook():
                self.gather()
                self.wait()
            env._pre_sync_hook = hook

    # ---- called by the env around every integration launch it enqueues ------------------------------------------------------
    def next_launch(self):
        """``(epoch, obs_ptr, reward_ptr, pulled_min)`` of the ring slice the next launch writes."""
        self.epoch = e = self.epoch + 1
        s = e % self.SLOTS
        if self.mode == "push":
            if self._slot_seq[s]:
                # the copy that last read this slice (epoch e - SLOTS) has long finished; this returns at once
                if self._slot_seq[s] < 0:
                    self._N.peer_event_wait(self._ev_copy[s].cuda_event)      # (a direct push: its copy_done event)
                else:
                    self._N.peer_ctx_wait(self._ctx, self._slot_seq[s])
            if self.is_learner or self._direct:
                # the learner's own rows need no transfer: its launch writes them straight into its slice of the gathered
                # block (a local device-to-device copy would run on SMs and hold up the next step's CTAs: 36.6 vs 27.6 us);
                # host-step mode: every rank's launch does (see attach)
                o, r = self._learner_out[e & 1]
            else:
                o, r = self._slot_ptrs[s]
            return e, o, r, 0
        obs_v, rew_v = self._slot_views[s]
        return e, obs_v.data_ptr(), rew_v.data_ptr(), max(0, e - self.SLOTS)

    def after_launch(self, main_stream_raw):
        """push mode: enqueue the transfer of the epoch just launched (events + copy engine, no kernel)."""
        if self.mode != "push":
            return
        e = self.epoch
        s = e % self.SLOTS
        self._epoch_np[s] = e               # (source, destination and sizes of slot s were fixed at construction)
        if self._direct:
            self._N.peer_push(self._push[s], main_stream_raw, self._side_raw)
            self._slot_seq[s] = -1
        else:
            self._slot_seq[s] = self._N.peer_ctx_push(self._ctx, self._push[s], main_stream_raw)

    # ---- learner ------------------------------------------------------------------------------------------------------
    def gather(self):
        """Learner, once per published step, in order: pull mode enqueues the gather kernel on the side stream; push mode
        has nothing to do (every rank pushes its own rows)."""
        if self.mode == "pull" and self.is_learner:
            self._N.peer_pull(self._pull_args, self._side_raw)

    pull = gather

    def wait(self):
        """Learner: the current stream waits until the rows of the latest epoch of every rank are in ``obs_all`` /
        ``reward_all``.  Other ranks: nothing."""
        if not self.is_learner:
            return
        if self.mode == "pull":
            torch.cuda.current_stream().wait_stream(self.side)
        else:
            self._N.peer_wait(self.flags.data_ptr(), self.world, self.epoch, self._N.ptr(self._status))

    def close(self):
        if self.mode == "push" and getattr(self, "_ctx", None) is not None:
            self._N.peer_ctx_wait(self._ctx, max(0, max(self._slot_seq)))
            self._N.peer_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def check(self):
        """Synchronise and raise if an exchange step timed out (a rank never delivered)."""
        if self.mode == "push" and getattr(self, "_ctx", None) is not None:
            self._N.peer_ctx_wait(self._ctx, max(0, max(self._slot_seq)))   # every queued transfer has been issued
        self.side.synchronize()
        torch.cuda.current_stream().synchronize()
        if int(self._status.item()) & self._N.STATUS_PEER_TIMEOUT:
            raise RuntimeError("quantrl_b200: the learner gather timed out waiting for a rank's step result")
