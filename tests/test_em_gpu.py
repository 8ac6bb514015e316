"""GPU parity of the fused Euler–Maruyama kernel (qrl_em_step, through the C ABI) against the oracle and the
reference-generated fixtures.  Tolerance: FP64, rtol 1e-12 pathwise when the reference's own draws are fed
(north_star: "pathwise for stochastic envs when the reference's Wiener increments are fed in")."""
import os

import numpy as np
import pytest
import torch

import em_oracle as eo
import systems as sy
from helpers import dev, em_step
from quantrl_b200 import _native as N

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-12, 1e-14


@pytest.fixture(scope="module")
def g_em(golden_dir):
    return np.load(os.path.join(golden_dir, "em_linear_env.npz"))


@pytest.mark.parametrize("tag", ["n4", "n8"])
def test_fed_matches_reference_fixture(native, cuda, g_em, tag):
    g = g_em
    A0, Au, nz, w = g[f"{tag}_A0"], g[f"{tag}_Au"], g[f"{tag}_nz"], g[f"{tag}_weights"]
    K, S, ssz = int(g[f"{tag}_K"]), int(g[f"{tag}_S"]), float(g[f"{tag}_ssz"])
    Ws = np.transpose(g[f"{tag}_Ws"], (1, 0, 2))            # (S,E,n)
    ON = np.transpose(g[f"{tag}_obs_noises"], (1, 0, 2))    # (S+1,E,n)
    x = g[f"{tag}_x0"].copy()
    cum = np.zeros(x.shape[0])
    for a in range(S // K):
        t0 = a * K
        A = sy.affine_A(A0, Au, g[f"{tag}_actions"][:, a])
        out = em_step(x, A, nz, ssz, K, W=Ws[t0:t0 + K], obs_noise=ON[t0:t0 + K + 1],
                      reward_kind=N.REWARD_NEG_WSQ_OBS, reward_w=w, rewards_cum=cum)
        ref_S = np.transpose(g[f"{tag}_States"][:, t0:t0 + K + 1], (1, 0, 2))
        ref_O = np.transpose(g[f"{tag}_Observations"][:, t0:t0 + K + 1], (1, 0, 2))
        ref_R = g[f"{tag}_Reward"][:, t0 + 1:t0 + K + 1].T
        np.testing.assert_allclose(out["states"].cpu().numpy(), ref_S, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(out["observations"].cpu().numpy(), ref_O, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(out["reward"].cpu().numpy()[1:], ref_R, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(out["reward_last"].cpu().numpy(), ref_R[-1], rtol=RTOL, atol=ATOL)
        np.testing.assert_array_equal(out["x_last"].cpu().numpy(), out["states"][-1].cpu().numpy())
        np.testing.assert_array_equal(out["obs_last"].cpu().numpy(), out["observations"][-1].cpu().numpy())
        cum = cum + ref_R[-1]
        np.testing.assert_allclose(out["rewards_cum"].cpu().numpy(), cum, rtol=1e-12)
        mm = out["minmax"].cpu().numpy()
        assert mm[0] == ref_O.min() or abs(mm[0] - ref_O.min()) < 1e-12
        assert abs(mm[1] - ref_O.max()) < 1e-12
        x = out["x_last"].cpu().numpy()


@pytest.mark.parametrize("E,n,K", [(1, 4, 7), (33, 4, 50), (257, 8, 20), (5, 3, 11), (9, 1, 5), (130, 2, 16),
                                   (70, 6, 9), (40, 16, 12), (3, 13, 4)])
def test_fed_matches_oracle_seeded(native, cuda, E, n, K):
    rng = np.random.default_rng(E * 1000 + n)
    A = 0.4 * rng.standard_normal((E, n, n))
    nz = rng.uniform(0.01, 0.1, (E, n))
    W = 0.1 * rng.standard_normal((K, E, n))
    ON = 0.01 * rng.standard_normal((K + 1, E, n))
    x0 = rng.standard_normal((E, n))
    ref = eo.em_interval_vec(x0, A, nz, W, 0.01)
    out = em_step(x0, A, nz, 0.01, K, W=W, obs_noise=ON)
    np.testing.assert_allclose(out["states"].cpu().numpy(), ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out["observations"].cpu().numpy(), ref + ON, rtol=RTOL, atol=ATOL)
    assert int(out["nan"].item()) == 0


def test_time_dependent_and_shared_coefficients(native, cuda):
    rng = np.random.default_rng(5)
    E, n, K = 19, 4, 13
    A = 0.4 * rng.standard_normal((K, E, n, n))
    nz = rng.uniform(0.01, 0.1, (K, E, n))
    W = 0.1 * rng.standard_normal((K, E, n))
    x0 = rng.standard_normal((E, n))
    out = em_step(x0, A, nz, 0.02, K, W=W)
    np.testing.assert_allclose(out["states"].cpu().numpy(), eo.em_interval_vec(x0, A, nz, W, 0.02), rtol=RTOL, atol=ATOL)
    np.testing.assert_array_equal(out["observations"].cpu().numpy(), out["states"].cpu().numpy())  # no obs noise
    # shared (n,n) drift and (n,) prefixes (the reference's single-env hook shapes, stochastic.py:244-284)
    out = em_step(x0, A[0, 0], nz[0, 0], 0.02, K, W=W)
    ref = eo.em_interval_vec(x0, np.broadcast_to(A[0, 0], (E, n, n)), np.broadcast_to(nz[0, 0], (E, n)), W, 0.02)
    np.testing.assert_allclose(out["states"].cpu().numpy(), ref, rtol=RTOL, atol=ATOL)


def test_philox_pathwise_matches_oracle_stream(native, cuda):
    """In-kernel Philox4x32-10 + Box-Muller reproduces oracle/em_oracle.py:philox_normals (same integer stream;
    log/sincos differ by a few ulp between libm and CUDA, hence rtol 1e-10)."""
    rng = np.random.default_rng(8)
    E, n, K, dt = 37, 4, 25, 0.01
    A = 0.4 * rng.standard_normal((E, n, n))
    nz = rng.uniform(0.01, 0.1, (E, n))
    std = rng.uniform(0.01, 0.02, (E, n))
    x0 = rng.standard_normal((E, n))
    ph = dict(seed=0x1234567890ABCDEF, episode=3, t_idx=400, env_offset=1000)
    z_w, z_o = eo.philox_noise_block(ph["seed"], ph["episode"], ph["t_idx"], K, ph["env_offset"], E, n)
    ref = eo.em_interval_vec(x0, A, nz, np.sqrt(dt) * z_w, dt)
    out = em_step(x0, A, nz, dt, K, philox=ph, obs_std=std)
    np.testing.assert_allclose(out["states"].cpu().numpy(), ref, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["observations"].cpu().numpy(), ref + std * z_o, rtol=1e-10, atol=1e-12)


def test_philox_normals_accuracy_against_libm(native, cuda):
    """The generator itself, read out of the kernel exactly.  With A = 0, x0 = 0, dt = 1: nz = 1 makes the first state
    row 0 + 1 * (1 * z0) = z0 of the call with tau = t_idx; nz = 0 and obs_std = 1 make Observations[1] = 0 + z1.  Both
    kernels (tiled: tables in shared memory; generic: tables behind L1) against the extended-precision libm evaluation
    of the same stream (oracle/em_oracle.py:philox_normals): |dz| <= 5 * 2^-53 * radius (philox.cuh's accuracy claim, the
    same bound oracle/philox_port.c is held to on the CPU), and tiled == generic bit for bit."""
    E, n, K = 4096, 4, 1
    ph = dict(seed=0xFEEDFACE12345678, episode=5, t_idx=77, env_offset=3)
    z0s, z1s = [], []
    for flags in (0, N.EM_FORCE_GENERIC):
        out = em_step(np.zeros((E, n)), np.zeros((n, n)), np.ones(n), 1.0, K, philox=ph, obs_std=np.ones(n), flags=flags)
        z0s.append(out["states"][1].cpu().numpy())
        out = em_step(np.zeros((E, n)), np.zeros((n, n)), np.zeros(n), 1.0, K, philox=ph, obs_std=np.ones(n), flags=flags)
        z1s.append(out["observations"][1].cpu().numpy())
    np.testing.assert_array_equal(z0s[0], z0s[1])
    np.testing.assert_array_equal(z1s[0], z1s[1])
    tau = np.full((1, 1, 1), ph["t_idx"], dtype=np.uint64)
    env = (np.uint64(ph["env_offset"]) + np.arange(E, dtype=np.uint64)).reshape(1, -1, 1)
    comp = np.arange(n, dtype=np.uint64).reshape(1, 1, -1)
    r0, r1 = eo.philox_normals(ph["seed"], ph["episode"], tau, env, comp)
    radius = np.hypot(r0[0], r1[0])
    for z, r in ((z0s[0], r0[0]), (z1s[0], r1[0])):
        err = np.abs(z - r)
        assert np.max(err / radius) < 5 * 2.0 ** -53, np.max(err / radius) * 2.0 ** 53
        assert np.mean(err > 2 * np.spacing(np.abs(r))) < 0.01
    assert np.max(np.abs(np.hypot(z0s[0], z1s[0]) - radius) / radius) < 7 * 2.0 ** -53


def test_philox_sharding_invariance_full_size(native, cuda):
    """BASELINE config 3 size (E=4096, n=4, K=100): two half shards with env_offset == one full launch, bitwise."""
    rng = np.random.default_rng(1)
    E, n, K, dt = 4096, 4, 100, 0.01
    A0, Au = sy.affine_matrices(n, 1, seed=0)
    u = rng.uniform(-1, 1, (E, 1))
    A = sy.affine_A(A0, Au, u)
    nz = np.full((n,), 0.05)
    std = np.full((n,), 0.01)
    x0 = np.ones((E, n))
    ph = dict(seed=1, episode=0, t_idx=0, env_offset=0)
    full = em_step(x0, A, nz, dt, K, philox=ph, obs_std=std)
    h = E // 2
    lo = em_step(x0[:h], A[:h], nz, dt, K, philox=ph, obs_std=std)
    hi = em_step(x0[h:], A[h:], nz, dt, K, philox=dict(ph, env_offset=h), obs_std=std)
    assert torch.equal(full["states"][:, :h], lo["states"]) and torch.equal(full["states"][:, h:], hi["states"])
    assert torch.equal(full["observations"][:, h:], hi["observations"])
    # a different episode / seed gives a different path
    other = em_step(x0, A, nz, dt, K, philox=dict(ph, episode=1), obs_std=std)
    assert not torch.equal(full["states"][1:], other["states"][1:])


def test_philox_increment_moments_full_size(native, cuda):
    """Moment matching for the in-kernel RNG: recover z = (x_{k+1} - M x_k) / (nz sqrt(dt)) and the measurement
    noise and test mean/var/kurtosis/independence."""
    E, n, K, dt = 4096, 4, 100, 0.01
    A0, _ = sy.affine_matrices(n, 1, seed=0)
    nz, std = 0.05, 0.01
    out = em_step(np.ones((E, n)), A0, np.full((n,), nz), dt, K, philox=dict(seed=7, episode=2, t_idx=1000), obs_std=np.full((n,), std))
    X = out["states"]
    M = torch.eye(n, dtype=torch.float64, device="cuda") + dev(A0) * dt
    z = (X[1:] - torch.einsum("ij,kej->kei", M, X[:-1])) / (nz * np.sqrt(dt))
    v = (out["observations"] - X) / std
    for s in (z.flatten(), v.flatten()):
        m = s.numel()
        assert abs(s.mean().item()) < 5 / np.sqrt(m)
        assert abs(s.var().item() - 1) < 5 * np.sqrt(2 / m)
        assert abs((s ** 4).mean().item() - 3) < 5 * np.sqrt(96 / m)
        assert s.abs().max().item() < 7.0
    zc = z.reshape(K, E * n)
    assert abs(torch.corrcoef(torch.stack([zc[0], zc[1]]))[0, 1].item()) < 0.02          # across time
    assert abs(torch.corrcoef(torch.stack([z[:, 0, 0], z[:, 1, 0]]))[0, 1].item()) < 0.4  # across envs (K=100 samples)
    assert abs(torch.corrcoef(torch.stack([z.flatten(), v[1:].flatten()]))[0, 1].item()) < 0.01


def test_linearity_and_noise_free_euler_full_size(native, cuda):
    """Size-independent properties at E=4096, K=100: the fed map is linear in (x0, W); with W = 0 it is explicit
    Euler, x_K = (I + A dt)^K x0."""
    rng = np.random.default_rng(3)
    E, n, K, dt = 4096, 4, 100, 0.01
    A = 0.5 * rng.standard_normal((E, n, n))
    nz = rng.uniform(0.01, 0.1, (E, n))
    xa, xb = rng.standard_normal((E, n)), rng.standard_normal((E, n))
    Wa, Wb = 0.1 * rng.standard_normal((K, E, n)), 0.1 * rng.standard_normal((K, E, n))
    sa = em_step(xa, A, nz, dt, K, W=Wa)["states"]
    sb = em_step(xb, A, nz, dt, K, W=Wb)["states"]
    sab = em_step(xa + xb, A, nz, dt, K, W=Wa + Wb)["states"]
    torch.testing.assert_close(sab, sa + sb, rtol=1e-11, atol=1e-12)
    s0 = em_step(xa, A, nz, dt, K, W=np.zeros((K, E, n)))["x_last"]
    M = torch.eye(n, dtype=torch.float64, device="cuda") + dev(A) * dt
    ref = torch.einsum("eij,ej->ei", torch.linalg.matrix_power(M, K), dev(xa))
    torch.testing.assert_close(s0, ref, rtol=1e-11, atol=1e-12)


def test_edge_cases(native, cuda):
    n = 4
    A, nz = np.zeros((n, n)), np.zeros(n)
    # K = 0: row 0 only (observations_0 = states_0 + noise[0], base.py:426)
    x0 = np.arange(8.0).reshape(2, 4)
    out = em_step(x0, A, nz, 0.01, 0, obs_noise=np.full((1, 2, 4), 0.5))
    np.testing.assert_array_equal(out["obs_last"].cpu().numpy(), x0 + 0.5)
    np.testing.assert_array_equal(out["x_last"].cpu().numpy(), x0)
    # K = 0 with A = nz = NULL (how reset() asks for observations_0)
    xk = dev(x0)
    ol = torch.empty_like(xk)
    a = N.EmArgs(); a.n, a.n_envs, a.K, a.rng_mode, a.seed = 4, 2, 0, N.RNG_PHILOX, 3
    a.x0, a.obs_last = N.ptr(xk), N.ptr(ol)
    N.em_step(a); torch.cuda.synchronize()
    np.testing.assert_array_equal(ol.cpu().numpy(), x0)   # obs_std NULL -> no measurement noise
    # E = 0: nothing to do, no launch, no error
    before = N.launch_count()
    a = N.EmArgs(); a.n, a.n_envs, a.K = 4, 0, 5
    N.em_step(a)
    assert N.launch_count() == before
    # NaN / Inf detection and NaN-blind min/max (the reference's comparison is False for NaN, base.py:495-499)
    x0 = np.array([[1.0, 2.0, np.nan, 4.0], [1e300, 1e300, 1e300, 1e300]])
    out = em_step(x0, 1e10 * np.ones((n, n)), nz, 1.0, 3, W=np.zeros((3, 2, n)))
    assert int(out["nan"].item()) == 1
    mm = out["minmax"].cpu().numpy()
    assert mm[1] == np.inf and mm[0] == 1.0
    # trajectory off
    out = em_step(np.ones((3, n)), A, nz, 0.01, 4, W=np.zeros((4, 3, n)), traj=False)
    np.testing.assert_array_equal(out["x_last"].cpu().numpy(), np.ones((3, n)))


def test_invalid_arguments_raise(native, cuda):
    a = N.EmArgs(); a.n, a.n_envs, a.K = 4, 2, 1
    with pytest.raises(N.NativeError, match="x0 is NULL"):
        N.em_step(a)
    a.n = 17
    with pytest.raises(N.NativeError, match="n must be in 1..16"):
        N.em_step(a)
    a.n, a.rng_mode = 4, 9
    with pytest.raises(N.NativeError, match="bad rng_mode"):
        N.em_step(a)
    with pytest.raises(N.NativeError, match="CUDA tensor"):
        N.ptr(torch.zeros(3))


@pytest.mark.parametrize("E,n,K", [(4096, 4, 100), (300, 8, 37), (65, 3, 12), (1000, 2, 64), (33, 5, 7), (5000, 4, 20),
                                   (9001, 4, 30), (31, 7, 3), (2, 6, 1)])
def test_tiled_kernel_equals_generic_kernel_bitwise(native, cuda, E, n, K):
    """The tiled warp-specialised kernel and the generic lane-per-component kernel run the same operations in the
    same order: States/Observations must agree bit for bit (Philox and fed), rewards to rounding."""
    rng = np.random.default_rng(E + n)
    A = 0.4 * rng.standard_normal((E, n, n))
    nz = rng.uniform(0.01, 0.1, (E, n))
    std = rng.uniform(0.01, 0.02, (E, n))
    x0 = rng.standard_normal((E, n))
    w = rng.uniform(0.5, 1.5, n)
    for mode in ("philox", "fed"):
        kw = dict(philox=dict(seed=99, episode=1, t_idx=300, env_offset=17), obs_std=std) if mode == "philox" else \
            dict(W=0.1 * rng.standard_normal((K, E, n)), obs_noise=0.01 * rng.standard_normal((K + 1, E, n)))
        t = em_step(x0, A, nz, 0.01, K, reward_kind=N.REWARD_NEG_WSQ_OBS, reward_w=w, flags=N.EM_FORCE_TILED, **kw)
        g = em_step(x0, A, nz, 0.01, K, reward_kind=N.REWARD_NEG_WSQ_OBS, reward_w=w, flags=N.EM_FORCE_GENERIC, **kw)
        for key in ("states", "observations", "x_last", "obs_last", "minmax"):
            assert torch.equal(t[key], g[key]), (mode, key)
        for key in ("reward", "reward_last", "rewards_cum"):
            torch.testing.assert_close(t[key], g[key], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("flags", [0, N.EM_FORCE_GENERIC, N.EM_PACK_TMA])
@pytest.mark.parametrize("idxs,pitch", [([0, 1, 2, 3, 4, 5, 6], 7), ([0, 1, 2, 3, 4, 5, 6], 8), ([-1], 1), ([-1], 2),
                                        ([0, 3, -1, 1], 4), ([0, 3, -1, 1], 6)])
def test_fused_cache_rows_match_oracle_pack_rows(native, cuda, flags, idxs, pitch):
    """Packed rows written by the integration launch == oracle pack_rows (BaseSB3Env.update_data layout), written
    at a row offset of a larger (E, shape_T, n_sel) cache like the env does."""
    rng = np.random.default_rng(21)
    E, n, K = 77, 4, 23
    A = 0.4 * rng.standard_normal((E, n, n))
    nz = rng.uniform(0.01, 0.1, (E, n))
    x0 = rng.standard_normal((E, n))
    W, ON = 0.1 * rng.standard_normal((K, E, n)), 0.01 * rng.standard_normal((K + 1, E, n))
    T_step = 0.5 + np.arange(K + 1) * 0.01
    acts = rng.uniform(-1, 1, (E, 1))
    cum0 = rng.standard_normal(E)
    out = em_step(x0, A, nz, 0.01, K, W=W, obs_noise=ON, reward_kind=N.REWARD_NEG_WSQ_OBS, reward_w=np.ones(n),
                  rewards_cum=cum0, flags=flags,
                  pack=dict(T_step=T_step, actions=acts, idxs=idxs, rows_total=K + 11, row0=5 + (pitch & 1), pitch=pitch))
    obs = eo.em_interval_vec(x0, A, nz, W, 0.01) + ON
    cum = cum0 - np.sum(obs[-1] ** 2, -1)
    full = eo.pack_rows(T_step, acts, obs, None, cum)
    got = out["packed"].cpu().numpy()
    r0 = 5 + (pitch & 1)
    np.testing.assert_allclose(got[:, r0:r0 + K + 1, :len(idxs)], full[:, :, idxs], rtol=1e-12, atol=1e-14)
    # nothing outside the block is touched; padding columns of a pitched cache are unspecified (untouched or zero)
    assert np.isnan(got[:, :r0]).all() and np.isnan(got[:, r0 + K + 1:]).all()
    pad = got[:, r0:r0 + K + 1, len(idxs):]
    assert np.all(np.isnan(pad) | (pad == 0.0))


@pytest.mark.parametrize("flags", [N.EM_FORCE_TILED, N.EM_FORCE_GENERIC, N.EM_FORCE_TILED | N.EM_ITEM_OUTPUT])
def test_last_block_bookkeeping_advances_dyn_and_delivers_minmax(native, cuda, flags):
    """ABI 2 ``finish_counter`` protocol: the last block of the launch adds ``dyn_advance`` to dyn[0], copies this
    launch's [min, max] to ``minmax_out``, re-arms the accumulator and leaves the counter at zero — three launches in
    a row (as a replayed CUDA graph issues them) equal three launches with host-side bookkeeping, bit for bit."""
    rng = np.random.default_rng(5)
    E, n, K = 300, 4, 40
    A = torch.as_tensor(0.4 * rng.standard_normal((E, n, n)), device="cuda")
    nz = torch.as_tensor(rng.uniform(0.01, 0.1, (E, n)), device="cuda")
    std = torch.as_tensor(rng.uniform(0.01, 0.02, (E, n)), device="cuda")
    x = torch.as_tensor(rng.standard_normal((E, n)), device="cuda")
    x_ref = x.clone()
    dyn = torch.tensor([7, 3], dtype=torch.int64, device="cuda")
    counter = torch.zeros((1,), dtype=torch.int32, device="cuda")
    acc = torch.tensor([float("inf"), float("-inf")], dtype=torch.float64, device="cuda")
    mm_out = torch.full((2,), float("nan"), dtype=torch.float64, device="cuda")
    obs = torch.empty((K + 1, E, n), dtype=torch.float64, device="cuda")
    obs_ref = torch.empty_like(obs)

    def args(x0, obs_buf):
        a = N.EmArgs()
        a.n, a.n_envs, a.K, a.t_ssz = n, E, K, 0.01
        a.A, a.A_stride_e, a.nz, a.nz_stride_e = N.ptr(A), n * n, N.ptr(nz), n
        a.rng_mode, a.seed, a.obs_std, a.obs_std_stride_e = N.RNG_PHILOX, 11, N.ptr(std), n
        a.x0, a.x_last, a.observations = N.ptr(x0), N.ptr(x0), N.ptr(obs_buf)
        a.flags = flags
        return a

    for step in range(3):
        a = args(x, obs)
        a.dyn, a.finish_counter, a.dyn_advance = N.ptr(dyn), N.ptr(counter), K
        a.obs_minmax, a.minmax_out = N.ptr(acc), N.ptr(mm_out)
        N.em_step(a)
        b = args(x_ref, obs_ref)
        mm_ref = torch.tensor([float("inf"), float("-inf")], dtype=torch.float64, device="cuda")
        b.t_idx, b.episode, b.obs_minmax = 7 + step * K, 3, N.ptr(mm_ref)
        N.em_step(b)
        torch.cuda.synchronize()
        assert torch.equal(obs, obs_ref) and torch.equal(x, x_ref)
        assert dyn.tolist() == [7 + (step + 1) * K, 3] and counter.item() == 0
        assert torch.equal(mm_out, mm_ref) and mm_out[0].item() == obs.min().item() and mm_out[1].item() == obs.max().item()
        assert acc.tolist() == [float("inf"), float("-inf")]
