"""CPU: the plain-C restatement of the stochastic path (oracle/em_port.c) against the NumPy oracle and the reference's
own ``LinearEnv`` fixture — two independently written oracles have to agree before either checks the kernel."""
import os

import numpy as np
import pytest

import em_oracle as eo
import em_port
import systems as sy

pytestmark = pytest.mark.skipif(not em_port.usable(), reason="gcc (with OpenMP) not available")


def test_c_port_equals_numpy_oracle():
    rng = np.random.default_rng(0)
    for n, E, K in ((4, 37, 25), (8, 5, 10), (1, 3, 4), (16, 2, 3)):
        A = 0.4 * rng.standard_normal((E, n, n))
        nz = rng.uniform(0.01, 0.1, (E, n))
        W = 0.1 * rng.standard_normal((K, E, n))
        ON = 0.01 * rng.standard_normal((K + 1, E, n))
        w = rng.uniform(0.5, 1.5, n)
        x0 = rng.standard_normal((E, n))
        St, Ob, Rw = em_port.em_interval_fed(x0, A, nz, W, 0.01, obs_noise=ON, weights=w)
        ref = eo.em_interval_vec(x0, A, nz, W, 0.01)
        np.testing.assert_allclose(St, ref, rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(Ob, ref + ON, rtol=1e-12, atol=1e-14)     # obs = x + noise can cancel
        np.testing.assert_allclose(Rw, eo.reward_neg_weighted_sq(ref + ON, w), rtol=1e-12, atol=1e-15)
    # K = 0: row 0 only
    St, Ob, Rw = em_port.em_interval_fed(x0, A, nz, W[:0], 0.01, obs_noise=ON[:1])
    assert St.shape == (1, E, n) and np.array_equal(St[0], x0) and np.array_equal(Ob[0], x0 + ON[0])


@pytest.mark.parametrize("tag", ["n4", "n8"])
def test_c_port_matches_reference_linear_env(golden_dir, tag):
    """Fixture G3 (reference ``LinearEnv``, its own draws) through the C port, interval by interval."""
    g = np.load(os.path.join(golden_dir, "em_linear_env.npz"))
    A0, Au, nz, w = g[f"{tag}_A0"], g[f"{tag}_Au"], g[f"{tag}_nz"], g[f"{tag}_weights"]
    K, S, ssz = int(g[f"{tag}_K"]), int(g[f"{tag}_S"]), float(g[f"{tag}_ssz"])
    x = g[f"{tag}_x0"].copy()
    for a in range(S // K):
        t0 = a * K
        St, Ob, Rw = em_port.em_interval_fed(
            x, sy.affine_A(A0, Au, g[f"{tag}_actions"][:, a]), nz, np.transpose(g[f"{tag}_Ws"][:, t0:t0 + K], (1, 0, 2)), ssz,
            obs_noise=np.transpose(g[f"{tag}_obs_noises"][:, t0:t0 + K + 1], (1, 0, 2)), weights=w)
        np.testing.assert_allclose(np.transpose(St, (1, 0, 2)), g[f"{tag}_States"][:, t0:t0 + K + 1], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(np.transpose(Ob, (1, 0, 2)), g[f"{tag}_Observations"][:, t0:t0 + K + 1], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(Rw[1:].T, g[f"{tag}_Reward"][:, t0 + 1:t0 + K + 1], rtol=1e-12, atol=1e-14)
        x = St[-1]


def test_c_port_rng_mode_has_the_right_moments():
    """Noise-free drift + the port's own generator: an Ornstein–Uhlenbeck ensemble reaches the stationary variance
    ``sigma^2 / (2 theta)`` (up to the Euler–Maruyama bias) and the measurement noise has the requested std."""
    E, n, K, dt, theta, sigma = 20000, 2, 400, 0.01, 1.0, 0.5
    st = em_port.RngStepper(E, n, K, dt, -theta * np.eye(n), np.full(n, sigma), np.full(n, 0.2), np.ones(n), seed=3)
    for _ in range(3):
        st.step()
    var = st.St[-1].var(axis=0)
    np.testing.assert_allclose(var, sigma ** 2 / (2 * theta - theta ** 2 * dt), rtol=0.05)
    np.testing.assert_allclose((st.Ob[-1] - st.St[-1]).std(axis=0), 0.2, rtol=0.03)
    assert abs(np.corrcoef(st.St[-1][:, 0], st.St[-1][:, 1])[0, 1]) < 0.03


@pytest.mark.parametrize("substeps,key", [(1, "rk4"), (2, "rk4x2")])
def test_c_lyapunov_port_matches_rk4_on_the_reference_func(golden_dir, substeps, key):
    """oracle/lyap_port.c (optomech RHS + classic RK4 in plain C) against fixture G2 — RK4 driven by the reference's own
    ``env.func`` — and against the NumPy oracle."""
    import lyap_oracle as lo
    g = np.load(os.path.join(golden_dir, "lyap_optomech.npz"))
    tag = "E5"
    p, y0, acts, T, K = g[f"{tag}_p"], g[f"{tag}_y0"], g[f"{tag}_actions"], g[f"{tag}_T"], int(g[f"{tag}_K"])
    func = lo.make_func(2, 4, lambda t, mo, u: sy.optomech_mode_rates(p, t, mo, u), lambda t, mo, u: sy.optomech_A(p, t, mo, u),
                        lambda t, mo, u: sy.optomech_D(p, t, mo, u), p.shape[0])
    y, rows = y0.copy(), [y0[None].copy()]
    for a in range(acts.shape[0]):
        Y = em_port.lyap_rk4_interval(p, acts[a][:, 0], y, K, float(g[f"{tag}_ssz"]), substeps=substeps)
        np.testing.assert_allclose(Y, lo.rk4_interval(func, T[a * K:(a + 1) * K + 1], y, acts[a][:, 0], substeps=substeps),
                                   rtol=1e-12, atol=1e-13)
        rows.append(Y[1:])
        y = Y[-1]
    np.testing.assert_allclose(np.concatenate(rows, 0), g[f"{tag}_States_{key}_reffunc"], rtol=1e-12, atol=1e-13)


def test_c_philox_known_answers_and_polynomial_normals():
    """oracle/philox_port.c restates csrc/philox.cuh operation by operation.  (1) Random123's Philox4x32-10 known-answer
    vectors; (2) the kernel's table-driven Box–Muller (256-bin ln and sin / cos tables + short polynomials, third-order
    sqrt) against the extended-precision libm evaluation of the same stream (oracle/em_oracle.py:philox_normals): the
    accuracy claim of philox.cuh, |dz| <= 5 * 2^-53 * radius, checked without a GPU — and the moments of those normals."""
    assert em_port.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert em_port.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert em_port.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    seed, episode = 0x1234567890ABCDEF, 3
    tau = np.arange(0, 400, dtype=np.uint64).reshape(-1, 1, 1)
    tau[0] = 0xFFFFFFFF
    env = (np.uint64(17) + np.arange(64, dtype=np.uint64)).reshape(1, -1, 1)
    comp = np.arange(8, dtype=np.uint64).reshape(1, 1, -1)
    r0, r1 = eo.philox_normals(seed, episode, tau, env, comp)
    z0, z1 = em_port.philox_normals_poly(seed, episode, tau, env, comp)
    z0, z1 = z0.reshape(r0.shape), z1.reshape(r1.shape)
    radius = np.hypot(r0, r1)
    for z, r in ((z0, r0), (z1, r1)):
        err = np.abs(z - r)
        assert np.max(err) < 4e-15 and np.max(err / radius) < 5 * 2.0 ** -53
        assert np.mean(err > 2 * np.spacing(np.abs(r))) < 0.01          # more than 2 ulp of the value itself: < 1 %
    # radius alone (sin^2 + cos^2 = 1 up to the direction's error): -2 ln u1 and its square root within 2 ulp
    assert np.max(np.abs(np.hypot(z0, z1) - radius) / radius) < 7 * 2.0 ** -53
    z = np.concatenate([z0.ravel(), z1.ravel()])
    m = z.size
    assert abs(z.mean()) < 5 / np.sqrt(m) and abs(z.var() - 1) < 5 * np.sqrt(2 / m) and abs((z ** 4).mean() - 3) < 5 * np.sqrt(96 / m)
    assert abs(np.corrcoef(z0.ravel(), z1.ravel())[0, 1]) < 0.01


def test_polynomial_normals_pass_ks_and_drive_an_ou_process_to_its_stationary_covariance():
    """SURVEY.md §8c G5 on the CPU, with the kernel's own generator (C twin): a Kolmogorov–Smirnov test of the
    polynomial Box–Muller normals against N(0,1), and a 4-dimensional Ornstein–Uhlenbeck ensemble integrated by the
    oracle's Euler–Maruyama with those increments reaches the stationary covariance of the DISCRETE recurrence,
    ``Sigma = M Sigma M^T + dt diag(nz^2)`` (``scipy.linalg.solve_discrete_lyapunov``)."""
    from scipy.linalg import solve_discrete_lyapunov
    from scipy.stats import kstest
    seed, E, n, K, dt = 20251020, 4000, 4, 600, 0.01
    tau = np.arange(K, dtype=np.uint64).reshape(-1, 1, 1)
    env = np.arange(E, dtype=np.uint64).reshape(1, -1, 1)
    comp = np.arange(n, dtype=np.uint64).reshape(1, 1, -1)
    z0, z1 = em_port.philox_normals_poly(seed, 0, tau, env, comp)
    assert kstest(z0[::7], "norm").pvalue > 1e-3 and kstest(z1[::7], "norm").pvalue > 1e-3
    assert kstest(np.concatenate([z0[:200000], z1[:200000]]) ** 2, "chi2", args=(1,)).pvalue > 1e-3
    rng = np.random.default_rng(0)
    G = rng.standard_normal((n, n))
    A = (G - G.T) - 1.0 * np.eye(n)                       # rotation + unit damping: relaxation time 1 = 100 sub-steps
    nz = np.array([0.3, 0.5, 0.2, 0.4])
    Y = eo.em_interval_vec(np.zeros((E, n)), np.broadcast_to(A, (E, n, n)), np.broadcast_to(nz, (E, n)),
                           np.sqrt(dt) * z0.reshape(K, E, n), dt)
    M = np.eye(n) + A * dt
    Sigma = solve_discrete_lyapunov(M, dt * np.diag(nz ** 2))
    S = np.cov(Y[-1].T)
    scale = np.sqrt(np.outer(np.diag(Sigma), np.diag(Sigma)))
    assert np.max(np.abs(S - Sigma) / scale) < 0.1, (S, Sigma)
    assert np.max(np.abs(Y[-1].mean(0)) / np.sqrt(np.diag(Sigma))) < 0.08
