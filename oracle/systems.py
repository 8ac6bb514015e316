"""TEST INFRASTRUCTURE ONLY — NumPy statements of the canonical system hooks.

The reference ships no concrete physical system: ``get_mode_rates``/``get_A``/``get_D`` (quantrl/envs/
deterministic.py:748-818) and ``get_A``/``get_noise_prefixes`` (quantrl/envs/stochastic.py:244-284) are user hooks
that raise ``NotImplementedError``.  The systems below are the "user code" of BASELINE.json's configs.  Each is
written twice in this repository: here (NumPy, batched over envs, the form the reference's hooks take) and as a
CUDA device functor in ``quantrl_b200/csrc/systems.cuh``.  Parity is engine-vs-engine on identical formulas.

Parameter rows (one row per env, FP64) use the SAME column order as the CUDA functors:

``optomech``  (m=2 complex modes a,b; q=4 quadratures dX,dY,dQ,dP; 1 action u)
    p = [kappa, gamma, omega_m, Delta0, g0, n_th, A_l]
    da/dt = -(kappa/2 - i Delta0) a + i g0 a (b + b*) + A_l (1 + u)
    db/dt = -(gamma/2 + i omega_m) b + i g0 |a|^2
    Delta = Delta0 + 2 g0 Re b,  G = 2 g0 a = G_R + i G_I
    A = [[-kappa/2, -Delta, -G_I, 0], [Delta, -kappa/2, G_R, 0], [0, 0, -gamma/2, omega_m],
         [G_R, G_I, -omega_m, -gamma/2]]
    D = diag(kappa/2, kappa/2, gamma (n_th + 1/2), gamma (n_th + 1/2))

``kerr_chain``  (m=N complex modes; q=2N quadratures (X_0,Y_0,...,X_{N-1},Y_{N-1}); 1 action u)
    p = [kappa, chi, J, n_th, eps, Delta_0 .. Delta_{N-1}]
    da_j/dt = -(kappa/2 + i Delta_j) a_j - i chi |a_j|^2 a_j - i J (a_{j-1} + a_{j+1}) + [j == 0] eps (1 + u)
    Delta'_j = Delta_j + 2 chi |a_j|^2,  G_j = chi a_j^2 = G_R + i G_I
    A_jj = [[-kappa/2 + G_I, Delta' - G_R], [-Delta' - G_R, -kappa/2 - G_I]],  A_jk = [[0, J], [-J, 0]] (|j-k| = 1)
    D = kappa (n_th + 1/2) I

``affine``  (stochastic LinearEnv drift; n states; n_act actions)
    A(u) = A0 + sum_a u_a A_a   (time-invariant within an action interval), noise prefixes = constant vector
"""
import numpy as np

OPTOMECH_NPARAMS = 7


def optomech_default_params(E, seed=2, jitter=True):
    """BASELINE config 1/2 parameters: one row per env."""
    p = np.empty((E, OPTOMECH_NPARAMS))
    p[:] = [0.1, 1e-3, 1.0, -1.0, 1e-2, 10.0, 5.0]
    if jitter and E > 1:
        p[:, 3] = np.random.default_rng(seed).uniform(-1.2, -0.8, size=E)
    return p


def optomech_y0(p):
    """y = [Re a, Re b, Im a, Im b, vec(V)] with V(0) = diag(1/2, 1/2, n_th+1/2, n_th+1/2)."""
    E = p.shape[0]
    y = np.zeros((E, 20))
    V = np.zeros((E, 4, 4))
    V[:, 0, 0] = V[:, 1, 1] = 0.5
    V[:, 2, 2] = V[:, 3, 3] = p[:, 5] + 0.5
    y[:, 4:] = V.reshape(E, 16)
    return y


def optomech_mode_rates(p, t, modes, u):
    """modes (E,2) complex, u (E,) -> (E,2) complex."""
    kappa, gamma, om, D0, g0, nth, Al = (p[:, i] for i in range(7))
    a, b = modes[:, 0], modes[:, 1]
    da = -(kappa / 2 - 1j * D0) * a + 1j * g0 * a * (b + np.conj(b)) + Al * (1.0 + u)
    db = -(gamma / 2 + 1j * om) * b + 1j * g0 * (a.real * a.real + a.imag * a.imag)
    return np.stack([da, db], axis=1)


def optomech_A(p, t, modes, u):
    kappa, gamma, om, D0, g0, nth, Al = (p[:, i] for i in range(7))
    a, b = modes[:, 0], modes[:, 1]
    Delta = D0 + 2 * g0 * b.real
    GR, GI = 2 * g0 * a.real, 2 * g0 * a.imag
    A = np.zeros((p.shape[0], 4, 4))
    A[:, 0, 0] = -kappa / 2
    A[:, 0, 1] = -Delta
    A[:, 0, 2] = -GI
    A[:, 1, 0] = Delta
    A[:, 1, 1] = -kappa / 2
    A[:, 1, 2] = GR
    A[:, 2, 2] = -gamma / 2
    A[:, 2, 3] = om
    A[:, 3, 0] = GR
    A[:, 3, 1] = GI
    A[:, 3, 2] = -om
    A[:, 3, 3] = -gamma / 2
    return A


def optomech_D(p, t, modes, u):
    kappa, gamma, nth = p[:, 0], p[:, 1], p[:, 5]
    D = np.zeros((p.shape[0], 4, 4))
    D[:, 0, 0] = D[:, 1, 1] = kappa / 2
    D[:, 2, 2] = D[:, 3, 3] = gamma * (nth + 0.5)
    return D


def kerr_nparams(N):
    return 5 + N


def kerr_default_params(E, N, seed=4):
    rng = np.random.default_rng(seed)
    p = np.empty((E, kerr_nparams(N)))
    p[:, 0] = 0.2                          # kappa
    p[:, 1] = 1e-3                         # chi
    p[:, 2] = rng.uniform(0.3, 0.5, E)     # J
    p[:, 3] = 0.5                          # n_th
    p[:, 4] = 2.0                          # eps
    p[:, 5:] = rng.uniform(-1.0, 1.0, (E, N))
    return p


def kerr_y0(p, N):
    E = p.shape[0]
    q = 2 * N
    y = np.zeros((E, 2 * N + q * q))
    V = np.zeros((E, q, q))
    idx = np.arange(q)
    V[:, idx, idx] = (p[:, 3] + 0.5)[:, None]
    y[:, 2 * N:] = V.reshape(E, q * q)
    return y


def kerr_mode_rates(p, t, modes, u):
    N = modes.shape[1]
    kappa, chi, J, nth, eps = (p[:, i] for i in range(5))
    Dl = p[:, 5:5 + N]
    out = np.empty_like(modes)
    for j in range(N):
        a = modes[:, j]
        n2 = a.real * a.real + a.imag * a.imag
        nb = np.zeros_like(a)
        if j > 0:
            nb = nb + modes[:, j - 1]
        if j < N - 1:
            nb = nb + modes[:, j + 1]
        r = -(kappa / 2 + 1j * Dl[:, j]) * a - 1j * chi * n2 * a - 1j * J * nb
        if j == 0:
            r = r + eps * (1.0 + u)
        out[:, j] = r
    return out


def kerr_A(p, t, modes, u):
    N = modes.shape[1]
    q = 2 * N
    kappa, chi, J = p[:, 0], p[:, 1], p[:, 2]
    Dl = p[:, 5:5 + N]
    A = np.zeros((p.shape[0], q, q))
    for j in range(N):
        a = modes[:, j]
        n2 = a.real * a.real + a.imag * a.imag
        Dp = Dl[:, j] + 2 * chi * n2
        GR = chi * (a.real * a.real - a.imag * a.imag)
        GI = chi * (2 * a.real * a.imag)
        A[:, 2 * j, 2 * j] = -kappa / 2 + GI
        A[:, 2 * j, 2 * j + 1] = Dp - GR
        A[:, 2 * j + 1, 2 * j] = -Dp - GR
        A[:, 2 * j + 1, 2 * j + 1] = -kappa / 2 - GI
        for k in (j - 1, j + 1):
            if 0 <= k < N:
                A[:, 2 * j, 2 * k + 1] = J
                A[:, 2 * j + 1, 2 * k] = -J
    return A


def kerr_D(p, t, modes, u):
    N = modes.shape[1]
    q = 2 * N
    D = np.zeros((p.shape[0], q, q))
    idx = np.arange(q)
    D[:, idx, idx] = (p[:, 0] * (p[:, 3] + 0.5))[:, None]
    return D


def affine_matrices(n, n_act=1, seed=0):
    """SURVEY.md §8d config 3/4 drift: A0 = J - 0.05 I (J antisymmetric, 0.5 N(0,1)), A_a = 0.1 N(0,1)."""
    rng = np.random.default_rng(seed)
    G = 0.5 * rng.standard_normal((n, n))
    A0 = (G - G.T) / 2.0 * 2.0 ** 0.5 - 0.05 * np.eye(n)
    Au = 0.1 * rng.standard_normal((n_act, n, n))
    return A0, Au


def affine_A(A0, Au, u):
    """u (E, n_act) -> (E, n, n);  A0 (n,n) or (E,n,n);  Au (n_act,n,n)."""
    return A0 + np.einsum("ea,aij->eij", u, Au)
