"""GPU: the reference's OWN env class on the ``'b200'`` backend, on a B200.  The reference package is
imported, unmodified, from the staged copy ``oracle/_ref`` (oracle/stage_ref.py; /root/reference does not exist on the GPU
box) through the import shims of oracle/ref_shim.py; ``quantrl_b200.register()`` puts the backend instance and the solver
class into the reference's registries (backends/context_manager.py:38-39, solvers/context_manager.py:31-32)."""
import numpy as np
import pytest
import torch

import ref_shim
from ref_on_b200 import run_and_check

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (ref_shim.available() or ref_shim.staged_available()),
                                 reason="the staged reference (oracle/_ref) is absent: run oracle/stage_ref.py where /root/reference exists")]


def test_reference_env_class_runs_on_the_b200_backend_on_cuda(native, cuda, golden_dir):
    quantrl = ref_shim.import_reference()
    import quantrl.backends.context_manager as bcm
    import quantrl_b200
    backend, solver = quantrl_b200.register()
    assert bcm.get_backend_instance("b200") is backend and backend.device.type == "cuda"
    env = run_and_check(quantrl, golden_dir, "cuda", n_steps=3)
    assert env.States.is_cuda and env.Observations.is_cuda
    # the same system through the fused device functor of quantrl_b200's own env class: same trajectory
    from quantrl_b200.envs import LinearizedHOVecEnv
    import os
    g = np.load(os.path.join(golden_dir, "lyap_optomech.npz"))

    class Fused(LinearizedHOVecEnv):
        def reset_states(self):
            return torch.as_tensor(g["E5_y0"], device="cuda")

        def get_reward(self):
            O = self.Observations
            return -(O[..., 4] + O[..., 9] + O[..., 14] + O[..., 19])

    K, S, ssz = int(g["E5_K"]), int(g["E5_S"]), float(g["E5_ssz"])
    fused = Fused(name="optomech", desc="fused", params={}, num_modes=2, num_quads=4, t_norm_max=S * ssz + ssz / 2, t_norm_ssz=ssz,
                  t_norm_mul=1.0, n_envs=5, n_observations=20, n_properties=0, n_actions=1, action_maximums=[1.0], action_interval=K,
                  data_idxs=[0, 1, 2, 22], cache_all_data=False, dir_prefix="/tmp/qrl_test", ode_method="rk4", system="optomech",
                  system_params=g["E5_p"])
    fused.reset()
    for a in range(3):
        fused.step(g["E5_actions"][a])
    torch.testing.assert_close(fused.States, env.States, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(fused.data[:, :3 * K + 1], env.data[:, :3 * K + 1], rtol=1e-9, atol=1e-10)
    fused.close()


def test_b200_backend_ops_on_cuda_match_numpy(native, cuda):
    """SURVEY §8 a9 on the device: the op surface the reference's envs and solvers call, on CUDA tensors, against NumPy."""
    from quantrl_b200.backends import get_backend_instance
    b = get_backend_instance("b200", precision="double")
    assert b.device.type == "cuda"
    rng = np.random.default_rng(0)
    X, Y = rng.standard_normal((5, 4, 4)), rng.standard_normal((5, 4, 4))
    v, w = rng.standard_normal(7), rng.standard_normal(7)
    Z = X + 1j * Y
    t = lambda a, dtype="real": b.convert_to_typed(tensor=a, dtype=dtype)  # noqa: E731
    same = lambda got, ref, tol=1e-13: np.testing.assert_allclose(b.convert_to_numpy(got), ref, rtol=tol, atol=0)  # noqa: E731
    assert t(X).is_cuda and t(X).dtype == torch.float64 and t(Z, "complex").dtype == torch.complex128
    same(b.add(t(X), t(Y), out=None), X + Y)
    out = b.empty((5, 4, 4), "real")
    assert b.matmul(t(X), t(Y), out=out) is out and out.is_cuda
    same(out, X @ Y)
    same(b.dot(t(v), t(w), out=None), v @ w)
    same(b.norm(t(X), axis=2), np.linalg.norm(X, axis=2))
    same(b.transpose(t(X), 1, 2), np.swapaxes(X, 1, 2))
    same(b.repeat(t(X[:1]), repeats=3, axis=0), np.repeat(X[:1], 3, 0))
    same(b.concatenate((t(X), t(Y)), axis=1, out=None), np.concatenate((X, Y), 1))
    same(b.stack((t(X), t(Y)), axis=0, out=None), np.stack((X, Y), 0))
    same(b.update(t(X.copy()), 2, t(Y[2])), np.concatenate((X[:2], Y[2:3], X[3:])))
    same(b.reshape(t(X), (5, 16)), X.reshape(5, 16))
    same(b.real(t(Z, "complex")), X)
    same(b.imag(t(Z, "complex")), Y)
    same(b.conj(t(Z, "complex")), Z.conj())
    same(b.sqrt(t(np.abs(X))), np.sqrt(np.abs(X)))
    same(b.sum(t(X), axis=1), X.sum(1))
    same(b.cumsum(t(X), axis=2), X.cumsum(2))
    same(b.min(t(X)), X.min())
    same(b.max(t(X)), X.max())
    assert int(b.argmax(t(X))) == int(X.argmax())
    for name, args, ref in (("zeros", ((3, 2),), np.zeros((3, 2))), ("ones", ((3, 2),), np.ones((3, 2))), ("eye", (3,), np.eye(3)),
                            ("arange", (0.0, 1.0, 0.25), np.arange(0.0, 1.0, 0.25)), ("linspace", (0.0, 1.0, 11), np.linspace(0.0, 1.0, 11))):
        got = getattr(b, name)(*args, dtype="real")
        assert got.is_cuda
        same(got, ref)
    f = lambda i, Yv, args: Yv + (i + 1) * args[0]  # noqa: E731
    same(b.iterate_i(f, 4, b.zeros((3,), "real"), (2.0,)), np.full(3, 20.0))
    gen = b.generator(seed=11)
    z = b.normal(gen, (20000,), 0.0, 1.0, "real")
    assert z.is_cuda and z.dtype == torch.float64 and abs(float(z.mean())) < 0.03 and abs(float(z.std()) - 1.0) < 0.03
    u = b.uniform(gen, (2000, 2), low=-2.0, high=3.0, dtype="real")
    assert float(u.min()) >= -2.0 and float(u.max()) < 3.0
    k = b.integers(gen, (1000,), low=3, high=9, dtype="integer")
    assert k.dtype == torch.int64 and int(k.min()) >= 3 and int(k.max()) <= 8
    assert isinstance(b.convert_to_numpy(z, "real"), np.ndarray)
