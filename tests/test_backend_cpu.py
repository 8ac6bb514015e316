"""CPU: behaviour of ``B200Backend``'s plumbing ops against the reference's own ``NumPyBackend`` on
the same inputs (only where /root/reference exists).  The backend refuses to exist without a CUDA device; for this
test its device is pointed at the CPU — the ops are device-agnostic torch calls, the product always runs them on cuda."""
import numpy as np
import pytest
import torch

import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")


@pytest.fixture()
def backends(monkeypatch):
    ref_shim.load()
    from quantrl.backends.numpy import NumPyBackend
    from quantrl_b200.backends import B200Backend
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            B200Backend()
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    return B200Backend(precision="double", device="cpu"), NumPyBackend(precision="double")


def _same(b, got, ref, **kw):
    np.testing.assert_allclose(b.convert_to_numpy(got), np.asarray(ref), **({"rtol": 1e-14, "atol": 0} | kw))


def test_elementwise_and_linear_algebra_ops_match_the_numpy_backend(backends):
    b, r = backends
    rng = np.random.default_rng(0)
    X, Y = rng.standard_normal((5, 4, 4)), rng.standard_normal((5, 4, 4))
    v, w = rng.standard_normal(7), rng.standard_normal(7)
    Z = X + 1j * Y
    t = lambda a, dtype="real": b.convert_to_typed(tensor=a, dtype=dtype)  # noqa: E731
    _same(b, b.add(t(X), t(Y), out=None), r.add(X, Y, out=None))
    out_t, out_n = b.empty((5, 4, 4), "real"), r.empty((5, 4, 4), "real")
    assert b.matmul(t(X), t(Y), out=out_t) is out_t                          # ``out=`` buffers are honoured
    _same(b, out_t, r.matmul(X, Y, out=out_n), rtol=1e-13)
    _same(b, b.dot(t(v), t(w), out=None), r.dot(v, w, out=None), rtol=1e-13)
    _same(b, b.norm(t(X), axis=2), r.norm(X, axis=2), rtol=1e-13)
    _same(b, b.transpose(t(X), 1, 2), r.transpose(X, 1, 2))
    _same(b, b.transpose(t(X[0])), r.transpose(X[0]))
    _same(b, b.repeat(t(X[:1]), repeats=3, axis=0), r.repeat(X[:1], repeats=3, axis=0))
    _same(b, b.concatenate((t(X), t(Y)), axis=1, out=None), r.concatenate((X, Y), axis=1, out=None))
    _same(b, b.stack((t(X), t(Y)), axis=0, out=None), r.stack((X, Y), axis=0, out=None))
    _same(b, b.update(t(X.copy()), 2, t(Y[2])), r.update(X.copy(), 2, Y[2]))
    _same(b, b.reshape(t(X), (5, 16)), r.reshape(X, (5, 16)))
    _same(b, b.flatten(t(X)), r.flatten(X))
    _same(b, b.real(t(Z, "complex")), r.real(Z))
    _same(b, b.imag(t(Z, "complex")), r.imag(Z))
    _same(b, b.conj(t(Z, "complex")), r.conj(Z))
    _same(b, b.sqrt(t(np.abs(X))), r.sqrt(np.abs(X)))
    _same(b, b.sum(t(X), axis=None), r.sum(X, axis=None), rtol=1e-13)      # ``axis`` is required by the reference
    _same(b, b.sum(t(X), axis=1), r.sum(X, axis=1), rtol=1e-13)
    _same(b, b.cumsum(t(X), axis=2), r.cumsum(X, axis=2), rtol=1e-13)
    _same(b, b.min(t(X)), r.min(X))
    _same(b, b.max(t(X)), r.max(X))
    assert int(b.argmax(t(X))) == int(r.argmax(X))
    assert b.shape(t(X)) == tuple(r.shape(X)) == (5, 4, 4)


def test_constructors_and_dtypes_match_the_numpy_backend(backends):
    b, r = backends
    for name, args in (("zeros", ((3, 2),)), ("ones", ((3, 2),)), ("eye", (3,)), ("arange", (0.0, 1.0, 0.25)),
                       ("linspace", (0.0, 1.0, 11))):
        for dtype in ("real", "complex") if name in ("zeros", "ones", "eye") else ("real",):
            got, ref = getattr(b, name)(*args, dtype=dtype), getattr(r, name)(*args, dtype=dtype)
            _same(b, got, ref)
            assert b.convert_to_numpy(got).dtype == np.asarray(ref).dtype
    _same(b, b.diag([1.0, 2.0, 3.0], dtype="real"), r.diag([1.0, 2.0, 3.0], dtype="real"))
    assert b.empty((2, 3), "real").dtype == torch.float64 and b.empty((2,), "integer").dtype == torch.int64
    assert b.convert_to_typed([1, 2, 3], "real").dtype == torch.float64
    assert b.convert_to_numpy(torch.ones(3), "real").dtype == np.float64
    lst = b.convert_to_numpy([torch.ones(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)], "real")
    assert lst.shape == (2, 2)
    x = b.ones((2, 2), "real")
    assert b.is_typed(x, "real") and not b.is_typed(x, "complex") and not b.is_typed(np.ones(2))
    assert b.convert_to_typed(x, "real") is x                                  # no copy when already typed


def test_control_flow_helpers_match_the_numpy_backend(backends):
    b, r = backends
    f = lambda i, Y, args: Y + (i + 1) * args[0]  # noqa: E731
    _same(b, b.iterate_i(f, 4, b.zeros((3,), "real"), (2.0,)), r.iterate_i(f, 4, r.zeros((3,), "real"), (2.0,)))
    for cond in (True, False):
        assert b.if_else(cond, lambda a: a + 1, lambda a: a - 1, 5) == r.if_else(cond, lambda a: a + 1, lambda a: a - 1, 5)
    for name in ("transpose", "repeat", "add", "matmul", "dot", "concatenate", "stack", "update"):
        assert callable(getattr(b, "jit_" + name))                             # aliases of backends/base.py:577-607


def test_generator_protocol_follows_the_reference(backends):
    """One SeedSequence per backend instance, created from ``seed`` on the FIRST ``generator()`` call; every later call
    spawns the next child regardless of its argument (backends/base.py:101-120, numpy.py:54-59).  The streams themselves
    differ by design (torch Philox vs PCG64): shapes, dtypes, ranges and moments are compared."""
    b, r = backends
    g1, g2 = b.generator(seed=7), b.generator(seed=7)
    a1, a2 = b.normal(g1, (4000,), 0.0, 1.0, "real"), b.normal(g2, (4000,), 0.0, 1.0, "real")
    assert not torch.equal(a1, a2)                       # second child, although the same seed was passed
    assert r.generator(seed=7) is not None and r.seed_sequence.n_children_spawned == 1
    assert b.seed_sequence.n_children_spawned == 2 and b.seed_sequence.entropy == r.seed_sequence.entropy
    assert a1.dtype == torch.float64 and abs(float(a1.mean())) < 0.08 and abs(float(a1.std()) - 1.0) < 0.08
    u = b.uniform(g1, (4000, 2), low=-2.0, high=3.0, dtype="real")
    assert tuple(u.shape) == (4000, 2) and float(u.min()) >= -2.0 and float(u.max()) < 3.0 and abs(float(u.mean()) - 0.5) < 0.1
    k = b.integers(g1, (1000,), low=3, high=9, dtype="integer")
    assert k.dtype == torch.int64 and int(k.min()) >= 3 and int(k.max()) <= 8
    # a fresh instance with the same seed reproduces the first child's stream
    from quantrl_b200.backends import B200Backend
    b2 = B200Backend(precision="double", device="cpu")
    assert torch.equal(b2.normal(b2.generator(seed=7), (4000,), 0.0, 1.0, "real"), a1)
