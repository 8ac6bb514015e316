"""Registration of the ``'b200'`` backend and IVP solver in the reference's own registries.

The reference resolves ``backend_library`` through two module-level dicts that are consulted *before* its substring
dispatch: ``quantrl.backends.context_manager.BACKEND_INSTANCES`` (quantrl/backends/context_manager.py:14, 38-39) and
``quantrl.solvers.context_manager.IVP_SOLVERS`` (quantrl/solvers/context_manager.py:14, 31-32).  ``register()`` puts
the B200 backend instance and solver class there, so that

    import quantrl_b200; quantrl_b200.register()
    class MyEnv(quantrl.envs.deterministic.LinearizedHOVecEnv):
        backend_libraries = ['torch', 'jax', 'numpy', 'b200']      # the class-level allow-list (deterministic.py:515)
    MyEnv(..., backend_library='b200', ode_method='rk4')

runs the reference's own env class on this backend (stage-callback mode of ``B200IVPSolver``: the reference's Python
``func`` evaluated with torch on the device).  The fused single-launch path needs the env classes of
``quantrl_b200.envs`` (they declare device functors / batched hooks).
"""


def register(precision: str = "double", require_cuda: bool = True):
    """Insert the backend instance and the solver class into the reference's registries.  Returns
    ``(backend_or_None, solver_class)``.  Raises ``ImportError`` if ``quantrl`` is not importable."""
    import quantrl.backends.context_manager as bcm   # noqa: F401  (ImportError propagates: nothing to register into)
    import quantrl.solvers.context_manager as scm

    from .solvers import B200IVPSolver
    scm.IVP_SOLVERS["b200"] = B200IVPSolver
    backend = None
    import torch
    if torch.cuda.is_available():
        from .backends import get_backend_instance
        backend = get_backend_instance("b200", precision=precision)
        bcm.BACKEND_INSTANCES["b200"] = backend
    elif require_cuda:
        raise RuntimeError("quantrl_b200.register(): no CUDA device — the 'b200' backend has no CPU fallback")
    return backend, B200IVPSolver
