"""CPU: the ``'b200'`` entries land in the reference's own registries (only where /root/reference exists)."""
import pytest

import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")


def test_register_into_reference_registries():
    ref_shim.load()
    import quantrl.solvers.context_manager as scm
    import quantrl.backends.context_manager as bcm
    import torch
    from quantrl_b200.plugin import register
    from quantrl_b200.solvers import B200IVPSolver
    if torch.cuda.is_available():
        backend, solver = register()
        assert bcm.get_backend_instance("b200") is backend
    else:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            register()
        backend, solver = register(require_cuda=False)
        assert backend is None and "b200" not in bcm.BACKEND_INSTANCES
    assert solver is B200IVPSolver and scm.get_IVP_solver("b200") is B200IVPSolver
    # the solver honours the reference's BaseIVPSolver contract (solvers/base.py:60-67, 115-116)
    assert {"rk4", "dopri5", "bosh3", "RK23", "tsit5", "adaptive_heun", "heun", "fehlberg2", "dop853", "DOP853", "dopri8",
            "euler", "midpoint"} == set(B200IVPSolver.solver_methods)
    for key in ("method", "atol", "rtol", "is_stiff", "step_interval", "complex"):
        assert key in B200IVPSolver.default_solver_params


def test_backend_surface_matches_reference_base_backend():
    """Every public method of the reference's BaseBackend exists on B200Backend with the same parameter names."""
    import inspect
    ref_shim.load()
    from quantrl.backends.base import BaseBackend
    from quantrl_b200.backends import B200Backend
    for name, fn in inspect.getmembers(BaseBackend, predicate=inspect.isfunction):
        if name.startswith("_"):
            continue
        assert hasattr(B200Backend, name), f"missing backend method {name}"
        ref_params = [p for p in inspect.signature(fn).parameters if p != "self"]
        got_params = [p for p in inspect.signature(getattr(B200Backend, name)).parameters if p != "self"]
        assert ref_params == got_params, (name, ref_params, got_params)
