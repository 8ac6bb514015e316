"""CPU: the reference's OWN ``LinearizedHOVecEnv`` class runs on ``backend_library='b200'`` after ``register()`` — host
logic only (the backend's device is pointed at the CPU for this test: its ops are device-agnostic torch calls and
``B200IVPSolver`` takes the stage-callback mode for a reference env, which launches no kernel).  The GPU twin of this
test (tests/test_reference_on_b200_gpu.py) runs the same thing on cuda from the staged copy of the reference."""
import pytest
import torch

import ref_shim
from ref_on_b200 import run_and_check

pytestmark = pytest.mark.skipif(not (ref_shim.available() or ref_shim.staged_available()), reason="reference not present")


def test_reference_env_class_runs_on_the_b200_backend(monkeypatch, golden_dir):
    quantrl = ref_shim.import_reference()
    import quantrl.backends.context_manager as bcm
    import quantrl_b200
    from quantrl_b200.backends import B200Backend
    if not torch.cuda.is_available():
        _, solver = quantrl_b200.register(require_cuda=False)      # the solver class; no backend without a CUDA device
        monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
        bcm.BACKEND_INSTANCES["b200"] = B200Backend(precision="double", device="cpu")
        device = "cpu"
    else:
        quantrl_b200.register()
        device = "cuda"
    try:
        run_and_check(quantrl, golden_dir, device)
    finally:
        bcm.BACKEND_INSTANCES.pop("b200", None)
