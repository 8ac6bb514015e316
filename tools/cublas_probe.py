import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
which = sys.argv[1]
if which == "lib_first":
    from quantrl_b200 import _native as N
    N.lib(); print("fp64 peak", N.measure_fp64_peak(1024))
a = torch.randn(64, 1, dtype=torch.float64, device="cuda"); b = torch.randn(1, 16, dtype=torch.float64, device="cuda")
try:
    print(which, "f64 matmul ok", float((a @ b).sum()))
except Exception as e:
    print(which, "f64 matmul FAILED", str(e)[:200])
a = torch.randn(64, 64, device="cuda")
try:
    print(which, "f32 matmul ok", float((a @ a).sum()))
except Exception as e:
    print(which, "f32 matmul FAILED", str(e)[:200])
