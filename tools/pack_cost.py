import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_em_bench import bench
for pack, rot, name in ((0, False, "no packed rows, outputs reused (L2 resident)"), (0, True, "no packed rows, outputs to fresh memory"), (1, False, "packed rows (7 cols), S/O reused"), (1, True, "packed rows (7 cols), all fresh")):
    r = bench(4096, 4, 100, True, True, reps=40, pack=pack, rotate=rot)
    print(f"{name:48s} {r['ms']*1e3:7.1f} us")
