import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_em_bench import bench
for pack, flags, name in ((0, 0, "no packed rows"), (2, 0, "reward column only (direct)"), (2, 128, "reward column only (TMA staged)"), (1, 0, "all 7 columns (direct)"), (1, 128, "all 7 columns (TMA staged)")):
    r = bench(4096, 4, 100, True, True, reps=40, pack=pack, flags=flags)
    print(f"{name:48s} {r['ms']*1e3:7.1f} us")
