import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_em_bench import bench
for flags, name in ((0, "full"), (256, "full + odd-warp phase shift"), (12, "produce only"), (12 + 256, "produce only + phase shift")):
    r = bench(4096, 4, 100, True, True, reps=40, flags=flags)
    print(f"{name:36s} {r['ms']*1e3:7.1f} us")
