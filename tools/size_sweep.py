import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_em_bench import bench
for n in (4, 8):
    for E in (1024, 4096, 8192, 16384, 32768, 65536, 262144):
        t = bench(E, n, 100, True, True, reps=10, flags=0)
        g = bench(E, n, 100, True, True, reps=10, flags=1)
        print(f"n={n} E={E:7d} tile {t['ms']*1e3:8.1f} us {t['substeps_per_s']}  generic {g['ms']*1e3:8.1f} us {g['substeps_per_s']}")
