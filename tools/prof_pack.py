import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_em_bench import bench
pack, flags = int(sys.argv[1]), int(sys.argv[2])
print(json.dumps(bench(4096, 4, 100, True, True, reps=10, pack=pack, flags=flags, graph=False)))
