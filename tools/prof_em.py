"""Profile driver: a few launches of qrl_em_step at config-3 size (used under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_em_bench import bench
import json
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
philox = (sys.argv[3] != "fed") if len(sys.argv) > 3 else True
traj = (sys.argv[4] != "notraj") if len(sys.argv) > 4 else True
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
print(json.dumps(bench(E, n, 100, philox, traj, reps=10, flags=flags)))
