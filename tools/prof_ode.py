"""Profile driver: a few launches of qrl_ode_step (used under ncu).  argv: E [method] [flags] [system q m] [traj|notraj] [reward_kind]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_ode_bench import bench
E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
method = sys.argv[2] if len(sys.argv) > 2 else "rk4"
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
system = sys.argv[4] if len(sys.argv) > 4 else "optomech"
q = int(sys.argv[5]) if len(sys.argv) > 5 else 4
m = int(sys.argv[6]) if len(sys.argv) > 6 else 2
traj = (sys.argv[7] != "notraj") if len(sys.argv) > 7 else True
reward = int(sys.argv[8]) if len(sys.argv) > 8 else 0
print(json.dumps(bench(system, E, m, q, method=method, flags=flags, reps=5, traj=traj, reward_kind=reward)))
