set -x
T=${1:-r1l}
for E in 1024 4096 16384 65536 262144; do python tools/prof_ode.py $E rk4 0; done > gpurun_out/ode_quick_$T.log 2>&1
python tools/prof_ode.py 65536 rk4 0 optomech 4 2 notraj >> gpurun_out/ode_quick_$T.log 2>&1
python tools/prof_ode.py 65536 dopri5 0 >> gpurun_out/ode_quick_$T.log 2>&1
python tools/prof_ode.py 16384 rk4 0 kerr_chain 8 4 >> gpurun_out/ode_quick_$T.log 2>&1
cat gpurun_out/ode_quick_$T.log
ncu --set full --clock-control none --import-source on -k regex:rk4_kernel -s 3 -c 1 -f -o gpurun_out/prof_rk4_thread_$T python tools/prof_ode.py 65536 rk4 0 > gpurun_out/ncu_rk4_thread_$T.log 2>&1; tail -2 gpurun_out/ncu_rk4_thread_$T.log
ncu --set full --clock-control none --import-source on -k regex:rk4_lanes -s 3 -c 1 -f -o gpurun_out/prof_rk4_lanes_$T python tools/prof_ode.py 1024 rk4 0 > gpurun_out/ncu_rk4_lanes_$T.log 2>&1; tail -2 gpurun_out/ncu_rk4_lanes_$T.log
