set -x
T=${1:-r1m}
ncu --set full --clock-control none --import-source on -k regex:rk4_kernel -s 3 -c 1 -f -o gpurun_out/prof_rk4_thread_$T python tools/prof_ode.py 262144 rk4 1 > gpurun_out/ncu_rk4_thread_$T.log 2>&1; tail -2 gpurun_out/ncu_rk4_thread_$T.log
ncu --set full --clock-control none --import-source on -k regex:dopri5_kernel -s 3 -c 1 -f -o gpurun_out/prof_dopri5_$T python tools/prof_ode.py 65536 dopri5 0 > gpurun_out/ncu_dopri5_$T.log 2>&1; tail -2 gpurun_out/ncu_dopri5_$T.log
