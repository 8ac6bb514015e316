# One GPU call of round-2 evidence (usage: bash tools/run_round2_profile.sh r2k): bench line, ncu launch list, full capture of
# the dominant kernel inside bench.py (in place) and over the ring, the q = 8 row-lanes kernel, the n = 8 tiled kernel.
set -x
T=${1:-r2k}
python bench.py --steps 1000 --warmup 20 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err || tail -5 gpurun_out/bench_$T.err
head -c 400 gpurun_out/bench_$T.json; echo
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --only-headline > gpurun_out/plain_$T.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --only-headline > gpurun_out/ncu_ll_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_tile3 -s 60 -c 2 -f -o gpurun_out/prof_tile_$T python bench.py --steps 40 --warmup 5 --no-cpu-baseline --only-headline > gpurun_out/ncu_full_$T.log 2>&1; tail -2 gpurun_out/ncu_full_$T.log
python tools/prof_ring.py 40 && ncu --set full --clock-control none -k regex:em_tile3 -s 20 -c 2 -f -o gpurun_out/prof_ring_$T python tools/prof_ring.py 40 > gpurun_out/ncu_ring_$T.log 2>&1; tail -2 gpurun_out/ncu_ring_$T.log
python tools/prof_em.py 8192 8 && ncu --set full --clock-control none --import-source on -k regex:em_tile3 -s 5 -c 1 -f -o gpurun_out/prof_tile8_$T python tools/prof_em.py 8192 8 > gpurun_out/ncu_tile8_$T.log 2>&1; tail -2 gpurun_out/ncu_tile8_$T.log
python tools/prof_ode.py 16384 rk4 0 kerr_chain 8 4 && ncu --set full --clock-control none --import-source on -k regex:rk4_rows -s 2 -c 1 -f -o gpurun_out/prof_rows_$T python tools/prof_ode.py 16384 rk4 0 kerr_chain 8 4 > gpurun_out/ncu_rows_$T.log 2>&1; tail -2 gpurun_out/ncu_rows_$T.log
python tools/prof_ode.py 16384 dopri5 0 kerr_chain 8 4 && ncu --set full --clock-control none --import-source on -k regex:dopri5_rows -s 2 -c 1 -f -o gpurun_out/prof_dprows_$T python tools/prof_ode.py 16384 dopri5 0 kerr_chain 8 4 > gpurun_out/ncu_dprows_$T.log 2>&1; tail -2 gpurun_out/ncu_dprows_$T.log
{ for E in 1024 4096 16384; do python tools/prof_ode.py $E dopri5 0 kerr_chain 8 4; python tools/prof_ode.py $E dopri5 1 kerr_chain 8 4; done
  for E in 1024 4096 16384 65536; do python tools/prof_ode.py $E rk4 0 kerr_chain 8 4; python tools/prof_ode.py $E rk4 1 kerr_chain 8 4; done
  for E in 1024 65536 262144; do python tools/prof_ode.py $E rk4 0; done
  for E in 1024 65536; do python tools/prof_ode.py $E dopri5 0; done; } > gpurun_out/ode_quick_$T.log 2>&1
cat gpurun_out/ode_quick_$T.log | cut -c1-220
python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown_$T.log 2>&1; cat gpurun_out/e2e_breakdown_$T.log
