# One GPU call of a round's evidence: GPU test suite, bench line, ncu launch list + full capture of the dominant kernel,
# and the deterministic-path (qrl_ode_step) timings and captures.  usage: bash tools/run_round_profile.sh r1n
set -x
T=${1:-r1n}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 1000 --warmup 20 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err || tail -5 gpurun_out/bench_$T.err
cat gpurun_out/bench_$T.json | head -c 1500
python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/plain_$T.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_ll_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_tile_kernel -s 60 -c 2 -f -o gpurun_out/prof_tile_$T python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full_$T.log 2>&1; tail -2 gpurun_out/ncu_full_$T.log
# deterministic path: timings (CUDA-graph replay, pure device time), then one full capture per kernel
{ for E in 1024 2048 4096 16384 65536 262144; do python tools/prof_ode.py $E rk4 0; done
  python tools/prof_ode.py 1024 rk4 6; python tools/prof_ode.py 1024 rk4 1
  for E in 64 1024 8192 65536; do python tools/prof_ode.py $E dopri5 0; done
  python tools/prof_ode.py 1024 dopri5 1; python tools/prof_ode.py 65536 dopri5 2
  python tools/prof_ode.py 16384 rk4 0 kerr_chain 8 4; } > gpurun_out/ode_quick_$T.log 2>&1
cat gpurun_out/ode_quick_$T.log
ncu --set full --clock-control none --import-source on -k regex:rk4_kernel -s 3 -c 1 -f -o gpurun_out/prof_rk4_thread_$T python tools/prof_ode.py 262144 rk4 1 > gpurun_out/ncu_ode_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rk4_split -s 3 -c 1 -f -o gpurun_out/prof_rk4_split_$T python tools/prof_ode.py 1024 rk4 0 >> gpurun_out/ncu_ode_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dopri5_lanes -s 3 -c 1 -f -o gpurun_out/prof_dopri5_lanes_$T python tools/prof_ode.py 1024 dopri5 0 >> gpurun_out/ncu_ode_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dopri5_kernel -s 3 -c 1 -f -o gpurun_out/prof_dopri5_thread_$T python tools/prof_ode.py 65536 dopri5 1 >> gpurun_out/ncu_ode_$T.log 2>&1
grep -c "Report:" gpurun_out/ncu_ode_$T.log
# afterwards, in the build container (no GPU needed): per-source-line breakdowns of the captures
#   ncu -i gpurun_out/prof_tile_$T.ncu-rep --page source --csv --print-source sass,cuda > /tmp/page.csv
#   python tools/summarize_source_page.py /tmp/page.csv --kernel-regex em_tile_kernel --regions '<file:lo-hi=name,...>' > profiles/<round>_tile_source_hotspots.md
#   python tools/summarize_profiles.py            # raw-page summaries + profiles/traffic.json
# pending GPU parity cases of round 2 (run once, then merge into the collected files):
#   python -m pytest tests/pending_round2_gpu.py -m gpu -x -q
