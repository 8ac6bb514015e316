set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 1000 --warmup 20 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err || tail -5 gpurun_out/bench_r1k.err
cat gpurun_out/bench_r1k.json | head -c 1500
python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/plain_r1k.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1k.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_ll_r1k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_tile_kernel -s 60 -c 2 -f -o gpurun_out/prof_tile_r1k python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full_r1k.log 2>&1; tail -2 gpurun_out/ncu_full_r1k.log
