for f in 0 512 1024 1536; do
QRL_EM_DEBUG_FLAGS=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/s2_$f.json 2> gpurun_out/s2_$f.err
python -c "
import json
d=json.loads(open('gpurun_out/s2_$f.json').read().strip().splitlines()[-1]); print('flags $f', d['n_gpus'], d['value'], d['ms_per_step'])"
done
