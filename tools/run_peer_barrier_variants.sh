# A/B of the fused cross-GPU barrier (diagnostic flags of qrl_em_args.flags): 0 default, 2048 with the extra system fence
# after the flag stores, 1024 no wait at all; repeated to see the run-to-run spread
for rep in 1 2; do for f in 0 2048 1024; do
QRL_EM_DEBUG_FLAGS=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/s2_$f.json 2> gpurun_out/s2_$f.err
python -c "
import json
d=json.loads(open('gpurun_out/s2_$f.json').read().strip().splitlines()[-1]); print('flags $f', d['n_gpus'], '%.3e' % d['value'], d['ms_per_step'])"
done; done
