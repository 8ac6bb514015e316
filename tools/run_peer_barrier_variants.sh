# A/B of the fused cross-GPU barrier (diagnostic bits of qrl_em_args.flags, set through QRL_EM_DEBUG_FLAGS):
#   0     default: release fence + flag stores before the local bookkeeping, acquire-load polling
#   4096  plain volatile polling + a system fence after the loop
#   2048  extra system fence between the flag stores and the polling loop
#   1024  no wait at all (lower bound: peer stores + fences only)
# repeated to see the run-to-run spread; bench.py prints host-enqueue and device time per step on stderr
python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do for f in 0 4096 1024; do
QRL_EM_DEBUG_FLAGS=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 20 --no-cpu-baseline 2>&1 >/dev/null | grep "\[bench\] rank 0" | sed "s/^/flags $f /"
done; done
