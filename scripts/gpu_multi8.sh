#!/bin/bash
G=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi_check_g$G.log 2>&1; echo "check exit $?" >> gpurun_out/multi_check_g$G.log
timeout 300 $TR --master-port 29513 bench.py --gpus $G --steps 32 --warmup 3 > gpurun_out/bench_g${G}_allreduce.json 2> gpurun_out/bench_g${G}_allreduce.err
grep -h '^{' gpurun_out/multi_check_g$G.log; tail -n 2 gpurun_out/multi_check_g$G.log
cat gpurun_out/bench_g${G}_allreduce.json; tail -n 3 gpurun_out/bench_g${G}_allreduce.err
