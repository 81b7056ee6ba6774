#!/bin/bash
# usage: gpu_multi.sh G   -- sharded correctness + bench at G GPUs (both exchanges)
G=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi_check_g$G.log 2>&1; echo "check exit $?" >> gpurun_out/multi_check_g$G.log
timeout 600 $TR --master-port 29512 bench.py --gpus $G --steps 32 --warmup 3 --exchange allgather > gpurun_out/bench_g${G}_allgather.json 2> gpurun_out/bench_g${G}_allgather.err
timeout 600 $TR --master-port 29513 bench.py --gpus $G --steps 32 --warmup 3 --exchange allreduce > gpurun_out/bench_g${G}_allreduce.json 2> gpurun_out/bench_g${G}_allreduce.err
grep -h '^{' gpurun_out/multi_check_g$G.log; tail -n 2 gpurun_out/multi_check_g$G.log
cat gpurun_out/bench_g${G}_allgather.json gpurun_out/bench_g${G}_allreduce.json; tail -n 3 gpurun_out/bench_g${G}_allgather.err gpurun_out/bench_g${G}_allreduce.err
