#!/bin/bash
mkdir -p gpurun_out
run() { G=$1; ex=$2; TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 295$G$3"
timeout 300 $TR bench.py --gpus $G --steps 32 --warmup 3 --exchange $ex > gpurun_out/bench_g${G}_${ex}_final.json 2> gpurun_out/bench_g${G}_${ex}_final.err
grep -h '^{' gpurun_out/bench_g${G}_${ex}_final.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$ex', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])" || tail -n 5 gpurun_out/bench_g${G}_${ex}_final.err; }
run 8 p2p 1
run 8 allreduce 2
run 4 p2p 3
head -c 300 gpurun_out/bench_g8_p2p_final.json
