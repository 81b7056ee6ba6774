#!/bin/bash
G=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
for ex in p2p allreduce; do
timeout 300 $TR --master-port 29513 bench.py --gpus $G --steps 32 --warmup 3 --exchange $ex > gpurun_out/bench_g${G}_$ex.json 2> gpurun_out/bench_g${G}_$ex.err
grep -h '^{' gpurun_out/bench_g${G}_$ex.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$ex', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])" || tail -n 5 gpurun_out/bench_g${G}_$ex.err
done
EE_SYM_JS=256 timeout 300 $TR --master-port 29514 bench.py --gpus $G --steps 32 --warmup 3 --exchange p2p 2>/dev/null | grep -h '^{' | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('p2p-js256', d['n_gpus'], d['value'], d['ms_per_step'])"
