#!/bin/bash
G=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi_check_g$G.log 2>&1; echo "check exit $?" >> gpurun_out/multi_check_g$G.log
grep -h '^{' gpurun_out/multi_check_g$G.log; tail -n 4 gpurun_out/multi_check_g$G.log
for ex in p2p allreduce; do
timeout 300 $TR --master-port 29513 bench.py --gpus $G --steps 32 --warmup 3 --exchange $ex > gpurun_out/bench_g${G}_$ex.json 2> gpurun_out/bench_g${G}_$ex.err
python -c "import json;d=json.load(open('gpurun_out/bench_g${G}_$ex.json'));print('$ex', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])" || tail -n 5 gpurun_out/bench_g${G}_$ex.err
done
