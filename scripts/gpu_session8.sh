#!/bin/bash
# 2 GPUs: sharded runs against the 1-GPU result, then the 2-GPU bench line (peer path) with breakdown and self-check
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/s8_multi_check_g2.log 2>&1
echo "check rc=$?" >> gpurun_out/s8_multi_check_g2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 32 --warmup 3 > gpurun_out/s8_bench_g2.json 2> gpurun_out/s8_bench_g2.err
echo "bench rc=$?" >> gpurun_out/s8_bench_g2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 32 --warmup 3 --exchange allreduce > gpurun_out/s8_bench_g2_allreduce.json 2> gpurun_out/s8_bench_g2_allreduce.err
grep -h "check\|ok" gpurun_out/s8_multi_check_g2.log | tail -12
tail -c 1800 gpurun_out/s8_bench_g2.json; tail -3 gpurun_out/s8_bench_g2.err
python - <<'PY'
import json
for f in ('gpurun_out/s8_bench_g2.json','gpurun_out/s8_bench_g2_allreduce.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d.get('parity_rel'), d['e2e']['value'] if d['e2e'] else None)
    except Exception as e: print(f, 'ERR', e)
PY
