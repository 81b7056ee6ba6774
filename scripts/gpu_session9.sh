#!/bin/bash
# 8 GPUs: the bench line (peer path) with breakdown and self-check, then the NCCL all-reduce variant
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 32 --warmup 3 > gpurun_out/s9_bench_g8.json 2> gpurun_out/s9_bench_g8.err
echo "bench rc=$?" >> gpurun_out/s9_bench_g8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 32 --warmup 3 --exchange allreduce > gpurun_out/s9_bench_g8_allreduce.json 2> gpurun_out/s9_bench_g8_allreduce.err
python - <<'PY'
import json
for f in ('gpurun_out/s9_bench_g8.json','gpurun_out/s9_bench_g8_allreduce.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d.get('parity_rel'), d['e2e']['value'] if d['e2e'] else None, json.dumps(d.get('p2p_breakdown')))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/s9_bench_g8.err
