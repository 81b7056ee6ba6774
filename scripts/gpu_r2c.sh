#!/bin/bash
# round 2, session C (8 GPUs): correctness of the sharded paths against 1 GPU, the bench line on the peer path with its
# breakdown and self-check, and the NCCL all-reduce variant
mkdir -p gpurun_out
G=${1:-8}
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/c_pytest_multi_g$G.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest_multi_g$G.log
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $G --steps 32 --warmup 3 > gpurun_out/c_bench_g$G.json 2> gpurun_out/c_bench_g$G.err
echo "bench rc=$?" >> gpurun_out/c_bench_g$G.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $G --steps 32 --warmup 3 --exchange allreduce > gpurun_out/c_bench_g${G}_allreduce.json 2> gpurun_out/c_bench_g${G}_allreduce.err
tail -4 gpurun_out/c_pytest_multi_g$G.log
python - <<PY
import json
for f in ('gpurun_out/c_bench_g$G.json','gpurun_out/c_bench_g${G}_allreduce.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d.get('parity_rel'), d['e2e']['value'] if d['e2e'] else None, json.dumps(d.get('p2p_breakdown')))
    except Exception as e: print(f, 'ERR', e)
PY
grep -c "NCCL INFO" gpurun_out/c_bench_g$G.err; grep "nranks\|NVLS" gpurun_out/c_bench_g$G.err | head -5; tail -3 gpurun_out/c_bench_g$G.err | cut -c1-300
