#!/bin/bash
# round 2, session M (G GPUs): the bench line on the peer path (scaling table)
mkdir -p gpurun_out
G=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $G --steps 32 --warmup 3 > gpurun_out/m_bench_g$G.json 2> gpurun_out/m_bench_g$G.err
echo "bench rc=$?" >> gpurun_out/m_bench_g$G.err
python - <<PY
import json
d=json.loads(open('gpurun_out/m_bench_g$G.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d.get('parity_rel'), d['e2e']['value'], json.dumps(d.get('p2p_breakdown'))[:600])
PY
tail -n 2 gpurun_out/m_bench_g$G.err | cut -c1-200
