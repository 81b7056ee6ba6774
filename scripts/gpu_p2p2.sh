#!/bin/bash
G=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi_check_g$G.log 2>&1; echo "check exit $?" >> gpurun_out/multi_check_g$G.log
grep -h '32768' gpurun_out/multi_check_g$G.log | cut -c1-200; tail -n 1 gpurun_out/multi_check_g$G.log
for cfg in "p2p 512" "p2p 256" "p2p 128" "allreduce 128"; do set -- $cfg
EE_SYM_JS=$2 timeout 300 $TR --master-port 29513 bench.py --gpus $G --steps 24 --warmup 3 --exchange $1 2>/dev/null | grep -h '^{' | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1 js$2', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
