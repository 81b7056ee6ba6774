#!/bin/bash
# round 2, session AB (1 GPU): reduce kernel with deeper load unrolling -- mid sizes and the headline
mkdir -p gpurun_out
timeout 600 python scripts/mid_probe.py 4096,8192,16384 4,128,4,4 > gpurun_out/ab_mid_probe.jsonl 2> gpurun_out/ab_mid_probe.err
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/ab_bench_n1.json 2> gpurun_out/ab_bench_n1.err
python - <<'PY'
import json
for l in open('gpurun_out/ab_mid_probe.jsonl'):
    d=json.loads(l)
    print(' ', d['n'], d['variant'], d.get('error') or ('b2b %.4f ms  sync %.4f ms  frac %.3f'%(d['ms_per_step_back_to_back'], d['ms_per_step_host_sync_each'], d['frac_of_dfma_peak'])))
d=json.loads(open('gpurun_out/ab_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_rel')}, d['roofline']['frac'])
PY
