#!/bin/bash
# round 2, session E (1 GPU): new GPU tests, mid-size probe with and without programmatic dependent launch, headline check
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ships_gpu.py tests/test_configs_gpu.py tests/test_nbody_gpu.py -m gpu -q -k "adaptive_method or transitions or mid_size or symmetric or throughput or c4" > gpurun_out/e_pytest_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest_new.log
timeout 600 python scripts/mid_probe.py 4096,8192,16384 > gpurun_out/e_mid_probe_pdl.jsonl 2> gpurun_out/e_mid_probe.err
EE_PDL=0 timeout 600 python scripts/mid_probe.py 4096 plain 4,32,16,4 2,32,16,4 4,256,2,16 > gpurun_out/e_mid_probe_nopdl.jsonl 2>> gpurun_out/e_mid_probe.err
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err
tail -25 gpurun_out/e_pytest_new.log
python - <<'PY'
import json
for f in ('gpurun_out/e_mid_probe_pdl.jsonl','gpurun_out/e_mid_probe_nopdl.jsonl'):
    print(f)
    for l in open(f):
        d=json.loads(l)
        print(' ', d['n'], d['variant'], d.get('error') or ('b2b %.4f ms  sync %.4f ms  frac %.3f  rel %.1e'%(d['ms_per_step_back_to_back'], d['ms_per_step_host_sync_each'], d['frac_of_dfma_peak'], d['accel_rel_vs_first'])))
try:
    d=json.loads(open('gpurun_out/e_bench_n1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','parity_rel','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])
except Exception as e: print('ERR', e)
PY
tail -3 gpurun_out/e_mid_probe.err gpurun_out/e_bench_n1.err
