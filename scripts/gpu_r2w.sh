#!/bin/bash
# round 2, session W (1 GPU): mid-size probe after the reduce shape follows the system size
mkdir -p gpurun_out
timeout 600 python scripts/mid_probe.py 2048,4096,8192,16384 plain 4,128,4,4 > gpurun_out/w_mid_probe.jsonl 2> gpurun_out/w_mid_probe.err
timeout 600 python -m pytest tests/test_configs_gpu.py -m gpu -q -k "mid_size" > gpurun_out/w_pytest.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/w_mid_probe.jsonl'):
    d=json.loads(l)
    print(' ', d['n'], d['variant'], d.get('error') or ('b2b %.4f ms  sync %.4f ms  frac %.3f  rel %.1e'%(d['ms_per_step_back_to_back'], d['ms_per_step_host_sync_each'], d['frac_of_dfma_peak'], d['accel_rel_vs_first'])))
PY
tail -3 gpurun_out/w_pytest.log
