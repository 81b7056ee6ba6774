#!/bin/bash
# round 2, session F (1 GPU): full GPU suite after the round's changes, mid-size probe with the one-launch kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest.log
timeout 600 python scripts/mid_probe.py 2048,4096,8192,16384 plain 4,32,16,4 2,32,16,4 4,64,8,4 4,128,4,4 > gpurun_out/f_mid_probe_fused.jsonl 2> gpurun_out/f_mid_probe.err
EE_SYM_FUSED=0 timeout 600 python scripts/mid_probe.py 4096,16384 4,128,4,4 > gpurun_out/f_mid_probe_unfused.jsonl 2>> gpurun_out/f_mid_probe.err
tail -14 gpurun_out/f_pytest.log
python - <<'PY'
import json
for f in ('gpurun_out/f_mid_probe_fused.jsonl','gpurun_out/f_mid_probe_unfused.jsonl'):
    print(f)
    for l in open(f):
        d=json.loads(l)
        print(' ', d['n'], d['variant'], d.get('error') or ('b2b %.4f ms  sync %.4f ms  frac %.3f  rel %.1e'%(d['ms_per_step_back_to_back'], d['ms_per_step_host_sync_each'], d['frac_of_dfma_peak'], d['accel_rel_vs_first'])))
PY
tail -n 3 gpurun_out/f_mid_probe.err
