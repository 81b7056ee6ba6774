#!/bin/bash
# round 2, session V (1 GPU): pair-symmetric inner-loop experiment -- headline timing + parity
mkdir -p gpurun_out
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/v_bench_n1.json 2> gpurun_out/v_bench_n1.err
timeout 600 python -m pytest tests/test_configs_gpu.py -m gpu -q -k "c4 or mid_size" > gpurun_out/v_pytest.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/v_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_rel')}, d['roofline']['frac'])
PY
tail -3 gpurun_out/v_pytest.log
