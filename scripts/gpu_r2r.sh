#!/bin/bash
# round 2, session R (1 GPU): the round's closing run -- full GPU suite, smoke, bench (ours + reference arm), launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r_smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/r_bench_reference.json 2> gpurun_out/r_bench_reference.err
timeout 900 python bench.py > gpurun_out/r_bench_n1.json 2> gpurun_out/r_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r_launches_n65536.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-parity-kernel > gpurun_out/r_ncu_launches.log 2>&1
tail -8 gpurun_out/r_pytest.log; cat gpurun_out/r_smoke.log | tail -2
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r_bench_n1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','parity_rel','gpu_launches')}, d['roofline']['frac'], d['roofline']['frac_of_nominal'], d['e2e']['value'], d['clocks'])
    print(json.dumps(d.get('extras'))[:3500])
    print(json.dumps(d.get('cpu_baseline'))[:600])
except Exception as e: print('ERR', e)
PY
cut -c1-400 gpurun_out/r_bench_reference.json; tail -n 3 gpurun_out/r_bench_n1.err
