#!/bin/bash
# round 2, re-entry session A (1 GPU): full GPU suite, the bench line, the reference arm, launch list, ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 900 python bench.py > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err
echo "bench rc=$?" >> gpurun_out/a_bench_n1.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/a_bench_reference.json 2> gpurun_out/a_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/a_launches_n65536.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-parity-kernel > gpurun_out/a_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_accel_sym|k_sym_reduce" -s 600 -c 2 -f -o gpurun_out/a_prof_sym \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-parity-kernel > gpurun_out/a_ncu_sym.log 2>&1
tail -12 gpurun_out/a_pytest.log
tail -3 gpurun_out/a_bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/a_bench_n1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','parity_rel','gpu_launches')}, d['roofline']['frac'], d['e2e'], d['clocks'])
    print(json.dumps(d.get('extras'))[:3000])
    print(json.dumps(d.get('cpu_baseline'))[:800])
except Exception as e: print('ERR', e)
PY
cat gpurun_out/a_bench_reference.json | cut -c1-600
grep -c k_accel_sym gpurun_out/a_launches_n65536.csv; tail -4 gpurun_out/a_launches_n65536.csv
tail -2 gpurun_out/a_ncu_sym.log
