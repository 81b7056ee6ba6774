#!/bin/bash
# session 4: memcheck of the failing clone test, microbenchmark v3 (operand reuse), in-kernel phase profile, full suite, bench
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_nbody_gpu.py -x -q -m gpu -k "clone_during_startup" > gpurun_out/s4_memcheck.log 2>&1
timeout 300 tools/bin/fp64_mix_bench > gpurun_out/s4_fp64mix.jsonl 2>&1
export EE_DEV_AIDS=1
for share in 1/1 3/8; do
  tag=$(echo $share | tr '/' '_')
  EE_SYM_PROF=1 EE_SYM_RANGE=$share timeout 300 python scripts/one_step.py 2>&1 | tail -4 > gpurun_out/s4_symprof_$tag.log
done
unset EE_DEV_AIDS
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s4_pytest.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err
echo "bench rc=$?" >> gpurun_out/s4_bench.err
grep -n "Invalid\|ERROR SUMMARY\|at 0x\|in k_\|by thread\|Address" gpurun_out/s4_memcheck.log | head -30
cat gpurun_out/s4_fp64mix.jsonl
cat gpurun_out/s4_symprof_*.log
tail -8 gpurun_out/s4_pytest.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s4_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_rel')}, d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d['extras']['C2_planner_loop'])[:1500])
PY
