#!/bin/bash
# round 2, session S (1 GPU): new ship-system test, run-ahead tests, planner loop after the early-launch rule
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ships_gpu.py tests/test_nbody_gpu.py tests/test_zz_small_sizes_gpu.py -m gpu -q -k "few_and_with_more or planner or run_ahead or snapshot or clone or small_kernel or spline_solution" > gpurun_out/s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s_pytest.log
python - > gpurun_out/s_planner.json 2> gpurun_out/s_planner.err <<'PY'
import json, sys
sys.path.insert(0, '.')
import bench
import ephemeris_explorer_b200 as ee
s = ee.formats.load_system(bench.ROOT / "tests" / "golden" / "systems" / "full_solar_system_2433282.5.json")
for _ in range(3):
    print(json.dumps(bench.planner_loop_c2(ee, s, 0)), flush=True)
PY
tail -6 gpurun_out/s_pytest.log
python - <<'PY'
import json
for l in open('gpurun_out/s_planner.json'):
    d=json.loads(l); print(d['gpu']['steps_per_s'], d['gpu']['gpu_launches'], d['cpu_oracle']['steps_per_s'], d['speedup_vs_cpu_oracle'], d['splines_bit_exact'])
PY
tail -n 3 gpurun_out/s_planner.err
