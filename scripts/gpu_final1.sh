#!/bin/bash
# final single-GPU session of the round: full GPU test suite, smoke, bench (both arms), other configs, ncu evidence
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1
timeout 600 python bench.py --steps 32 --warmup 3 > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err
timeout 300 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
timeout 900 python scripts/bench_configs.py > gpurun_out/bench_configs_final.json 2> gpurun_out/bench_configs_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 40 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_accel_sym|k_sym_reduce" -s 590 -c 2 -f -o gpurun_out/prof_sym_final \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
tail -n 3 gpurun_out/pytest_gpu_final.log; cat gpurun_out/smoke_final.log | tail -n 2; cat gpurun_out/bench_n1_final.json; cat gpurun_out/bench_configs_final.json; tail -n 4 gpurun_out/launches_final.csv | cut -c1-200
