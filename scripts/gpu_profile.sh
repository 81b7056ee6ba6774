#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/bench_configs.py > gpurun_out/bench_configs.json 2> gpurun_out/bench_configs.err; echo "configs exit $?" >> gpurun_out/bench_configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 40 --csv --log-file gpurun_out/launches_sym.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches_sym.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accel_sym -s 295 -c 1 -f -o gpurun_out/prof_sym \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_sym.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_small_steps|k_ships_step_to" -c 3 -f -o gpurun_out/prof_small_ships \
    python scripts/bench_configs.py --quick > gpurun_out/ncu_small.log 2>&1
cat gpurun_out/bench_configs.json; tail -n 3 gpurun_out/bench_configs.err; tail -n 6 gpurun_out/launches_sym.csv; tail -n 2 gpurun_out/ncu_sym.log gpurun_out/ncu_small.log
