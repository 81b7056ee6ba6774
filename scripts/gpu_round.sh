#!/bin/bash
# One GPU-box session: tests, bench, ncu launch list + full capture of the dominant kernel.  Outputs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1; nproc >> gpurun_out/smi.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 32 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?" >> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accel_fast -s 295 -c 2 -f -o gpurun_out/prof_accel \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n1.json; tail -n 2 gpurun_out/bench_n1.err; cat gpurun_out/bench_ref.json; tail -n 5 gpurun_out/launches.csv; tail -n 3 gpurun_out/ncu_full.log
