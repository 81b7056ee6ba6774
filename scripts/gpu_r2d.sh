#!/bin/bash
# round 2, session D (1 GPU): new GPU tests (ship methods, analytics, mid sizes), launch lists of a mid-size step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ships_gpu.py tests/test_configs_gpu.py -m gpu -q -x -k "adaptive_method or transitions or mid_size or knot_for_knot" > gpurun_out/d_pytest_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest_new.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches_n4096_mid.csv python scripts/one_step.py 4096 > gpurun_out/d_one_step.log 2>&1
EE_DEV_AIDS=1 EE_SYM=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches_n4096_plain.csv python scripts/one_step.py 4096 >> gpurun_out/d_one_step.log 2>&1
timeout 300 python scripts/step_floor.py > gpurun_out/d_step_floor.jsonl 2>&1
tail -15 gpurun_out/d_pytest_new.log
tail -12 gpurun_out/d_launches_n4096_mid.csv | cut -d, -f5,10,15 ; tail -6 gpurun_out/d_launches_n4096_plain.csv | cut -d, -f5,10,15
cat gpurun_out/d_step_floor.jsonl
