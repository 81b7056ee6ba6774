#!/bin/bash
# round 2, session H (1 GPU): ship kernel v2 -- parity, timing, ncu of the Verner87 kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ships_gpu.py tests/test_configs_gpu.py -m gpu -q -k "ship or relative or adaptive or transitions or c5" > gpurun_out/h_pytest_ships.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest_ships.log
timeout 900 python scripts/ships_probe.py 1024 0,6,2,7 > gpurun_out/h_ships_probe.jsonl 2> gpurun_out/h_ships_probe.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ships_step_to -c 1 -f -o gpurun_out/h_prof_ships \
    python scripts/ships_probe.py 1024 0 > gpurun_out/h_ncu_ships.log 2>&1
tail -5 gpurun_out/h_pytest_ships.log
cat gpurun_out/h_ships_probe.jsonl; tail -n 3 gpurun_out/h_ships_probe.err; tail -n 2 gpurun_out/h_ncu_ships.log | cut -c1-200
