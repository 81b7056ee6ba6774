#!/bin/bash
# round 2, session Q (1 GPU): five-level warp-parallel bisection -- parity, timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ships_gpu.py tests/test_configs_gpu.py -m gpu -q -k "ship or relative or adaptive or transitions or c5" > gpurun_out/q_pytest_ships.log 2>&1
echo "pytest rc=$?" >> gpurun_out/q_pytest_ships.log
timeout 900 python scripts/ships_probe.py 1024 0 > gpurun_out/q_ships_probe.jsonl 2> gpurun_out/q_ships_probe.err
tail -5 gpurun_out/q_pytest_ships.log
cat gpurun_out/q_ships_probe.jsonl; tail -n 3 gpurun_out/q_ships_probe.err
