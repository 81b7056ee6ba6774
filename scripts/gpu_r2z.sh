#!/bin/bash
# round 2, session Z (1 GPU): ship kernel v7 (compiled-in group count, running per-lane sums, one vote per attempt)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ships_gpu.py tests/test_configs_gpu.py -m gpu -q -k "ship or relative or adaptive or transitions or c5 or few_and" > gpurun_out/z_pytest_ships.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z_pytest_ships.log
timeout 900 python scripts/ships_probe.py 1024 0,6,2,7 > gpurun_out/z_ships_probe.jsonl 2> gpurun_out/z_ships_probe.err
tail -5 gpurun_out/z_pytest_ships.log
cat gpurun_out/z_ships_probe.jsonl | cut -c1-330; tail -n 3 gpurun_out/z_ships_probe.err
