#!/bin/bash
# round 2, session N (1 GPU): ship kernel v6 (wider lock-step, 255-register cap) -- parity, timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ships_gpu.py tests/test_configs_gpu.py -m gpu -q -k "ship or relative or adaptive or transitions or c5" > gpurun_out/n_pytest_ships.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n_pytest_ships.log
timeout 900 python scripts/ships_probe.py 1024 0,6,2,7,5 > gpurun_out/n_ships_probe.jsonl 2> gpurun_out/n_ships_probe.err
tail -5 gpurun_out/n_pytest_ships.log
cat gpurun_out/n_ships_probe.jsonl; tail -n 3 gpurun_out/n_ships_probe.err
