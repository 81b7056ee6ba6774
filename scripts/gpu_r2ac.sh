#!/bin/bash
# round 2, session AC (1 GPU): ship kernel with 3 / 43 / 140 bodies (shared-memory and global-memory caches), all ship tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ships_gpu.py tests/test_configs_gpu.py -m gpu -q -k "ship or relative or adaptive or transitions or c5 or few_and" --durations=5 > gpurun_out/ac_pytest_ships.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ac_pytest_ships.log
tail -12 gpurun_out/ac_pytest_ships.log
