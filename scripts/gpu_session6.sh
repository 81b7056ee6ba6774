#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_nbody_gpu.py -x -q -m gpu -k "clone_during_startup" > gpurun_out/s6_blocking.log 2>&1
export EE_DEV_AIDS=1
EE_SYM_PROF=1 EE_SYM_PROF_DUMP=1 EE_SYM_RANGE=3/8 timeout 300 python scripts/one_step.py 2>&1 | tail -300 > gpurun_out/s6_symprof_dump.log
unset EE_DEV_AIDS
timeout 900 python -m pytest tests/test_nbody_gpu.py -q -m gpu -k "pair_kernel or blanes or snapshot or planner or run_ahead or state_async" > gpurun_out/s6_pytest_new.log 2>&1
timeout 900 python -m pytest tests/test_ships_gpu.py tests/test_configs_gpu.py -q -m gpu > gpurun_out/s6_pytest_ships.log 2>&1
grep -n "^E \|passed\|failed" gpurun_out/s6_blocking.log | head
tail -3 gpurun_out/s6_pytest_new.log; tail -3 gpurun_out/s6_pytest_ships.log
tail -4 gpurun_out/s6_symprof_dump.log
