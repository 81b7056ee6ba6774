#!/bin/bash
# round 2, session B (1 GPU): mid-size variants of the pair-symmetric kernel, full ncu captures (sym pair, ships, small)
mkdir -p gpurun_out
timeout 600 python scripts/mid_probe.py > gpurun_out/b_mid_probe.jsonl 2> gpurun_out/b_mid_probe.err
timeout 600 python -m pytest tests/test_nbody_gpu.py -m gpu -q -x -k "throughput or symmetric or blanes" > gpurun_out/b_pytest_thr.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_accel_sym|k_sym_reduce" -s 590 -c 4 -f -o gpurun_out/b_prof_sym \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-parity-kernel > gpurun_out/b_ncu_sym.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_small_steps -s 3 -c 2 -f -o gpurun_out/b_prof_small \
    python -m pytest tests/test_configs_gpu.py -m gpu -q -x -k "c2" > gpurun_out/b_ncu_small.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_ships_step_to -c 1 -f -o gpurun_out/b_prof_ships \
    python -m pytest tests/test_configs_gpu.py -m gpu -q -x -k "c5" > gpurun_out/b_ncu_ships.log 2>&1
cat gpurun_out/b_mid_probe.jsonl; tail -3 gpurun_out/b_mid_probe.err
tail -5 gpurun_out/b_pytest_thr.log
tail -3 gpurun_out/b_ncu_sym.log | cut -c1-300; tail -3 gpurun_out/b_ncu_small.log | cut -c1-300; tail -3 gpurun_out/b_ncu_ships.log | cut -c1-300
ls -la gpurun_out
