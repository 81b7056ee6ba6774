#!/bin/bash
# round 2, session AA (1 GPU): ncu --set full of the two kernels of a 4 096-body step
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_accel_sym|k_sym_reduce" -s 578 -c 8 -f -o gpurun_out/aa_prof_mid \
    python scripts/one_step.py 4096 > gpurun_out/aa_ncu_mid.log 2>&1
tail -n 3 gpurun_out/aa_ncu_mid.log | cut -c1-200
