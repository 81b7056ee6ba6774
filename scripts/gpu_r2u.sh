#!/bin/bash
# round 2, session U (1 GPU): in-kernel phase profile of the pair-symmetric kernel at 65 536 bodies; racecheck after the fix
mkdir -p gpurun_out
EE_DEV_AIDS=1 EE_SYM_PROF=1 timeout 300 python scripts/one_step.py 2>&1 | tail -4 > gpurun_out/u_symprof.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 8 python scripts/ships_tiny.py > gpurun_out/u_racecheck.log 2>&1
cat gpurun_out/u_symprof.log
grep -E "RACECHECK SUMMARY|tiny ok" gpurun_out/u_racecheck.log
