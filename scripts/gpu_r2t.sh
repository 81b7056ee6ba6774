#!/bin/bash
# round 2, session T (1 GPU): compute-sanitizer over the round's new kernels (memcheck, racecheck, initcheck)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 8 python scripts/ships_tiny.py > gpurun_out/t_memcheck.log 2>&1
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 8 python scripts/ships_tiny.py > gpurun_out/t_racecheck.log 2>&1
timeout 900 compute-sanitizer --tool synccheck --print-limit 8 python scripts/ships_tiny.py > gpurun_out/t_synccheck.log 2>&1
for f in memcheck racecheck synccheck; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|tiny ok|Error|error|hazard" gpurun_out/t_$f.log | head -12; done
