#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/s7_clone.log
for sc in no_clone startup_clone startup_clone_sync steady_clone orig_after_clone startup_clone_generic_only startup_clone_two_takes; do
  timeout 120 python scripts/clone_debug.py $sc >> gpurun_out/s7_clone.log 2>&1 || echo "scenario $sc FAILED" >> gpurun_out/s7_clone.log
done
timeout 600 python scripts/sym_variants.py 4,256,2,16 4,128,3,16 > gpurun_out/s7_variants.jsonl 2> gpurun_out/s7_variants.err
export EE_DEV_AIDS=1
EE_SYM_PROF=1 EE_SYM_RANGE=3/8 timeout 300 python scripts/one_step.py 2>&1 | tail -4 > gpurun_out/s7_symprof.log
unset EE_DEV_AIDS
grep -n "scenario\|illegal" gpurun_out/s7_clone.log | cut -c1-200
cat gpurun_out/s7_variants.jsonl gpurun_out/s7_symprof.log
