#!/bin/bash
# session 5: full suite after the allocator fix; schedule sweep (guided 1/2, maxc 16/64); phase profile with min/max; per-kernel times
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s5_pytest.log
timeout 600 python scripts/sym_variants.py 4,256,2,16 4,128,3,16 > gpurun_out/s5_variants.jsonl 2> gpurun_out/s5_variants.err
export EE_DEV_AIDS=1
for share in 1/1 3/8; do
  tag=$(echo $share | tr '/' '_')
  EE_SYM_PROF=1 EE_SYM_RANGE=$share timeout 300 python scripts/one_step.py 2>&1 | tail -3 > gpurun_out/s5_symprof_$tag.log
  EE_SYM_RANGE=$share timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(accel_sym|sym_reduce)' -s 578 -c 8 --csv --log-file gpurun_out/s5_launches_share_$tag.csv python scripts/one_step.py > gpurun_out/s5_ncu_share_$tag.log 2>&1
done
unset EE_DEV_AIDS
tail -5 gpurun_out/s5_pytest.log
cat gpurun_out/s5_variants.jsonl
cat gpurun_out/s5_symprof_*.log
grep -h "k_" gpurun_out/s5_launches_share_*.csv | awk -F'","' '{print $5, $NF}' | head -20
