#!/bin/bash
# session 3: full GPU suite (glibc pow, restore size, pooled buffers), microbenchmark v2, per-kernel times of a 1/8 share, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
timeout 300 tools/bin/fp64_mix_bench > gpurun_out/s3_fp64mix.jsonl 2>&1
export EE_DEV_AIDS=1
for share in 1/1 3/8; do
  tag=$(echo $share | tr '/' '_')
  EE_SYM_RANGE=$share timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(accel_sym|sym_reduce)' -s 578 -c 8 --csv --log-file gpurun_out/s3_launches_share_$tag.csv python scripts/one_step.py > gpurun_out/s3_ncu_share_$tag.log 2>&1
done
unset EE_DEV_AIDS
timeout 900 python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
echo "bench rc=$?" >> gpurun_out/s3_bench.err
tail -5 gpurun_out/s3_pytest.log
cat gpurun_out/s3_fp64mix.jsonl
grep -h "k_" gpurun_out/s3_launches_share_*.csv | cut -d, -f5,14- | head -40
tail -c 2500 gpurun_out/s3_bench.json
tail -3 gpurun_out/s3_bench.err
