#!/bin/bash
# session 1: new tests + variant sweep + ncu of the two leading variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
timeout 900 python -m pytest tests/test_configs_gpu.py tests/test_nbody_gpu.py tests/test_zz_small_sizes_gpu.py -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
timeout 600 python scripts/sym_variants.py > gpurun_out/s1_variants.jsonl 2> gpurun_out/s1_variants.err
export EE_DEV_AIDS=1
for v in 4,256,2,16 4,128,3,16; do
  tag=$(echo $v | tr ',' '_')
  EE_SYM_VARIANT=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_accel_sym -s 290 -c 1 -o gpurun_out/s1_prof_$tag -f python scripts/one_step.py > gpurun_out/s1_ncu_$tag.log 2>&1
done
tail -5 gpurun_out/s1_pytest.log
cat gpurun_out/s1_variants.jsonl
