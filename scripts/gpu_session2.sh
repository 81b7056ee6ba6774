#!/bin/bash
# session 2: full GPU suite, variant sweep after the reduce fix, FP64 mix microbenchmark, 1-GPU bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
timeout 300 tools/bin/fp64_mix_bench > gpurun_out/s2_fp64mix.jsonl 2>&1
timeout 600 python scripts/sym_variants.py 4,256,2,16 4,128,3,16 4,256,1,16 > gpurun_out/s2_variants.jsonl 2> gpurun_out/s2_variants.err
timeout 900 python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
echo "bench rc=$?" >> gpurun_out/s2_bench.err
tail -5 gpurun_out/s2_pytest.log
cat gpurun_out/s2_fp64mix.jsonl gpurun_out/s2_variants.jsonl
tail -c 3000 gpurun_out/s2_bench.json
tail -5 gpurun_out/s2_bench.err
