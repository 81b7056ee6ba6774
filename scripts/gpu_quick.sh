#!/bin/bash
# quick session: selected tests + bench (sym vs plain)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nbody_gpu.py -m gpu -q -k "symmetric or throughput" > gpurun_out/pytest_sym.log 2>&1; tail -n 5 gpurun_out/pytest_sym.log
timeout 300 python bench.py --steps 32 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_sym.json 2> gpurun_out/bench_sym.err; cat gpurun_out/bench_sym.json; tail -n 3 gpurun_out/bench_sym.err
EE_SYM=0 timeout 300 python bench.py --steps 32 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; cat gpurun_out/bench_plain.json
