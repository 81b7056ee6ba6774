#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nbody_gpu.py -m gpu -q -k "symmetric or reference_systems or spline or step_to or clone" > gpurun_out/pytest_quick.log 2>&1; tail -n 5 gpurun_out/pytest_quick.log
timeout 300 python bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json; tail -n 3 gpurun_out/bench_quick.err
