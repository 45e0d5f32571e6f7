#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "batchnorm or infonce" > gpurun_out/r02_t2.log 2>&1; tail -5 gpurun_out/r02_t2.log
timeout 600 python -m pytest tests/test_golden_gpu.py tests/test_bench_shape_gpu.py -x -q > gpurun_out/r02_t2b.log 2>&1; tail -5 gpurun_out/r02_t2b.log
for w in onion18_huge amazon_nouser; do timeout 300 python scripts/profile_step.py $w > gpurun_out/r02_prof_$w.log 2>&1; head -45 gpurun_out/r02_prof_$w.log; done
