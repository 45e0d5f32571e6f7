#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t4.log 2>&1; tail -6 gpurun_out/r02_t4.log
for w in ml1m onion18_huge amazon_nouser; do timeout 300 python scripts/profile_step.py $w > gpurun_out/r02_prof_$w.log 2>&1; head -24 gpurun_out/r02_prof_$w.log; done
