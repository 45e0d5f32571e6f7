#!/bin/bash
# short final pass (1 GPU): GPU tests, full bench line, per-kernel timings of the ML-1M step, ncu launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_final.log 2>&1; tail -3 gpurun_out/r02_gputest_final.log
timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu_final.json 2> gpurun_out/r02_bench_1gpu_final.err; echo "bench rc=$?"
timeout 300 python scripts/profile_step.py ml1m > gpurun_out/r02_prof_ml1m_final.log 2>&1; head -14 gpurun_out/r02_prof_ml1m_final.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches_ncu_final.csv \
  python bench.py --no-eval --extra-configs '' --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r02_bench_under_ncu_final.log 2>&1; echo "ncu rc=$?"
