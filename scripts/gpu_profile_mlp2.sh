#!/bin/bash
# GPU-box profiling pass of the fused gather + SB-MLP kernels (run through gpurun); outputs under gpurun_out/
set -x
mkdir -p gpurun_out
python scripts/bench_mlp2.py > gpurun_out/r02_mlp2_attr.log 2>&1
tail -40 gpurun_out/r02_mlp2_attr.log
for k in bwd fwd; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp2_${k} -s 20 -c 1 -f \
    -o gpurun_out/r02_mlp2_${k} python scripts/bench_mlp2.py > gpurun_out/r02_mlp2_${k}_ncu.log 2>&1
  tail -3 gpurun_out/r02_mlp2_${k}_ncu.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches_ncu.csv \
  python bench.py --no-eval --extra-configs '' --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r02_bench_under_ncu.log
ls -la gpurun_out
