#!/bin/bash
# Round-2 final GPU pass (1 GPU): tests, bench (all configs + eval sweep + CPU arm), per-kernel step profiles, ncu launch
# list of the bench command, ncu --set full captures of the two fused kernels, role profile.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_final.log 2>&1; tail -3 gpurun_out/r02_gputest_final.log
timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu_final.json 2> gpurun_out/r02_bench_1gpu_final.err; echo "bench rc=$?"
for w in ml1m onion18_huge amazon_nouser; do timeout 300 python scripts/profile_step.py $w > gpurun_out/r02_prof_${w}_final.log 2>&1; done
head -12 gpurun_out/r02_prof_ml1m_final.log
timeout 300 python scripts/trace_mlp2.py 2>&1 | grep -v median > gpurun_out/r02_mlp2_roles_final.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches_ncu_final.csv \
  python bench.py --no-eval --extra-configs '' --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r02_bench_under_ncu_final.log 2>&1
for k in bwd fwd; do
  ONLY=none timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp2_${k} -s 20 -c 1 -f \
    -o gpurun_out/r02_mlp2_${k}_final python scripts/bench_mlp2.py > gpurun_out/r02_mlp2_${k}_final_ncu.log 2>&1
  tail -2 gpurun_out/r02_mlp2_${k}_final_ncu.log
done
ls -la gpurun_out | tail -12
