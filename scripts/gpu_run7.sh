#!/bin/bash
mkdir -p gpurun_out
ONLY=none timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp2_bwd -s 20 -c 1 -f \
    -o gpurun_out/r02_mlp2_bwd_v2 python scripts/bench_mlp2.py > gpurun_out/r02_mlp2_bwd_v2_ncu.log 2>&1
tail -3 gpurun_out/r02_mlp2_bwd_v2_ncu.log
