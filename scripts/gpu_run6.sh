#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ONLY=none timeout 300 python scripts/bench_mlp2.py 2>&1 | tee gpurun_out/r02_mlp2_attr2.log
for i in 1 2; do timeout 300 python bench.py --no-eval --extra-configs '' --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('early-resolve', l['value'], l['ms_per_step'], l['e2e']['value'], l['paper_batch']['ms_per_step']); print([(k['family'], k['avg_us']) for k in l['kernel_families'][:4]])"; done
