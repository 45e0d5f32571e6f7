#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_golden_gpu.py tests/test_bench_shape_gpu.py -x -q 2>&1 | tail -3
ONLY=none timeout 300 python scripts/bench_mlp2.py 2>&1 | grep -v "poll=1" | head -8
SBR_MLP2_OCC=2 ONLY=none timeout 300 python scripts/bench_mlp2.py 2>&1 | grep "fwd.*poll=0"
timeout 300 python bench.py --no-eval --extra-configs '' --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['value'], l['ms_per_step'], l['e2e']['value'], l['paper_batch'])"
