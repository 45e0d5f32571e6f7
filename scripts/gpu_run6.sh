#!/bin/bash
mkdir -p gpurun_out
ONLY=none timeout 300 python scripts/bench_mlp2.py 2>&1 | tee gpurun_out/r02_mlp2_attr2.log
timeout 300 python bench.py --no-eval --extra-configs '' --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['value'], l['ms_per_step'], l['e2e']['value'], l['paper_batch']); print([(k['family'], k['share'], k['avg_us']) for k in l['kernel_families'][:8]])"
SBR_MLP2_BN_TAIL=0 timeout 300 python bench.py --no-eval --extra-configs '' --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no tail', l['value'], l['ms_per_step'])"
timeout 300 python bench.py --no-eval --extra-configs '' --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tail again', l['value'], l['ms_per_step'])"
