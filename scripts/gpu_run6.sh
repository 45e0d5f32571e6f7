#!/bin/bash
for i in 1 2; do timeout 300 python bench.py --no-eval --extra-configs '' --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e-pipelined', l['value'], l['ms_per_step'], l['e2e'], l['paper_batch'])"; done
