#!/bin/bash
for i in 1 2 3 4 5; do timeout 300 python bench.py --no-eval --extra-configs '' --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run', l['value'], l['ms_per_step'], l['e2e']['value'], l['paper_batch']['ms_per_step'], l['paper_batch']['e2e'])"; done
