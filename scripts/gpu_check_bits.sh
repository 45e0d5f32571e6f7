#!/bin/bash
# bit-packed GEMM with TMA-loaded bit words: parity, kernel timing, ML-1M step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gemm_bits" 2>&1 | tail -5 || exit 1
timeout 120 python scripts/profile_gemm_bits.py 2>&1 | tee gpurun_out/r02_gemm_bits_tma.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --no-eval --extra-configs '' --no-cpu-baseline --steps 50 --warmup 10 > gpurun_out/r02_bench_bits_tma.json 2> gpurun_out/r02_bench_bits_tma.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_bits_tma.json'):
    if l.startswith('{'):
        d = json.loads(l); print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'paper', d.get('paper_batch'))
PY
