#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --extra-configs '' --no-eval > gpurun_out/r02_bench_${N}gpu_short.json 2> gpurun_out/r02_bench_${N}gpu_short.err; echo "rc=$?"
tail -c 600 gpurun_out/r02_bench_${N}gpu_short.err
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_${N}gpu_short.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','n_gpus')}, l['config'].get('collective'), l['e2e'])
PY
