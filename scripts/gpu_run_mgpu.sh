#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/r02_mgpu_test.log 2>&1; tail -15 gpurun_out/r02_mgpu_test.log
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "rc=$?"
tail -c 1500 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','n_gpus')}, l['config'].get('collective'), l['e2e'])
for k,v in l.get('configs',{}).items(): print(k, {a:v.get(a) for a in ('value','ms_per_step','collective','error')})
print(json.dumps(l.get('eval',{}))[:1500])
PY
