#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_${N}gpu_full.json 2> gpurun_out/r02_bench_${N}gpu_full.err; echo "rc=$?"
tail -c 400 gpurun_out/r02_bench_${N}gpu_full.err
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_${N}gpu_full.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','n_gpus')}, l['config'].get('collective'), l['e2e'])
for k,v in l.get('configs',{}).items(): print(k, {a:v.get(a) for a in ('value','ms_per_step','error')})
e=l.get('eval',{})
print({k:(v if not isinstance(v,(list,dict)) else '...') for k,v in e.items()})
print(e.get('item_sharded_val_split')); print([ (s['U'],s['I'],s['D'],s.get('shards'),round(s['ms'],2)) for s in e.get('sweep',[])])
PY
