"""Per-call CUDA-event breakdown of one fused train step of a BASELINE workload (same set-up as bench.py).

    python scripts/profile_step.py [ml1m|onion18_huge|amazon_nouser] [batch]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sibrar_b200  # noqa
import bench
from sibrar_b200 import ops, workloads
from sibrar_b200.sbnet import SingleBranchNet
from sibrar_b200.trainer import FusedTrainer

name = sys.argv[1] if len(sys.argv) > 1 else "ml1m"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
dev = torch.device("cuda", 0)
corpus, conf, learn, _, _ = workloads.build(name)
train = corpus.dataset("train")
torch.manual_seed(1234)
model = SingleBranchNet.build_from_conf(conf, train).to(dev).train()
tr = FusedTrainer(model, learn, n_negative_samples=bench.N_NEG)
coo = train.interaction_matrix
d = lambda a, t: torch.from_numpy(np.ascontiguousarray(a).astype(t)).to(dev)
csr = train.user_sampling_matrix_train
coo_u, coo_i = d(coo.row, np.int32), d(coo.col, np.int32)
ip, ix, items = d(csr.indptr, np.int64), d(csr.indices, np.int32), d(train.items_in_split, np.int32)
step = torch.zeros(1, dtype=torch.int64, device=dev)
u = torch.empty(B, dtype=torch.int64, device=dev)
i = torch.empty((B, 11), dtype=torch.int64, device=dev)
ops.tick(step)
ops.sample_batch(coo_u, coo_i, ip, ix, items, B, 10, 1000, step, u, i)
for _ in range(5):
    tr.step(u, i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    tr.step(u, i)
b.record()
torch.cuda.synchronize()
print(f"{name} B={B}: {a.elapsed_time(b) / 10:.3f} ms/step unprofiled (graph replay, warm L2)")
tr.cuda_graph, tr.branches = False, 0
n_prof = 3
with bench.CallProfiler(ops, torch) as prof:
    for _ in range(n_prof):
        torch.cuda._sleep(20_000_000)
        tr.step(u, i)
agg = prof.summary()
tot = sum(v[0] for v in agg.values()) / n_prof
print(f"sum of kernels {tot:.3f} ms/step, {sum(v[1] for v in agg.values()) / n_prof:.0f} calls/step")
nnz = float(train.user_sampling_matrix_train.nnz)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    us = v[0] / v[1] * 1e3
    w = bench.algorithmic_work(k, nnz)
    rate = ""
    if w is not None and us > 0:
        rate = f"{w[0] / us / 1e6:8.1f} TFLOP/s {w[1] / us / 1e3:8.1f} GB/s"
    print(f"{v[0] / n_prof / tot * 100:5.1f}%  {us:8.1f} us x {v[1] / n_prof:4.1f}  {rate}  {k}")
