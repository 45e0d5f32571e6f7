"""Per-call CUDA-event breakdown of one fused train step (same workload as bench.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sibrar_b200  # noqa
import bench
from sibrar_b200 import ops
from sibrar_b200.sbnet import SingleBranchNet
from sibrar_b200.synthetic import SynCorpus
from sibrar_b200.trainer import FusedTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
corpus = SynCorpus("ml1m", "cold_start_item", seed=42)
train = corpus.dataset("train")
torch.manual_seed(1234)
model = SingleBranchNet.build_from_conf(bench.ml1m_model_conf(), train).to(dev).train()
tr = FusedTrainer(model, bench.LEARN, n_negative_samples=bench.N_NEG)
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
for _ in range(20):
    tr.step(u, i)
b.record()
torch.cuda.synchronize()
print(f"B={B}: {a.elapsed_time(b) / 20:.3f} ms/step unprofiled")
with bench.CallProfiler(ops, torch) as prof:
    for _ in range(5):
        tr.step(u, i)
agg = prof.summary()
tot = sum(v[0] for v in agg.values()) / 5
print(f"sum of kernels {tot:.3f} ms/step, {sum(v[1] for v in agg.values()) / 5:.0f} calls/step")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0] / 5 / tot * 100:5.1f}%  {v[0] / v[1] * 1e3:8.1f} us x {v[1] / 5:4.1f}  {k}")
