"""Isolated timings of the fused gather + SB-MLP kernels (sbr_mlp2_fwd / sbr_mlp2_bwd) on the arguments the real ML-1M
step passes (item side: 180 224 rows, user side: 16 384 rows), with the SBR_MLP2_DEBUG attribution switches
(1 = no gradient flush, 2 = no gather loads, 4 = no dy / z loads, 8 = no output stores, 32 = no L2 prefetch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sibrar_b200  # noqa
from sibrar_b200 import ops, workloads
from sibrar_b200.sbnet import SingleBranchNet
from sibrar_b200.trainer import FusedTrainer

B = int(os.environ.get("B", 16384))
dev = torch.device("cuda", 0)
corpus, conf, learn, _, _ = workloads.build("ml1m")
train = corpus.dataset("train")
torch.manual_seed(1234)
model = SingleBranchNet.build_from_conf(conf, train).to(dev).train()
tr = FusedTrainer(model, learn, n_negative_samples=10)
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
for _ in range(3):
    tr.step(u, i)
torch.cuda.synchronize()
calls = {}
orig_call = ops.call
def spy(name, *a):
    if name in ("sbr_mlp2_fwd", "sbr_mlp2_fwd_bn", "sbr_mlp2_bwd"):
        calls[(name, int(a[1]))] = a
    return orig_call(name, *a)
ops.call = spy
tr.step(u, i, apply_optimizer=False)
torch.cuda.synchronize()
ops.call = orig_call
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(key, label, reps=10):
    a = calls[key]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for r in range(reps + 2):
        flush.zero_()
        if r >= 2:
            ev[r - 2][0].record()
        orig_call(key[0], *a)
        if r >= 2:
            ev[r - 2][1].record()
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) * 1e3 for x, y in ev)
    print(f"{label:60s} median {t[len(t) // 2]:7.1f} us  min {t[0]:7.1f} us", flush=True)

os.environ["SBR_MLP2_DEBUG"] = "0"
for key in sorted(calls):
    timeit(key, f"{key[0]} rows={key[1]}")
only = os.environ.get("ONLY")  # e.g. "sbr_mlp2_bwd:180224:0" = one call, one debug mask (ncu captures)
for key in sorted(calls):
    for dbg in (0, 1, 2, 4, 8, 15, 32):
        if only and only != f"{key[0]}:{key[1]}:{dbg}":
            continue
        os.environ["SBR_MLP2_DEBUG"] = str(dbg)
        timeit(key, f"{key[0]} rows={key[1]} debug={dbg}")
os.environ["SBR_MLP2_DEBUG"] = "0"
