"""Isolated timings of the item-side gather kernels of the train step (forward gather, sorted-run backward) under the
tuning knobs of gather.cu (SBR_SEG_RPG / SBR_SEG_BPS / SBR_SEG_MINB), on the arguments the real step passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sibrar_b200  # noqa
from sibrar_b200 import workloads
from sibrar_b200 import ops
from sibrar_b200.sbnet import SingleBranchNet
from sibrar_b200.synthetic import SynCorpus
from sibrar_b200.trainer import FusedTrainer

B = 16384
dev = torch.device("cuda", 0)
corpus, _conf, _learn, _, _ = workloads.build("ml1m")
train = corpus.dataset("train")
torch.manual_seed(1234)
model = SingleBranchNet.build_from_conf(_conf, train).to(dev).train()
tr = FusedTrainer(model, _learn, n_negative_samples=workloads.N_NEG)
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
    if name in ("sbr_row_gather_bwd_segmented", "sbr_row_gather_fwd"):
        rows = [x for x in a if isinstance(x, int) and x >= 100000]
        if rows:
            calls[name] = a
    return orig_call(name, *a)
ops.call = spy
tr.step(u, i)
torch.cuda.synchronize()
ops.call = orig_call
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(name, label, reps=10):
    a = calls[name]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for r in range(reps + 2):
        flush.zero_()
        if r >= 2:
            ev[r - 2][0].record()
        orig_call(name, *a)
        if r >= 2:
            ev[r - 2][1].record()
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) * 1e3 for x, y in ev)
    print(f"{label:50s} median {t[len(t) // 2]:7.1f} us  min {t[0]:7.1f} us", flush=True)

timeit("sbr_row_gather_fwd", "forward gather")
for minb, bps in ((3, 3), (3, 4), (4, 4)):
    for rpg in (8, 12, 0):
        os.environ.update(SBR_SEG_MINB=str(minb), SBR_SEG_BPS=str(bps), SBR_SEG_RPG=str(rpg))
        timeit("sbr_row_gather_bwd_segmented",
               f"seg_reduce min blocks/SM={minb} blocks/SM={bps} rows/chunk={rpg or 'balanced'}")
