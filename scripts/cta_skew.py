"""Start skew / tail of the persistent grid of the item-side sbr_mlp2_bwd INSIDE the real ML-1M step (both entities'
branches running): SBR_MLP2_DEBUG=64 makes every CTA stamp %globaltimer at its start and end."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SBR_MLP2_DEBUG"] = "64"
import numpy as np, torch
import sibrar_b200  # noqa
from sibrar_b200 import _lib, workloads
from sibrar_b200.sbnet import SingleBranchNet
from sibrar_b200.synthetic import sample_batch
from sibrar_b200.trainer import FusedTrainer

dev = "cuda"
corpus, conf, learn, _, _ = workloads.build("ml1m")
train = corpus.dataset("train")
model = SingleBranchNet.build_from_conf(conf, train).to(dev).train()
tr = FusedTrainer(model, learn, n_negative_samples=workloads.N_NEG, cuda_graph=True)
rng = np.random.default_rng(0)
u, i = sample_batch(train, 16384, rng, workloads.N_NEG)
u, i = torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev)
for _ in range(6):
    tr.step(u, i)
torch.cuda.synchronize()
buf = (C.c_ulonglong * (2 * 296))()
_lib.lib().sbr_mlp2_cta_times_read(buf, 296)
t = np.array(buf, dtype=np.int64).reshape(296, 2)
t0 = t[:, 0].min()
s, e = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
print(f"item mlp2_bwd inside the graph step: grid spans {e.max():.1f} us; CTA starts: median {np.median(s):.1f} us, "
      f"p90 {np.percentile(s, 90):.1f}, max {s.max():.1f}; CTA durations: median {np.median(e - s):.1f} us, "
      f"min {(e - s).min():.1f}, max {(e - s).max():.1f}")
late = s > 5
print(f"{late.sum()} of 296 CTAs start more than 5 us after the first one; their mean start {s[late].mean() if late.any() else 0:.1f} us, "
      f"mean end {e[late].mean() if late.any() else 0:.1f} us; early CTAs mean end {e[~late].mean():.1f} us")
