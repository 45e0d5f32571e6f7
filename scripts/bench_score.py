"""Isolated timing of the fused score / loss / BatchNorm kernel and the BatchNorm backward at the train step's item
shape (B = 16384, n = 11, D = 64), L2 flushed before every launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sibrar_b200  # noqa
from sibrar_b200 import ops
dev = "cuda"
B, n, D = 16384, 11, 64
g = torch.Generator(device="cpu").manual_seed(0)
zu = torch.randn(B, D, generator=g).to(dev)
zi = torch.randn(B * n, D, generator=g).to(dev)
mi = torch.cat([torch.zeros(D), torch.ones(D)]).to(dev)
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
logits = torch.empty(B, n, device=dev)
loss = torch.zeros(1, dtype=torch.float64, device=dev)
deu, dei = torch.empty_like(zu), torch.empty_like(zi)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def bn(z):
    return dict(z=z, mean_invstd=mi, gamma=gamma, beta=beta,
                sums=torch.zeros(ops.BN_SUM_REPLICAS * 2 * D, device=dev))
bu, bi = bn(zu), bn(zi)

def timeit(label, fn, nbytes, reps=10):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for r in range(reps + 2):
        flush.zero_()
        if r >= 2:
            ev[r - 2][0].record()
        fn()
        if r >= 2:
            ev[r - 2][1].record()
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) * 1e3 for x, y in ev)
    med = t[len(t) // 2]
    print(f"{label:56s} median {med:7.1f} us  min {t[0]:7.1f} us  {nbytes / med / 1e3:7.0f} GB/s", flush=True)

nb = 2 * B * D * 4 * (1 + n) + B * n * 4
for generic in (True, False):
    if generic:
        os.environ["SBR_SCORE_GENERIC"] = "1"
    else:
        os.environ.pop("SBR_SCORE_GENERIC", None)
    timeit(f"score_loss_bn bpr ({'loop' if generic else 'registers'})",
           lambda: ops.score_loss_bn(None, bu, None, bi, B, n, D, "bpr", False, 0.0, logits, loss, deu, dei), nb)
rows = B * n
dz16 = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
sums = torch.zeros(ops.BN_SUM_REPLICAS * 2 * D, device=dev)
for scalar in (True, False):
    if scalar:
        os.environ["SBR_NORM_SCALAR"] = "1"
    else:
        os.environ.pop("SBR_NORM_SCALAR", None)
    tag = "scalar" if scalar else "vector"
    timeit(f"bn_bwd_apply [180224 x 64] -> bf16 ({tag})",
           lambda: ops.bn_bwd_apply(dei, None, None, zi, mi, gamma, sums, rows, D, dz_bf16=dz16, dgamma=dg, dbeta=db,
                                    n_replicas=ops.BN_SUM_REPLICAS), rows * D * 10)
    t_rows = 3706
    dyt = torch.randn(t_rows, D, device=dev)
    yt = torch.randn(t_rows, D, device=dev)
    o16 = torch.empty(t_rows, D, dtype=torch.bfloat16, device=dev)
    cs = torch.zeros(D, device=dev)
    timeit(f"actgrad_colsum [3706 x 64] relu, zero_dy, bf16 out ({tag})",
           lambda: ops.actgrad_colsum(dyt, yt, "relu", t_rows, D, out_bf16=o16, colsum=cs, zero_dy=True),
           t_rows * D * 14)
