"""driver for ncu / timing: the bit-packed interaction projection (forward, split-K sliced) and its wgrad at the ML-1M
shapes and splits of the train step.  SBR_LIB_PATH selects a diagnostic build (scripts/gemm_bits_variants.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, torch, sibrar_b200
from sibrar_b200 import ops
rows, d, out = 3706, 6040, 64
m = sp.random(rows, d, density=0.045, format="csr", random_state=1); m.data[:] = 1
bits, bits_t = ops.pack_bits(m, "cuda"), ops.pack_bits(m.T.tocsr(), "cuda")
w = torch.randn(out, ops.pad8(d), device="cuda").to(torch.bfloat16)[:, :d]
dz = torch.randn(rows, out, device="cuda").to(torch.bfloat16)
split = ops.effective_splits(d, int(os.environ.get("FWD_SPLIT", 5)))
wsplit = ops.effective_splits(rows, int(os.environ.get("WGRAD_SPLIT", 3)))
part = torch.empty((split, rows, out), device="cuda")
gw = torch.zeros(out, d, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def fwd(): ops.gemm_bits(bits, w, rows, out, d, out_f32=part.view(split * rows, out), split_k=split, split_stride=rows * out)
def wgrad(): ops.gemm_bits(bits_t, dz, d, out, rows, b_mn=True, out_f32=gw, transpose_out=True, atomic_out=True, split_k=wsplit)
for name, fn in (("fwd", fwd), ("wgrad", wgrad)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):  # cold L2, one launch per measurement
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name}: cold single launch {np.median(ts):.1f} us, 20 back to back {a.elapsed_time(b) / 20 * 1e3:.1f} us per launch")
