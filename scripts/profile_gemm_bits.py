"""driver for ncu / timing: the bit-packed interaction projection (forward, split-K sliced) and its wgrad"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, torch, sibrar_b200
from sibrar_b200 import ops
rows, d, out = 3706, 6040, 64
m = sp.random(rows, d, density=0.045, format="csr", random_state=1); m.data[:] = 1
bits, bits_t = ops.pack_bits(m, "cuda"), ops.pack_bits(m.T.tocsr(), "cuda")
w = torch.randn(out, ops.pad8(d), device="cuda").to(torch.bfloat16)[:, :d]
dz = torch.randn(rows, out, device="cuda").to(torch.bfloat16)
split = ops.effective_splits(d, 10)
part = torch.empty((split, rows, out), device="cuda")
gw = torch.zeros(out, d, device="cuda")
def fwd(): ops.gemm_bits(bits, w, rows, out, d, out_f32=part.view(split * rows, out), split_k=split, split_stride=rows * out)
def wgrad(): ops.gemm_bits(bits_t, dz, d, out, rows, b_mn=True, out_f32=gw, transpose_out=True, atomic_out=True, split_k=6)
for name, fn in (("fwd", fwd), ("wgrad", wgrad)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / 20 * 1e3:.1f} us")
