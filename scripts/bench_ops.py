"""Micro-timings of single C-ABI ops at the train step's shapes (back-to-back launches, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sibrar_b200
from sibrar_b200 import ops
dev = "cuda"
BF16, F32 = torch.bfloat16, torch.float32

def timeit(name, fn, nbytes, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1e3
    print(f"{name:60s} {us:8.1f} us  {nbytes / us / 1e3:8.1f} GB/s ({nbytes / 1e6:.1f} MB)", flush=True)

M = int(sys.argv[1]) if len(sys.argv) > 1 else 180224
x = torch.randn(M, 64, device=dev).to(BF16)
w = torch.randn(64, 64, device=dev).to(BF16)
bias = torch.randn(64, device=dev)
y16 = torch.empty(M, 64, dtype=BF16, device=dev)
y32 = torch.empty(M, 64, dtype=F32, device=dev)
stats = torch.zeros(128, device=dev)
timeit("gemm M x64x64 -> bf16 (+bias+relu)", lambda: ops.gemm(x, w, M, 64, 64, bias=bias, act="relu", out_bf16=y16), M * 64 * 4)
timeit("gemm M x64x64 -> bf16 (plain)", lambda: ops.gemm(x, w, M, 64, 64, out_bf16=y16), M * 64 * 4)
timeit("gemm M x64x64 -> f32 (plain)", lambda: ops.gemm(x, w, M, 64, 64, out_f32=y32), M * 64 * 6)
timeit("gemm M x64x64 -> f32 + colstats", lambda: ops.gemm(x, w, M, 64, 64, bias=bias, out_f32=y32, colstats=stats), M * 64 * 6)
timeit("gemm dgrad b_mn -> f32", lambda: ops.gemm(x, w, M, 64, 64, b_mn=True, out_f32=y32), M * 64 * 6)
gw = torch.zeros(64, 64, device=dev)
for split in (1, 37, 148, 296):
    timeit(f"gemm wgrad 64x64xM split_k={split}", lambda: ops.gemm(x, y16, 64, 64, M, a_mn=True, b_mn=True, out_f32=gw, transpose_out=True, atomic_out=True, split_k=split), M * 64 * 4)
timeit("torch copy bf16->bf16 (reference point)", lambda: y16.copy_(x), M * 64 * 4)
timeit("torch matmul bf16", lambda: torch.matmul(x, w.t(), out=y16), M * 64 * 4)

# ---- small table-level kernels
rows = 3706
dy = torch.randn(rows, 64, device=dev)
yt = torch.randn(rows, 64, device=dev)
o16 = torch.empty(rows, 64, dtype=BF16, device=dev)
cs = torch.zeros(64, device=dev)
timeit("actgrad_colsum [3706 x 64] (+zero_dy, bf16 out, colsum)", lambda: ops.actgrad_colsum(dy, yt, "relu", rows, 64, out_bf16=o16, colsum=cs, zero_dy=True), rows * 64 * 14)
big = torch.randn(M, 64, device=dev)
mi = torch.cat([torch.zeros(64, device=dev), torch.ones(64, device=dev)])
g1 = torch.ones(64, device=dev)
sums = torch.zeros(128, device=dev)
timeit("bn_apply [M x 64] f32 -> f32", lambda: ops.bn_apply(big, mi, g1, g1, None, M, 64, out_f32=y32), M * 64 * 8)
timeit("bn_bwd_reduce [M x 64]", lambda: ops.bn_bwd_reduce(big, None, None, y32, mi, M, 64, sums), M * 64 * 8)
timeit("bn_bwd_apply [M x 64] -> bf16", lambda: ops.bn_bwd_apply(big, None, None, y32, mi, g1, sums, M, 64, dz_bf16=y16), M * 64 * 10)
