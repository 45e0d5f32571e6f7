"""User-side (in-batch) InfoNCE at the train batch: tensor-core route (ops.infonce_gemm) vs the CUDA-core kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sibrar_b200  # noqa
from sibrar_b200 import ops

dev = "cuda"
for n, D in ((4096, 64), (16384, 64), (16384, 128)):
    e = torch.randn(1, n, 2, D, device=dev)
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    de = torch.zeros_like(e)
    for route in ("1", "0"):
        if route == "0" and n > 4096:
            continue  # (the O(n^2 D) CUDA-core kernels take seconds at n = 16 384)
        os.environ["SBR_INFONCE_GEMM"] = route
        for _ in range(2):
            ops.infonce(e, 1, n, D, 0.5, 1.0, acc, de, accumulate=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            ops.infonce(e, 1, n, D, 0.5, 1.0, acc, de, accumulate=False)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        flops = 2.0 * n * n * (3 * D + 2 * D)
        print(f"n={n} D={D} route={'gemm' if route == '1' else 'cuda-core'}: {ms * 1e3:9.1f} us  "
              f"({flops / ms / 1e9:.0f} TFLOP/s of GEMM work, {6.0 * n * n * 2 / ms / 1e6:.0f} GB/s of n^2 traffic)", flush=True)
