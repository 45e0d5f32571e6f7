"""small driver for ncu: one fused score-GEMM + mask + top-k launch"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sibrar_b200
from sibrar_b200 import ops
U, I, D, k = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (37888, 100000, 64, 10))]
g = torch.Generator().manual_seed(7)
u16 = (torch.randn(U, D, generator=g) / math.sqrt(D)).to(torch.bfloat16).cuda()
i16 = torch.randn(I, D, generator=g).to(torch.bfloat16).cuda()
ip = torch.arange(0, (U + 1) * 100, 100, dtype=torch.int64, device="cuda")
ix = torch.sort(torch.randint(0, I, (U, 100), device="cuda", dtype=torch.int32), dim=1).values.reshape(-1).contiguous()
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    vals, idx = ops.topk_scores_masked(u16, i16, U, I, D, ip, ix, k)
    b.record()
    torch.cuda.synchronize()
    print(f"U={U} I={I} D={D} k={k}: {a.elapsed_time(b):.3f} ms  {2.0*U*I*D/a.elapsed_time(b)/1e9:.1f} TFLOP/s")
