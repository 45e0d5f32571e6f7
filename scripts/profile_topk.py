"""small driver for ncu / timing: fused score-GEMM + mask + top-k launches.
usage: profile_topk.py U I D k [reps]"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sibrar_b200
from sibrar_b200 import ops
U, I, D, k = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (37888, 100000, 64, 10))]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
g = torch.Generator().manual_seed(7)
u16 = (torch.randn(U, D, generator=g) / math.sqrt(D)).to(torch.bfloat16).cuda()
i16 = torch.randn(I, D, generator=g).to(torch.bfloat16).cuda()
ip = torch.arange(0, (U + 1) * 100, 100, dtype=torch.int64, device="cuda")
ix = torch.sort(torch.randint(0, I, (U, 100), device="cuda", dtype=torch.int32), dim=1).values.reshape(-1).contiguous()
best = 1e9
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    vals, idx = ops.topk_scores_masked(u16, i16, U, I, D, *( (None, None) if os.environ.get("NOSEEN") else (ip, ix)), k)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print(f"U={U} I={I} D={D} k={k} dbg={os.environ.get('SBR_TOPK_DEBUG', '0')} cap={os.environ.get('SBR_TOPK_CAP', '-')} noseen={os.environ.get('NOSEEN', '0')}: {best:.3f} ms  "
      f"{2.0*U*I*D/best/1e9:.1f} TFLOP/s  {U*I/best/1e6:.1f} Gscores/s", flush=True)
