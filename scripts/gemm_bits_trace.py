"""per-CTA phase trace of the bit-packed GEMM (csrc built with -DSBR_GB_TRACE as csrc/variants/lib_trace.so):
clock64 offsets from the kernel entry (cycles) at the points the kernel stamps, %globaltimer at entry / exit (ns)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("SBR_LIB_PATH", os.path.join(ROOT, "sibrar---single-branch-recommender_b200", "csrc", "variants", "lib_trace.so"))
sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp, torch, sibrar_b200
from sibrar_b200 import ops, _lib
rows, d, out = 3706, 6040, 64
m = sp.random(rows, d, density=0.045, format="csr", random_state=1); m.data[:] = 1
bits = ops.pack_bits(m, "cuda")
w = torch.randn(out, ops.pad8(d), device="cuda").to(torch.bfloat16)[:, :d]
split = ops.effective_splits(d, 5)
part = torch.empty((split, rows, out), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def fwd(): ops.gemm_bits(bits, w, rows, out, d, out_f32=part.view(split * rows, out), split_k=split, split_stride=rows * out)
NAMES = {1: "setup done", 2: "PDL wait done", 11: "first B TMA issued", 9: "first bit group landed", 10: "first A stage filled",
         3: "MMA: K block 0 ready", 4: "MMA: K block 4 ready", 13: "MMA: K block 8 ready", 14: "MMA: K block 12 ready",
         5: "MMA: all issued", 6: "epilogue: accumulator ready", 7: "epilogue: stored", 8: "all warps done"}
for cold in (True, False):
    for _ in range(3): fwd()
    torch.cuda.synchronize()
    if cold: flush.zero_()
    fwd(); torch.cuda.synchronize()
    buf = np.zeros(1024 * 16, dtype=np.uint64)
    rc = _lib.lib().sbr_debug_gemm_bits_trace(buf.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    t = buf.reshape(1024, 16)[:29 * split].astype(np.int64)
    g0 = t[:, 0].min()
    print(f"== {'cold L2' if cold else 'warm L2'}: {len(t)} CTAs; entry spread {(t[:, 0].max() - g0) / 1e3:.2f} us, "
          f"last exit {(t[:, 12].max() - g0) / 1e3:.2f} us after the first entry; CTA lifetime median {np.median(t[:, 12] - t[:, 0]) / 1e3:.2f} us")
    for k in (1, 2, 11, 9, 10, 3, 4, 13, 14, 5, 6, 7, 8):
        v = t[:, k]
        print(f"  {NAMES[k]:30s} median {np.median(v):8.0f}  min {v.min():8d}  max {v.max():8d} cycles")
