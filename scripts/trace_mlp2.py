"""Where do the three roles of block 0 of sbr_mlp2_fwd / sbr_mlp2_bwd spend their time?  (item side of the ML-1M step)
SBR_MLP2_DEBUG bit 16 selects the PROF instantiation, which accumulates the SM cycles spent inside every barrier wait;
loop - sum(waits) is the role's own work.  Extra masks (e.g. EXTRA=32: no L2 prefetch) are OR-ed in."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ONLY"] = "none"
import runpy
g = runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_mlp2.py"))
import torch
from sibrar_b200 import _lib
calls, orig_call = g["calls"], g["orig_call"]
SITES = {
    "sbr_mlp2_fwd": {"P": {1: "wait X0 stage free", 8: "work resolve rows", 9: "work gather"},
                     "M": {1: "wait X0 ready", 2: "wait final acc drained", 3: "wait hidden act written", 8: "work issue"},
                     "E": {1: "wait hidden acc ready", 2: "wait final acc ready", 8: "work loop top", 9: "work hidden act",
                           10: "work z store + colstats"}},
    "sbr_mlp2_bwd": {"P": {1: "wait X0 stage free", 8: "work publish + resolve", 9: "work gather"},
                     "M": {2: "wait X0 ready", 3: "wait Y1 written", 4: "wait dz written", 5: "wait dY1 written", 8: "work issue"},
                     "E": {1: "wait acc(Y1) ready", 2: "wait acc(dY1) ready", 3: "wait acc(dX0) ready",
                           8: "work loop top", 9: "work Y1", 10: "work dY1", 11: "work dX0 store"},
                     "Z": {1: "wait DZ buffer free", 8: "work first dz loads", 9: "work dz"}},
}
extra = int(os.environ.get("EXTRA", "0"))
N = 64
buf = (C.c_ulonglong * N)()
for key in sorted(calls):
    if key[1] < 100000:
        continue
    os.environ["SBR_MLP2_DEBUG"] = str(16 | extra)
    for _ in range(2):
        orig_call(key[0], *calls[key])
    torch.cuda.synchronize()
    _lib.lib().sbr_mlp2_trace_read(buf, N)
    os.environ["SBR_MLP2_DEBUG"] = "0"
    tiles = (key[1] + 127) // 128
    per_cta = -(-tiles // min(tiles, 296))
    print(f"{key[0]} rows={key[1]} debug={16 | extra} (block 0: {per_cta} tiles), cycles per tile (~1900 cycles = 1 us)")
    for r, role in enumerate("PMEZ"):
        if role not in SITES[key[0].replace("_bn", "")]:
            continue
        v = [int(buf[r * 16 + i]) for i in range(16)]
        print(f"  {role}: loop {v[0] / per_cta:8.0f}   " + "   ".join(f"{name} {v[i] / per_cta:.0f}" for i, name in SITES[key[0].replace("_bn", "")][role].items()))
