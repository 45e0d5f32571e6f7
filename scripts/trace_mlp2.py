"""Per-tile timeline of block 0 of sbr_mlp2_bwd (item side of the ML-1M step) from the in-kernel %globaltimer stamps
(SBR_MLP2_DEBUG bit 16).  Events: producer 1 tile start, 2 rows resolved, 3 X0 stage free, 4 X0 published, 5 first dz
loads issued, 6 DZ buffer free, 7 dz published; MMA 16 loop top, 17 accumulator free, 18 X0 ready, 19 Y1 ready, 20 dz
ready, 21 dY1 ready, 22 wgrad issued; epilogue 32 loop top, 33 acc(Y1) ready, 34 Y1 written, 35 acc(dY1) ready, 36 dY1
written, 37 acc(dX0) ready, 38 dX0 stored."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ONLY"] = "none"
import runpy
g = runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_mlp2.py"))
import torch
from sibrar_b200 import _lib
calls, orig_call = g["calls"], g["orig_call"]
key = ("sbr_mlp2_bwd", 180224)
buf = (C.c_ulonglong * 4096)()
_lib.lib().sbr_mlp2_trace_read(buf, 4096)
os.environ["SBR_MLP2_DEBUG"] = "16"
orig_call(key[0], *calls[key])
torch.cuda.synchronize()
n = _lib.lib().sbr_mlp2_trace_read(buf, 4096)
os.environ["SBR_MLP2_DEBUG"] = "0"
ev = sorted(((int(buf[i]) >> 8, int(buf[i]) & 0xff) for i in range(n)))
t0 = ev[0][0]
names = {1: "P tile start", 2: "P rows resolved", 3: "P X0 stage free", 4: "P X0 published", 5: "P dz loads issued",
         6: "P DZ buffer free", 7: "P dz published", 16: "M loop top", 17: "M acc free", 18: "M X0 ready",
         19: "M Y1 ready", 20: "M dz ready", 21: "M dY1 ready", 22: "M wgrad issued", 32: "E loop top",
         33: "E acc(Y1) ready", 34: "E Y1 written", 35: "E acc(dY1) ready", 36: "E dY1 written", 37: "E acc(dX0) ready",
         38: "E dX0 stored"}
print(f"{n} events, block 0, total {(ev[-1][0] - t0) / 1e3:.1f} us")
last = {}
for t, e in ev:
    role = "P" if e < 16 else ("M" if e < 32 else "E")
    d = t - last.get(role, t0)
    last[role] = t
    print(f"{(t - t0) / 1e3:8.2f} us  (+{d / 1e3:6.2f} in role)  {'      ' * ('PME'.index(role))}{names.get(e, e)}")
