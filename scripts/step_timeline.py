"""Timeline of ONE replayed CUDA-graph train step (ML-1M shape, B = 16 384) from the timeline build of the library
(`make -C sibrar---single-branch-recommender_b200/csrc stamps`; SBR_LIB_PATH selects it): every kernel records
%globaltimer right behind its griddepcontrol.wait, i.e. when its inputs are ready.  Prints the kernels in time order with
the gap to the previous stamp; kernel names are recovered from (source file, line of the PDL macro)."""
import ctypes as C, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "sibrar---single-branch-recommender_b200", "csrc")
os.environ.setdefault("SBR_LIB_PATH", os.path.join(CSRC, "libsibrar_b200_stamps.so"))
import numpy as np, torch
import sibrar_b200  # noqa
from sibrar_b200 import _lib, workloads
from sibrar_b200.sbnet import SingleBranchNet
from sibrar_b200.synthetic import sample_batch
from sibrar_b200.trainer import FusedTrainer


def fid(path):
    h = 2166136261
    for ch in path.encode():
        h = ((h ^ ch) * 16777619) & 0xffffffff
    return h & 0xffff


def kernel_names():
    """(file id, line) -> name of the enclosing __global__ function"""
    out = {}
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith(".cu"):
            continue
        lines = open(os.path.join(CSRC, f)).read().split("\n")
        name = "?"
        for n, line in enumerate(lines, 1):
            m = re.search(r"([A-Za-z_0-9]+_kernel)\(", line) if "<<<" not in line and "sbr_launch" not in line else None
            if m:
                name = m.group(1)
            if "SBR_PDL_ENTRY()" in line or "SBR_PDL_WAIT()" in line:
                out[(fid(f), n)] = name
    return out


dev = "cuda"
B = int(os.environ.get("B", 16384))
corpus, conf, learn, _, _ = workloads.build("ml1m")
train = corpus.dataset("train")
model = SingleBranchNet.build_from_conf(conf, train).to(dev).train()
tr = FusedTrainer(model, learn, n_negative_samples=workloads.N_NEG, cuda_graph=True)
rng = np.random.default_rng(0)
u, i = sample_batch(train, B, rng, workloads.N_NEG)
u, i = torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev)
for _ in range(6):
    tr.step(u, i)
torch.cuda.synchronize()
buf = torch.zeros(1 + 2 * 4000, dtype=torch.int64, device=dev)
lib = _lib.lib()
lib.sbr_debug_stamps.argtypes = [C.c_void_p]
assert lib.sbr_debug_stamps(buf.data_ptr()) == 0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
best = None
for rep in range(3):
    flush.zero_()
    buf.zero_()
    torch.cuda.synchronize()
    tr.step(u, i)
    torch.cuda.synchronize()
    h = buf.cpu().numpy()
    n = int(h[0])
    ev = sorted((int(h[1 + 2 * k]), int(h[2 + 2 * k])) for k in range(n))
    span = (ev[-1][0] - ev[0][0]) / 1e3
    if best is None or span < best[0]:
        best = (span, ev)
lib.sbr_debug_stamps(None)
names = kernel_names()
span, ev = best
t0 = ev[0][0]
print(f"{len(ev)} kernels, first stamp -> last stamp {span:.1f} us (B = {B}; the last kernel's own run time is not included)")
prev = t0
for t, tag in ev:
    grid, tag = tag >> 32, tag & 0xffffffff
    nm = names.get((tag >> 16, tag & 0xffff), f"file {tag >> 16:#x} line {tag & 0xffff}")
    print(f"{(t - t0) / 1e3:8.1f} us  (+{(t - prev) / 1e3:6.1f})  {nm}  grid {grid}")
    prev = t
