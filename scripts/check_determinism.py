"""forward determinism: the logits of one golden step must be bit-identical whatever ran before / in another process"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, sibrar_b200
import tests.test_golden_gpu as T
from tests.golden_util import state_dict_of

def run(steps):
    spec, g, corpus, model = T._build("pairwise_bn2"); model.to("cuda").train(); tr = T._trainer(model, spec)
    for s in steps:
        T._load(model, state_dict_of(g, "sd0/") if s == 0 else state_dict_of(g, f"s{s-1}/sd/"))
        u, i, mods, keep = T._translate(model, g, s)
        for gr in tr.grads.values(): gr.zero_()
        tr.step(u, i, mods, keep, apply_optimizer=False); torch.cuda.synchronize()
    return tr.logits.cpu().numpy().copy()
a = run([0, 1, 2]); b = run([2]); c = run([2])
print("forward determinism: seq-vs-fresh", np.abs(a - b).max(), "fresh-vs-fresh", np.abs(c - b).max())
