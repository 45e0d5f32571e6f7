"""debug helper (GPU): per-tensor relative errors of one golden case / step against the reference fixture and the
numpy oracle's intermediates"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sibrar_b200
from oracle import sbnet_oracle as O
import tests.test_golden_gpu as T
from tests.golden_util import state_dict_of, step_inputs

name = sys.argv[1] if len(sys.argv) > 1 else "pairwise_bn2"
s = int(sys.argv[2]) if len(sys.argv) > 2 else 0
spec, g, corpus, model = T._build(name)
model.to("cuda").train()
tr = T._trainer(model, spec)
sd = state_dict_of(g, "sd0/") if s == 0 else state_dict_of(g, f"s{s-1}/sd/")
T._load(model, sd)
u, i, mods, keep = T._translate(model, g, s)

# oracle intermediates
net = O.OracleSBNet(spec["model"], corpus.dataset("train"))
p = {k: v.astype(np.float64) if v.dtype.kind == "f" else v for k, v in sd.items()}
uu, ii, omods, onames, odrop = step_inputs(g, s)
r = net.train_step_fwd_bwd(p, uu, ii, omods, onames, odrop, loss_kind=spec["rec_loss"], n_items=corpus.n_items, neg_train=spec["n_neg"])
net2 = O.OracleSBNet(spec["model"], corpus.dataset("train"))
r2 = net2.train_step_fwd_bwd(p, uu, ii, omods, onames, odrop, loss_kind=spec["rec_loss"], n_items=corpus.n_items, neg_train=spec["n_neg"], emu=O.Bf16Emulation())

# run the CUDA step piecewise
rt = tr.rt
from sibrar_b200 import ops
ops.tick(rt.step_dev); rt.arena.reset()
Eu = tr.user.embed(u, True, mods.get("user"), keep.get("user"))
Ei = tr.item.embed(i, True, mods.get("item"), keep.get("item"))
def rel(a, b):
    a = np.asarray(a, np.float64).reshape(-1); b = np.asarray(b, np.float64).reshape(-1)
    return float(np.abs(a-b).max() / max(1e-12, np.abs(b).max()))
for ent_name, E in (("user", Eu), ("item", Ei)):
    ent = net.ent[ent_name]
    if hasattr(ent, "cache"):
        print(ent_name, "E rel", rel(E.cpu().numpy(), ent.cache["E"]), "scale", np.abs(ent.cache["E"]).max())
for g_ in tr.grads.values(): g_.zero_()
tr.read_losses()
tr.step(u, i, mods, keep, apply_optimizer=False)
torch.cuda.synchronize()
print("losses", tr.read_losses(), {k: float(g[f"s{s}/{k}"]) for k in ("rec_loss", "reg_loss", "loss")})
gold = state_dict_of(g, f"s{s}/grad/")
params = dict(model.named_parameters())
for k, gg in gold.items():
    got = tr.grads[id(params[k])].cpu().numpy()
    og = r2["grads"].get(k)
    cos = float((got.reshape(-1) @ gg.reshape(-1)) / max(1e-30, np.linalg.norm(got) * np.linalg.norm(gg)))
    print(f"{k[-70:]:70s} max {np.abs(gg).max():9.3e} err/max {rel(got, gg):8.2e} cos {cos:7.4f} | emu: cuda-vs-emu {rel(got, og) if og is not None else -1:8.2e} emu-vs-gold {rel(og, gg) if og is not None else -1:8.2e}")
print("logits cuda-vs-emu", rel(tr.logits.cpu().numpy(), r2["logits"]), "emu-vs-gold", rel(r2["logits"], g[f"s{s}/logits"]), "loss emu", r2["loss"])
