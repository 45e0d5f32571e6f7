"""Helpers shared by the golden-fixture tests (CPU oracle tests and GPU parity tests)."""
import os

import numpy as np

from oracle.make_golden import CASES, GOLDEN_DIR, seeded_state, unpack_f16  # noqa: F401  (nothing of the reference)


def load_case(name):
    from sibrar_b200.synthetic import SynCorpus
    spec = CASES[name]
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False))
    if spec.get("seeded_init") is not None:
        # compact fixture: weights regenerated from (name, shape, seed), gradients stored as scaled fp16, and of the
        # post-step state only the BatchNorm running statistics (the optimizer is pinned by the small cases)
        shapes = {k[len("shape/"):]: v for k, v in g.items() if k.startswith("shape/")}
        sd0 = seeded_state(shapes, spec["seeded_init"])
        g.update({f"sd0/{k}": v for k, v in sd0.items()})
        for s in range(spec["steps"]):
            for k in [k for k in g if k.startswith(f"s{s}/grad16/")]:
                pname = k[len(f"s{s}/grad16/"):]
                g[f"s{s}/grad/{pname}"] = unpack_f16(g[k], g[f"s{s}/gscale/{pname}"])
            for k, v in sd0.items():
                g.setdefault(f"s{s}/sd/{k}", v)
    corpus = SynCorpus(**spec["corpus"])
    return spec, g, corpus


def state_dict_of(g, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in g.items() if k.startswith(prefix)}


def step_inputs(g, s):
    mods, names, drop = {}, {}, {}
    for ent in ("user", "item"):
        if f"s{s}/mods_{ent}" in g:
            mods[ent] = g[f"s{s}/mods_{ent}"].astype(np.int64)
            names[ent] = [str(x) for x in g[f"s{s}/mod_names_{ent}"]]
        if f"s{s}/drop_{ent}" in g:
            drop[ent] = g[f"s{s}/drop_{ent}"].astype(np.float32)
    return g[f"s{s}/u"], g[f"s{s}/i"], mods, names, drop
