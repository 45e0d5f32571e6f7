"""Helpers shared by the golden-fixture tests (CPU oracle tests and GPU parity tests)."""
import os

import numpy as np

from oracle.make_golden import CASES, GOLDEN_DIR  # noqa: F401  (CASES only; nothing of the reference is imported)


def load_case(name):
    from sibrar_b200.synthetic import SynCorpus
    spec = CASES[name]
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False))
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
