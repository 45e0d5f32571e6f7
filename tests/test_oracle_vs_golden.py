"""Pins the numpy oracle (oracle/sbnet_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py): logits, losses, every gradient, parameters after each optimizer step, BN running
statistics, eval representations, masked scores, top-k and per-user metrics."""
import numpy as np
import pytest

from oracle import sbnet_oracle as O
from tests.golden_util import CASES, load_case, state_dict_of, step_inputs


WIDE = {"fp32_noise": 0.0}  # set per case: fp32 round-off of the reference relative to max-abs (wide layers)


def _close(a, b, rtol=2e-4, atol=2e-6, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    # the reference accumulates K = 512 .. 1200 products in fp32: entries near zero carry round-off proportional to
    # the tensor's scale, not to their own value
    tol = atol + rtol * np.abs(b) + WIDE["fp32_noise"] * (np.abs(b).max() if b.size else 0.0)
    assert (err <= tol).all(), f"{what}: max err {err.max():.3e} (ref max {np.abs(b).max():.3e})"


@pytest.mark.parametrize("name", list(CASES))
def test_train_steps_match_reference(name):
    spec, g, corpus = load_case(name)
    WIDE["fp32_noise"] = 2e-5 if spec.get("seeded_init") is not None else 0.0
    ds = corpus.dataset("train")
    net = O.OracleSBNet(spec["model"], ds)
    p = {k: v.astype(np.float64) if v.dtype.kind == "f" else v for k, v in state_dict_of(g, "sd0/").items()}
    state = {}
    for s in range(spec["steps"]):
        u, i, mods, names, drop = step_inputs(g, s)
        r = net.train_step_fwd_bwd(p, u, i, mods, names, drop, loss_kind=spec["rec_loss"], n_items=ds.n_items,
                                   neg_train=spec["n_neg"])
        _close(r["logits"], g[f"s{s}/logits"], what=f"s{s} logits")
        _close(r["rec_loss"], g[f"s{s}/rec_loss"], what=f"s{s} rec_loss")
        _close(r["reg_loss"], g[f"s{s}/reg_loss"], atol=1e-6, what=f"s{s} reg_loss")
        _close(r["loss"], g[f"s{s}/loss"], what=f"s{s} loss")
        gold_grads = state_dict_of(g, f"s{s}/grad/")
        gscale = max(1.0, max(float(np.abs(v).max()) for v in gold_grads.values()))
        for k, gg in gold_grads.items():
            got = r["grads"].get(k, np.zeros_like(gg))
            # atol: fp32 round-off of the reference (e.g. a bias in front of a BatchNorm has an exactly-zero
            # gradient; the reference holds ~1e-6 noise there)
            # (real-width fixtures store the gradients as scaled fp16: 2^-12 of each value, + the reference's fp32 noise)
            extra = 1.5e-3 * float(np.abs(gg).max()) if spec.get("seeded_init") is not None else 0.0
            _close(got, gg, rtol=1e-3, atol=3e-5 * gscale + extra, what=f"s{s} grad {k}")
        # the optimizer restatement is checked on the reference's own gradients: Adam turns the fp32 noise of
        # exactly-zero gradients into +-lr updates, which no independent implementation can reproduce
        grads = {k: v.astype(np.float64) for k, v in gold_grads.items()}
        if spec.get("seeded_init") is None:  # (the real-width fixtures do not apply the optimizer, see make_golden.py)
            O.adam_step(p, grads, state, spec["lr"], spec["wd"], s + 1, decoupled=spec["optimizer"] == "adamw")
        p.update(r["new_stats"])
        for k, v in state_dict_of(g, f"s{s}/sd/").items():
            _close(p[k], v, rtol=1e-3, atol=2e-5, what=f"s{s} param {k}")


@pytest.mark.parametrize("name", list(CASES))
def test_eval_matches_reference(name):
    spec, g, corpus = load_case(name)
    WIDE["fp32_noise"] = 2e-5 if spec.get("seeded_init") is not None else 0.0
    net = O.OracleSBNet(spec["model"], corpus.dataset("train"))
    val = corpus.dataset("val")
    p = state_dict_of(g, f"s{spec['steps'] - 1}/sd/")
    i_repr, _ = net.represent("item", val.items_in_split, p, False)
    u_repr, _ = net.represent("user", val.users_in_split, p, False)
    _close(i_repr, g["eval/i_repr"], rtol=1e-3, atol=1e-5, what="i_repr")
    _close(u_repr, g["eval/u_repr"], rtol=1e-3, atol=1e-5, what="u_repr")
    ex = val.exclude_data[val.users_in_split]
    vals, idx = O.masked_topk(g["eval/u_repr"], g["eval/i_repr"], ex, 5)
    gold_idx, gold_val = g["eval/topk_idx"], g["eval/topk_val"]
    _close(np.where(np.isinf(vals), -1e30, vals), np.where(np.isinf(gold_val), -1e30, gold_val), rtol=1e-4,
           atol=1e-6, what="topk values")
    # indices: identical wherever the reference's ranking is tie-free (torch.topk tie order is unspecified)
    s = np.sort(g["eval/scores"], axis=1)[:, ::-1][:, :6]
    tie_free = (np.abs(np.diff(s, axis=1)) > 1e-6 * np.maximum(1, np.abs(s[:, :5]))).all(1)
    assert tie_free.mean() > 0.9
    assert (idx[tie_free] == gold_idx[tie_free]).all()
    # metrics from the reference's own top-k must reproduce the reference's per-user metric vectors
    tgt = val.user_sampling_matrix[val.users_in_split][:, val.items_in_split]
    m = O.metrics_at_k(gold_idx, tgt, [1, 3, 5], n_items=val.n_items_in_split)
    for key, v in state_dict_of(g, "eval/raw/").items():
        _close(m[key], v, rtol=1e-5, atol=1e-6, what=f"raw {key}")
    for key, v in state_dict_of(g, "eval/metric/").items():
        got = m[key] if key.startswith("coverage") else m[key].mean()
        _close(got, v, rtol=1e-5, atol=1e-6, what=f"metric {key}")


def test_metric_restatement_pinned_against_reference_metrics():
    """ndcg / recall / precision of ``oracle/rmet_restated.py`` and of ``O.metrics_at_k`` against vectors produced by the
    reference's own ``eval/metrics.py:4-105`` (``oracle/make_metrics_golden.py``): users without targets, with more
    targets than k, with every item a target."""
    import os

    import scipy.sparse as sp
    import torch

    from oracle import rmet_restated as R
    from tests.golden_util import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "metrics_pin.npz"))
    logits, targets = torch.from_numpy(g["logits"]), torch.from_numpy(g["targets"].astype(np.float32))
    ks = [int(k) for k in g["ks"]]
    res, best = R.calculate(["ndcg", "recall", "precision"], logits, targets, k=ks, return_individual=True,
                            return_best_logit_indices=True)
    assert np.array_equal(best.numpy(), g[f"topk@{max(ks)}"])  # tie-free logits: torch.topk is unambiguous
    tgt = sp.csr_matrix(g["targets"])
    m = O.metrics_at_k(g[f"topk@{max(ks)}"], tgt, ks, n_items=targets.shape[1])
    for k in ks:
        for name in ("ndcg", "recall", "precision"):
            want = g[f"{name}@{k}"]
            assert np.abs(res[f"{name}_individual@{k}"].numpy() - want).max() < 1e-6, (name, k)
            assert np.abs(m[f"{name}@{k}"] - want).max() < 1e-6, (name, k)
