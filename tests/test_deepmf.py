"""DeepMatrixFactorization (algorithms/sgd_alg.py:1141-1242): the CPU oracle is pinned against fixtures generated from the
unmodified reference (``oracle/make_golden_deepmf.py``); the B200 path (``sibrar_b200.deepmf``) is compared with both."""
import os

import numpy as np
import pytest
import torch

from oracle.deepmf_oracle import OracleDeepMF
from oracle.make_golden import GOLDEN_DIR
from oracle.make_golden_deepmf import CASES

DEV = "cuda:0"


def _load(name):
    from sibrar_b200.synthetic import SynCorpus
    spec = CASES[name]
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))
    return spec, g, SynCorpus(**spec["corpus"])


def _sd(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


@pytest.mark.parametrize("name", list(CASES))
def test_deepmf_oracle_matches_reference_fixture(name):
    spec, g, corpus = _load(name)
    net = OracleDeepMF(spec["model"], corpus.dataset("train"))
    p = {k: v.astype(np.float64) for k, v in _sd(g, "sd0/").items()}
    r = net.step(p, g["s0/u"], g["s0/i"])
    assert np.abs(r["logits"] - g["s0/logits"]).max() < 1e-5
    assert abs(r["rec_loss"] - g["s0/rec_loss"]) < 1e-6 * max(1.0, abs(g["s0/rec_loss"]))
    for k, want in _sd(g, "s0/grad/").items():
        got = r["grads"].get(k, np.zeros_like(want))
        assert np.abs(got - want).max() < 1e-4 * max(1e-4, np.abs(want).max()), k
    p1 = {k: v.astype(np.float64) for k, v in _sd(g, "s0/sd/").items()}
    val = corpus.dataset("val")
    s = net.scores(p1, np.asarray(val.users_in_split), np.asarray(val.items_in_split))
    s[val.exclude_data[np.asarray(val.users_in_split)].toarray().astype(bool)] = -np.inf
    finite = np.isfinite(g["eval/scores"])
    assert np.array_equal(finite, np.isfinite(s))
    assert np.abs(s[finite] - g["eval/scores"][finite]).max() < 1e-5


def test_deepmf_state_dict_keys_and_no_cpu_fallback():
    from sibrar_b200.deepmf import DeepMatrixFactorization
    spec, g, corpus = _load("dmf_plain")
    model = DeepMatrixFactorization.build_from_conf(dict(spec["model"]), corpus.dataset("train"))
    assert sorted(model.state_dict().keys()) == sorted(_sd(g, "sd0/").keys())  # the reference's checkpoint keys
    assert model.u_layers == [corpus.dataset("train").n_items, 32, 16]
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, dtype=torch.int64), torch.zeros((2, 3), dtype=torch.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_deepmf_gpu_step_and_eval_match_reference(name):
    from oracle import sbnet_oracle as O
    from sibrar_b200.deepmf import DeepMatrixFactorization
    from sibrar_b200.evaluator import FullEvaluator
    spec, g, corpus = _load(name)
    train = corpus.dataset("train")
    model = DeepMatrixFactorization.build_from_conf(dict(spec["model"]), train)
    sd0 = _sd(g, "sd0/")
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd0.items()})
    model.to(DEV).train()
    u, i = torch.from_numpy(g["s0/u"]).to(DEV), torch.from_numpy(g["s0/i"]).to(DEV)
    opt = torch.optim.AdamW(model.parameters(), lr=spec["lr"], weight_decay=spec["wd"])
    sim = model(u, i)
    assert sim.requires_grad
    net = OracleDeepMF(spec["model"], train)
    emu = net.step({k: v.astype(np.float64) for k, v in sd0.items()}, g["s0/u"], g["s0/i"], emu=O.Bf16Emulation())
    got = sim.detach().cpu().numpy()
    assert got.min() >= spec["model"]["mu"] * (1 - 1e-6)  # the lower clamp
    assert np.abs(got - emu["logits"]).max() < 5e-3   # cosines: absolute
    assert np.abs(got - g["s0/logits"]).max() < 3e-2
    pos, neg = sim[:, :1].double(), sim[:, 1:].double()
    rec = torch.nn.functional.softplus(-(pos - neg)).mean()
    assert float(rec.detach()) == pytest.approx(float(g["s0/rec_loss"]), rel=1e-2)
    rec.backward()
    gscale = max(float(np.abs(v).max()) for v in emu["grads"].values())
    for k, p in model.named_parameters():
        e = emu["grads"][k]
        gp = p.grad.cpu().numpy()
        # a score within bf16 rounding of mu can fall on the other side of the clamp: its whole gradient term flips
        assert np.abs(gp - e).max() <= 2e-2 * np.abs(e).max() + 2e-3 * gscale, ("emulated", k)
        want = g[f"s0/grad/{k}"].reshape(-1).astype(np.float64)
        if np.abs(want).max() > 1e-2 * gscale:
            a = gp.reshape(-1).astype(np.float64)
            assert float(a @ want) / max(1e-30, np.linalg.norm(a) * np.linalg.norm(want)) > 0.95, k
    opt.step()
    model.check_errors()
    for k, v in model.state_dict().items():
        assert np.abs(v.cpu().numpy() - g[f"s0/sd/{k}"]).max() <= 2.1 * spec["lr"], k
    # ---- evaluation through FullEvaluator with the reference's updated weights
    model.load_state_dict({k: torch.from_numpy(v).to(DEV) for k, v in _sd(g, "s0/sd/").items()})
    model.refresh_shadows()
    val = corpus.dataset("val")
    res, (vals, idx) = FullEvaluator(dict(top_k=[1, 3, 5], metrics=["ndcg", "precision", "recall", "hitrate", "coverage"],
                                          calculate_std=False)).evaluate(model, val, return_topk=True)
    model.check_errors()
    ref_scores = g["eval/scores"]
    got_idx = idx.cpu().numpy()
    rows = np.arange(got_idx.shape[0])[:, None]
    picked = ref_scores[rows, np.clip(got_idx, 0, ref_scores.shape[1] - 1)]
    assert np.abs(picked - g["eval/topk_val"]).max() < 3e-2
    for k, v in res.items():
        assert abs(v - float(g[f"eval/metric/{k}"])) < 0.05, (k, v, float(g[f"eval/metric/{k}"]))
    # ---- API: representations + combine (clamped cosines) without a graph
    model.load_state_dict({k: torch.from_numpy(v).to(DEV) for k, v in sd0.items()})
    model.refresh_shadows()
    model.train()
    with torch.no_grad():
        again = model.combine_user_item_representations(model.get_user_representations(u),
                                                        model.get_item_representations(i))
    assert np.abs(again.cpu().numpy() - got).max() < 1e-4
    assert model.predict(u, i).shape == i.shape and not model.training
