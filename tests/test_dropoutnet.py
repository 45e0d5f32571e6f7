"""DropoutNet (algorithms/sgd_alg.py:1617-1761): the CPU oracle is pinned against fixtures generated from the unmodified
reference (``oracle/make_golden_dropoutnet.py``); the B200 path (``sibrar_b200.dropoutnet``) is compared with both."""
import os

import numpy as np
import pytest
import torch

from oracle.dropoutnet_oracle import OracleDropoutNet
from oracle.make_golden import GOLDEN_DIR
from oracle.make_golden_dropoutnet import CASES

DEV = "cuda:0"


def _load(name):
    from sibrar_b200.synthetic import SynCorpus
    spec = CASES[name]
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))
    return spec, g, SynCorpus(**spec["corpus"])


def _sd(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


@pytest.mark.parametrize("name", list(CASES))
def test_dropoutnet_oracle_matches_reference_fixture(name):
    spec, g, corpus = _load(name)
    net = OracleDropoutNet(spec["model"], corpus.dataset("train"))
    p = {k: v.astype(np.float64) for k, v in _sd(g, "sd0/").items()}
    r = net.step(p, g["s0/u"], g["s0/i"], g["s0/su"], g["s0/si"])
    assert np.abs(r["logits"] - g["s0/logits"]).max() < 1e-5 * max(1.0, np.abs(g["s0/logits"]).max())
    assert abs(r["rec_loss"] - g["s0/rec_loss"]) < 1e-6 * max(1.0, abs(g["s0/rec_loss"]))
    for k, want in _sd(g, "s0/grad/").items():
        got = r["grads"].get(k, np.zeros_like(want))
        assert np.abs(got - want).max() < 1e-5 * max(1e-3, np.abs(want).max()), k
    # the reference's generator: the same seed draws the same strategies, users first (sgd_alg.py:1683-1690)
    rng = np.random.default_rng(spec["model"]["sampling_seed"])
    assert np.array_equal(rng.choice([1, 2], size=spec["batch"], replace=True), g["s0/su"])
    assert np.array_equal(rng.choice([1, 2], size=spec["batch"], replace=True), g["s0/si"])
    # eval scores with the updated weights
    p1 = {k: v.astype(np.float64) for k, v in _sd(g, "s0/sd/").items()}
    val = corpus.dataset("val")
    s = net.scores(p1, np.asarray(val.users_in_split), np.asarray(val.items_in_split))
    s[val.exclude_data[np.asarray(val.users_in_split)].toarray().astype(bool)] = -np.inf
    finite = np.isfinite(g["eval/scores"])
    assert np.array_equal(finite, np.isfinite(s))
    assert np.abs(s[finite] - g["eval/scores"][finite]).max() < 1e-5 * max(1.0, np.abs(s[finite]).max())


def test_dropoutnet_config_and_state_dict_keys():
    from sibrar_b200.dropoutnet import DropoutNet, DropoutNetConfig, DropoutNetSamplingStrategy
    assert DropoutNetSamplingStrategy.list() == [1, 2]
    spec, g, corpus = _load("dn_vector_tag")
    with pytest.raises(KeyError):
        DropoutNetConfig.from_dict(dict(user={}, item={}))
    model = DropoutNet.build_from_conf(dict(spec["model"]), corpus.dataset("train"))
    assert sorted(model.state_dict().keys()) == sorted(_sd(g, "sd0/").keys())  # the reference's checkpoint keys
    model.eval()
    assert np.all(model.sample_training_strategy(5) == 1)
    model.train()
    assert np.array_equal(model.sample_training_strategy(spec["batch"]), g["s0/su"])  # the seeded stream of the reference
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, dtype=torch.int64), torch.zeros((2, 3), dtype=torch.int64))  # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_dropoutnet_gpu_step_and_eval_match_reference(name):
    from oracle import sbnet_oracle as O
    from sibrar_b200.dropoutnet import DropoutNet
    from sibrar_b200.evaluator import FullEvaluator
    spec, g, corpus = _load(name)
    train = corpus.dataset("train")
    model = DropoutNet.build_from_conf(dict(spec["model"]), train)
    sd0 = _sd(g, "sd0/")
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd0.items()})
    model.to(DEV).train()
    u, i = torch.from_numpy(g["s0/u"]).to(DEV), torch.from_numpy(g["s0/i"]).to(DEV)
    # ---- the reference's loop: logits (the model draws its own strategies: same seed, same draws) -> torch BPR ->
    # backward -> torch.optim.AdamW
    opt = torch.optim.AdamW(model.parameters(), lr=spec["lr"], weight_decay=spec["wd"])
    logits = model(u, i)
    assert logits.requires_grad
    # bf16 GEMM operands: against the fp32 fixture loosely, against the oracle with the kernels' rounding points tightly
    net = OracleDropoutNet(spec["model"], train)
    emu = net.step({k: v.astype(np.float64) for k, v in sd0.items()}, g["s0/u"], g["s0/i"], g["s0/su"], g["s0/si"],
                   emu=O.Bf16Emulation())
    got = logits.detach().cpu().numpy()
    scale = max(1e-3, np.abs(g["s0/logits"]).max())
    assert np.abs(got - emu["logits"]).max() < 5e-3 * scale
    assert np.abs(got - g["s0/logits"]).max() < 3e-2 * scale
    pos, neg = logits[:, :1].double(), logits[:, 1:].double()
    rec = torch.nn.functional.softplus(-(pos - neg)).mean()
    assert float(rec) == pytest.approx(float(g["s0/rec_loss"]), rel=1e-2)
    assert float(model.get_and_reset_other_loss()["reg_loss"].sum()) == 0.0
    rec.backward()
    gscale = max(float(np.abs(v).max()) for v in emu["grads"].values())
    for k, p in model.named_parameters():
        e = emu["grads"][k]
        gp = p.grad.cpu().numpy()
        assert np.abs(gp - e).max() <= 1e-2 * np.abs(e).max() + 1e-4 * gscale, ("emulated", k)
        want = g[f"s0/grad/{k}"].reshape(-1).astype(np.float64)
        if np.abs(want).max() > 1e-3 * gscale:  # direction against the fp32 reference
            a = gp.reshape(-1).astype(np.float64)
            assert float(a @ want) / max(1e-30, np.linalg.norm(a) * np.linalg.norm(want)) > 0.98, k
    opt.step()
    model.check_errors()
    for k, v in model.state_dict().items():
        assert np.abs(v.cpu().numpy() - g[f"s0/sd/{k}"]).max() <= 2.1 * spec["lr"], k  # (one Adam step moves <= lr)
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
    assert np.abs(picked - g["eval/topk_val"]).max() < 3e-2 * max(1.0, np.abs(g["eval/topk_val"]).max())
    for k, v in res.items():
        assert abs(v - float(g[f"eval/metric/{k}"])) < 0.05, (k, v, float(g[f"eval/metric/{k}"]))
    # ---- API: representations + combine without a graph, explicit strategies, eval mode keeps every preference
    model.load_state_dict({k: torch.from_numpy(v).to(DEV) for k, v in sd0.items()})
    model.refresh_shadows()
    model.train()
    with torch.no_grad():
        ur = model.get_user_representations(u, strategy=g["s0/su"])
        ir = model.get_item_representations(i, strategy=g["s0/si"])
        again = model.combine_user_item_representations(ur, ir)
    assert np.abs(again.cpu().numpy() - got).max() < 1e-4 * scale + 1e-5
    assert model.predict(u, i).shape == i.shape and not model.training
