"""Matrix-factorisation siblings (SGDMatrixFactorization / ItemFeature... / UserFeatureMatrixFactorization,
algorithms/sgd_alg.py:126-200, 1399-1614): the CPU oracle is pinned against fixtures generated from the unmodified
reference (``oracle/make_golden_mf.py``); the B200 path (``sibrar_b200.sibling``) is compared with both."""
import os

import numpy as np
import pytest
import torch

from oracle.make_golden import GOLDEN_DIR
from oracle.make_golden_mf import CASES
from oracle.mf_oracle import OracleFeatureMF

DEV = "cuda:0"


def _load(name):
    from sibrar_b200.synthetic import SynCorpus
    spec = CASES[name]
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))
    corpus = SynCorpus(**spec["corpus"])
    return spec, g, corpus


def _sd(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


@pytest.mark.parametrize("name", list(CASES))
def test_mf_oracle_matches_reference_fixture(name):
    spec, g, corpus = _load(name)
    net = OracleFeatureMF(spec["kind"], spec["model"], corpus.dataset("train"))
    p = {k: v.astype(np.float64) for k, v in _sd(g, "sd0/").items()}
    r = net.step(p, g["s0/u"], g["s0/i"])
    assert np.abs(r["logits"] - g["s0/logits"]).max() < 1e-5
    assert abs(r["rec_loss"] - g["s0/rec_loss"]) < 1e-6 * max(1.0, abs(g["s0/rec_loss"]))
    assert abs(r["reg_loss"] - g["s0/reg_loss"]) < 2e-6 * max(1.0, abs(g["s0/reg_loss"]))
    for k, want in _sd(g, "s0/grad/").items():
        got = r["grads"].get(k, np.zeros_like(want))
        assert np.abs(got - want).max() < 1e-5 * max(1.0, np.abs(want).max()), k
    # eval scores with the updated weights
    p1 = {k: v.astype(np.float64) for k, v in _sd(g, "s0/sd/").items()}
    val = corpus.dataset("val")
    s = net.scores(p1, np.asarray(val.users_in_split), np.asarray(val.items_in_split))
    s[val.exclude_data[np.asarray(val.users_in_split)].toarray().astype(bool)] = -np.inf
    finite = np.isfinite(g["eval/scores"])
    assert np.array_equal(finite, np.isfinite(s))
    assert np.abs(s[finite] - g["eval/scores"][finite]).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_mf_gpu_step_and_eval_match_reference(name):
    from sibrar_b200 import sibling
    from sibrar_b200.evaluator import FullEvaluator
    spec, g, corpus = _load(name)
    cls = {"ifmf": sibling.ItemFeatureMatrixFactorization, "ufmf": sibling.UserFeatureMatrixFactorization,
           "mf": sibling.SGDMatrixFactorization}[spec["kind"]]
    train = corpus.dataset("train")
    model = cls.build_from_conf(dict(spec["model"]), train)
    sd0 = _sd(g, "sd0/")
    assert sorted(model.state_dict().keys()) == sorted(sd0.keys())  # the reference's checkpoint keys
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd0.items()})
    model.to(DEV).train()
    u, i = torch.from_numpy(g["s0/u"]).to(DEV), torch.from_numpy(g["s0/i"]).to(DEV)
    # ---- the reference's loop: logits -> torch BPR (fp64) + reg -> backward -> torch.optim.AdamW
    opt = torch.optim.AdamW(model.parameters(), lr=spec["lr"], weight_decay=spec["wd"])
    logits = model(u, i)
    assert logits.requires_grad
    # the content tower runs bf16 GEMM operands when it has Linear layers; embeddings / biases / InfoNCE are fp32
    ltol = 2e-2 if spec["model"].get("intermediate_layers") else 1e-5
    assert np.abs(logits.detach().cpu().numpy() - g["s0/logits"]).max() < ltol
    pos, neg = logits[:, :1].double(), logits[:, 1:].double()
    rec = torch.nn.functional.softplus(-(pos - neg)).mean()
    reg = model.get_and_reset_other_loss()["reg_loss"].double().sum()
    assert float(rec) == pytest.approx(float(g["s0/rec_loss"]), rel=1e-5)
    assert float(reg) == pytest.approx(float(g["s0/reg_loss"]), rel=2e-2 if spec["model"].get("intermediate_layers")
                                       else 1e-4, abs=1e-6)
    (rec + reg).backward()
    # gradients: fp32 paths against the reference fixture; a content tower with Linear layers (bf16 GEMM operands)
    # tightly against the oracle run with the kernels' rounding points, loosely against the fp32 fixture
    net = OracleFeatureMF(spec["kind"], spec["model"], train)
    from oracle import sbnet_oracle as O
    tower = bool(spec["model"].get("intermediate_layers"))
    emu = net.step({k: v.astype(np.float64) for k, v in sd0.items()}, g["s0/u"], g["s0/i"], emu=O.Bf16Emulation()) \
        if tower else None
    for k, p in model.named_parameters():
        want = g[f"s0/grad/{k}"]
        got = p.grad.cpu().numpy()
        if tower:
            e = emu["grads"][k]
            assert np.abs(got - e).max() <= 1e-2 * max(1e-6, np.abs(e).max()) + 1e-7, ("emulated", k)
            assert np.abs(got - want).max() <= 2e-1 * max(1e-6, np.abs(want).max()) + 1e-7, k  # (summed InfoNCE: near-cancelling terms)
        else:
            assert np.abs(got - want).max() <= 1e-4 * max(1e-6, np.abs(want).max()) + 1e-7, k
    opt.step()
    for k, v in model.state_dict().items():
        want = g[f"s0/sd/{k}"]
        if k == "global_bias" or np.abs(g[f"s0/grad/{k}"]).max() < 1e-6:
            # BPR only sees score differences: d loss / d global_bias is exactly 0 up to rounding noise (~1e-9), and the
            # first Adam step turns noise / (|noise| + eps) into a fraction of lr -- not comparable
            assert np.abs(v.cpu().numpy() - want).max() <= 2.1 * spec["lr"]
            continue
        # (the first Adam step moves an entry by lr * g / (|g| + eps): where the gradient is rounding noise the sign is too)
        gref = g[f"s0/grad/{k}"]
        solid = np.abs(gref) > (2e-2 if not tower else 0.5) * np.abs(gref).max()  # (bf16 tower: only the largest entries)
        diff = np.abs(v.cpu().numpy() - want)
        assert diff[solid].max(initial=0.0) < 1e-4 * max(1.0, np.abs(want).max()) + 2e-3 * spec["lr"], k
        assert diff.max() <= 2.1 * spec["lr"], k
    # ---- evaluation through FullEvaluator (fused score + mask + top-k on the bias-augmented factors)
    model.load_state_dict({k: torch.from_numpy(v).to(DEV) for k, v in _sd(g, "s0/sd/").items()})
    model.refresh_shadows()
    val = corpus.dataset("val")
    res, (vals, idx) = FullEvaluator(dict(top_k=[1, 3, 5], metrics=["ndcg", "precision", "recall", "hitrate", "coverage"],
                                          calculate_std=False)).evaluate(model, val, return_topk=True)
    model.check_errors()
    # scores of the returned positions equal the reference's scores there (bf16 factors: 1e-2); metrics agree unless a
    # near-tie swaps a rank
    ref_scores = g["eval/scores"]
    got_idx = idx.cpu().numpy()
    rows = np.arange(got_idx.shape[0])[:, None]
    picked = ref_scores[rows, np.clip(got_idx, 0, ref_scores.shape[1] - 1)]
    assert np.abs(picked - g["eval/topk_val"]).max() < 3e-2 * max(1.0, np.abs(g["eval/topk_val"]).max())
    for k, v in res.items():
        assert abs(v - float(g[f"eval/metric/{k}"])) < 0.05, (k, v, float(g[f"eval/metric/{k}"]))
    # API: the tuple-returning representation calls combine to the same logits without a graph
    model.train()
    with torch.no_grad():
        again = model.combine_user_item_representations(model.get_user_representations(u),
                                                        model.get_item_representations(i))
    model.load_state_dict({k: torch.from_numpy(v).to(DEV) for k, v in sd0.items()})
    model.refresh_shadows()
    with torch.no_grad():
        again = model.combine_user_item_representations(model.get_user_representations(u),
                                                        model.get_item_representations(i))
    assert np.abs(again.cpu().numpy() - g["s0/logits"]).max() < max(ltol, 1e-4)
    assert net is not None
