"""TEST TOOLING -- golden fixtures of the matrix-factorisation siblings from the UNMODIFIED reference
(``/root/reference``, imported under ``oracle/ref_shims.py``): run in the build container, commit ``tests/golden/mf_*.npz``.

    python -m oracle.make_golden_mf
"""
import copy
import os
import sys

import numpy as np

from oracle import ref_shims
from oracle.make_golden import GOLDEN_DIR, PLAIN_FIXTURE_SHAPE, build_reference_datasets

MF_BASE = dict(embedding_dim=16, use_user_bias=False, use_item_bias=False, use_global_bias=False)
CASES = {
    # conf/single/algorithms/ifmf_ml1m_conf.yml: tag feature, item bias, aggregate_for_rec
    "mf_ifmf_genres": dict(kind="ifmf", corpus=dict(shape=PLAIN_FIXTURE_SHAPE, split_type="random", seed=41, scale=1.0),
                           model=dict(MF_BASE, use_item_bias=True, aggregate_for_rec=True, feature_name="genres",
                                      lambda_content=1e-4, temperature=0.1, embedding_loss_aggregator="mean",
                                      intermediate_layers=None), batch=24, n_neg=4, lr=1e-3, wd=1e-6),
    # vector feature with intermediate layers, profile-only scores, summed InfoNCE, global bias
    # (use_user_bias = True crashes in the reference outside UFMF: the [B, 1] bias of SGDMatrixFactorization is
    # broadcast against [B, n] logits as [B, 1, 1], sgd_alg.py:166, 190)
    "mf_ifmf_vector": dict(kind="ifmf", corpus=dict(shape=PLAIN_FIXTURE_SHAPE, split_type="random", seed=43, scale=1.0),
                           model=dict(MF_BASE, use_global_bias=True, aggregate_for_rec=False,
                                      feature_name="plot_mpnet", lambda_content=1e-4, temperature=0.5,
                                      embedding_loss_aggregator="sum", intermediate_layers=[20]),
                           batch=24, n_neg=4, lr=1e-3, wd=1e-6),
    # conf/single/algorithms/ufmf_ml1m_conf.yml: categorical user feature, user bias, aggregate_for_rec
    "mf_ufmf_country": dict(kind="ufmf", corpus=dict(shape=PLAIN_FIXTURE_SHAPE, split_type="random", seed=45, scale=1.0),
                            model=dict(MF_BASE, use_user_bias=True, aggregate_for_rec=True, feature_name="country",
                                       lambda_content=1e-4, temperature=0.1, embedding_loss_aggregator="mean",
                                       intermediate_layers=None), batch=24, n_neg=4, lr=1e-3, wd=1e-6),
    # plain SGDMatrixFactorization with item and global bias
    "mf_plain": dict(kind="mf", corpus=dict(shape=PLAIN_FIXTURE_SHAPE, split_type="random", seed=47, scale=1.0),
                     model=dict(MF_BASE, use_item_bias=True, use_global_bias=True),
                     batch=24, n_neg=4, lr=1e-3, wd=1e-6),
}


def run_case(name, spec):
    import torch
    ref_shims.install()
    from sibrar_b200.synthetic import SynCorpus, sample_batch
    from algorithms.sgd_alg import (ItemFeatureMatrixFactorization, SGDMatrixFactorization,
                                    UserFeatureMatrixFactorization)
    from train.rec_losses import RecommenderSystemLossesEnum
    from eval.eval import FullEvaluator, evaluate_recommender_algorithm
    from data.config_classes import EvalConfig
    from torch.utils.data import DataLoader

    torch.manual_seed(4321)
    torch.set_num_threads(1)
    corpus = SynCorpus(**spec["corpus"])
    dss = build_reference_datasets(corpus)
    cls = {"ifmf": ItemFeatureMatrixFactorization, "ufmf": UserFeatureMatrixFactorization,
           "mf": SGDMatrixFactorization}[spec["kind"]]
    model = cls.build_from_conf(copy.deepcopy(spec["model"]), dss["train"])
    with torch.no_grad():  # non-trivial biases (the reference initialises the global bias with 0)
        for k, p in model.named_parameters():
            if "bias" in k and "embedding_net" not in k:
                p.copy_(0.1 * torch.randn_like(p))
    out = {f"sd0/{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}
    loss_fn = RecommenderSystemLossesEnum["bpr"].value(n_items=dss["train"].n_items, aggregator="mean",
                                                       train_neg_strategy="uniform_recbole", neg_train=spec["n_neg"])
    opt = torch.optim.AdamW(model.parameters(), lr=spec["lr"], weight_decay=spec["wd"])
    rng = np.random.default_rng(77)
    model.train()
    u, i = sample_batch(corpus.dataset("train"), spec["batch"], rng, spec["n_neg"])
    labels = torch.zeros(i.shape, dtype=torch.float64)
    labels[:, 0] = 1.
    logits = model(torch.from_numpy(u), torch.from_numpy(i))
    rec = loss_fn.compute_loss(logits, labels)
    reg = model.get_and_reset_other_loss()["reg_loss"]
    reg = reg if torch.is_tensor(reg) else torch.tensor(float(reg))
    total = rec + reg
    total.backward()
    out["s0/u"], out["s0/i"] = u, i
    out["s0/logits"] = logits.detach().numpy().copy()
    out["s0/rec_loss"], out["s0/reg_loss"], out["s0/loss"] = (np.float64(rec.item()), np.float64(reg.item()),
                                                              np.float64(total.item()))
    for k, p in model.named_parameters():
        out[f"s0/grad/{k}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
    opt.step()
    for k, v in model.state_dict().items():
        out[f"s0/sd/{k}"] = v.detach().numpy().copy()
    # ---- evaluation with the updated weights
    val = dss["val"]
    ev_conf = EvalConfig(top_k=[1, 3, 5], metrics=["ndcg", "precision", "recall", "hitrate", "coverage"],
                         calculate_std=False)
    evaluator = FullEvaluator(ev_conf, dataset=val)
    model.eval()
    with torch.no_grad():
        i_repr = model.get_item_representations(torch.tensor(val.items_in_split))
        u_idx = torch.tensor(val.users_in_split)
        u_repr = model.get_user_representations(u_idx)
        scores = model.combine_user_item_representations(u_repr, i_repr)
        mask = torch.tensor(val.exclude_data[u_idx].toarray(), dtype=torch.bool)
        scores[mask] = -torch.inf
    out["eval/scores"] = scores.numpy()
    top = torch.topk(scores, 5, dim=-1)
    out["eval/topk_idx"], out["eval/topk_val"] = top.indices.numpy(), top.values.numpy()
    metrics = evaluate_recommender_algorithm(model, DataLoader(val, batch_size=7, shuffle=False), evaluator, device="cpu")
    for k, v in metrics.items():
        out[f"eval/metric/{k}"] = np.float64(v)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"[golden] {name}: {len(out)} arrays -> {path} ({os.path.getsize(path) / 1024:.0f} KiB); loss "
          f"{float(total.item()):.5f} reg {float(reg.item()):.5f} ndcg@5 {metrics.get('ndcg@5')}")


def main():
    import sibrar_b200  # noqa: F401
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if not only or name in only:
            run_case(name, spec)


if __name__ == "__main__":
    main()
