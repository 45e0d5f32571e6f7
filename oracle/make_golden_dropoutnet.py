"""TEST TOOLING -- golden fixtures of DropoutNet from the UNMODIFIED reference (``/root/reference``, imported under
``oracle/ref_shims.py``): run in the build container, commit ``tests/golden/dn_*.npz``.

    python -m oracle.make_golden_dropoutnet
"""
import copy
import os
import sys

import numpy as np

from oracle import ref_shims
from oracle.make_golden import GOLDEN_DIR, PLAIN_FIXTURE_SHAPE, build_reference_datasets

CASES = {
    # conf/single/algorithms/dropoutnet_ml1m_conf.yml shape: content = one feature per side, two preference layers
    "dn_vector_tag": dict(
        corpus=dict(shape=PLAIN_FIXTURE_SHAPE, split_type="random", seed=51, scale=1.0),
        model=dict(shared_common_dim=12, sampling_seed=7,
                   user=dict(features=[dict(feature_name="mpnet", embedding_dim=16, pre_embedding_layers=[24])],
                             preference_layers=[24, 16], common_hidden_layers=[20], activation_fn="relu"),
                   item=dict(features=[dict(feature_name="genres", embedding_dim=8),
                                       dict(feature_name="plot_mpnet", embedding_dim=16)],
                             preference_layers=[16], common_hidden_layers=[], activation_fn="relu")),
        batch=24, n_neg=4, lr=1e-3, wd=1e-6),
    # categorical user feature with post layers, tanh in the common net, no hidden layer on the user side
    "dn_categorical_tanh": dict(
        corpus=dict(shape=PLAIN_FIXTURE_SHAPE, split_type="random", seed=53, scale=1.0),
        model=dict(shared_common_dim=8, sampling_seed=11,
                   user=dict(features=[dict(feature_name="country", embedding_dim=8, post_embedding_layers=[8])],
                             preference_layers=[16], common_hidden_layers=[], activation_fn="tanh"),
                   item=dict(features=[dict(feature_name="plot_mpnet", embedding_dim=16, pre_embedding_layers=[32])],
                             preference_layers=[32, 16], common_hidden_layers=[16], activation_fn="tanh")),
        batch=24, n_neg=4, lr=1e-3, wd=1e-6),
}


def run_case(name, spec):
    import torch
    ref_shims.install()
    from sibrar_b200.synthetic import SynCorpus, sample_batch
    from algorithms.sgd_alg import DropoutNet
    from train.rec_losses import RecommenderSystemLossesEnum
    from eval.eval import FullEvaluator, evaluate_recommender_algorithm
    from data.config_classes import EvalConfig
    from torch.utils.data import DataLoader

    torch.manual_seed(4321)
    torch.set_num_threads(1)
    corpus = SynCorpus(**spec["corpus"])
    dss = build_reference_datasets(corpus)
    for ds in dss.values():  # DropoutNet needs the item side of the train interactions (data/dataset.py:269-273)
        ds.model_requires_item_interactions = True
        ds.item_sampling_matrix_train = ds.user_sampling_matrix_train.T.tocsr()
    model = DropoutNet.build_from_conf(copy.deepcopy(spec["model"]), dss["train"])
    out = {f"sd0/{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}
    loss_fn = RecommenderSystemLossesEnum["bpr"].value(n_items=dss["train"].n_items, aggregator="mean",
                                                       train_neg_strategy="uniform_recbole", neg_train=spec["n_neg"])
    opt = torch.optim.AdamW(model.parameters(), lr=spec["lr"], weight_decay=spec["wd"])
    rng = np.random.default_rng(77)
    model.train()
    u, i = sample_batch(corpus.dataset("train"), spec["batch"], rng, spec["n_neg"])
    labels = torch.zeros(i.shape, dtype=torch.float64)
    labels[:, 0] = 1.
    drawn = []
    orig = model.sample_training_strategy

    def recording(n_samples):
        s = orig(n_samples)
        drawn.append(np.asarray(s).copy())
        return s
    model.sample_training_strategy = recording
    logits = model(torch.from_numpy(u), torch.from_numpy(i))
    model.sample_training_strategy = orig
    assert len(drawn) == 2 and all(len(d) == spec["batch"] for d in drawn)
    rec = loss_fn.compute_loss(logits, labels)
    rec.backward()
    out["s0/u"], out["s0/i"], out["s0/su"], out["s0/si"] = u, i, drawn[0].astype(np.int64), drawn[1].astype(np.int64)
    out["s0/logits"] = logits.detach().numpy().copy()
    out["s0/rec_loss"] = np.float64(rec.item())
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
          f"{float(rec.item()):.5f} strategies {drawn[0][:8]} {drawn[1][:8]} ndcg@5 {metrics.get('ndcg@5')}")


def main():
    import sibrar_b200  # noqa: F401
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if not only or name in only:
            run_case(name, spec)


if __name__ == "__main__":
    main()
