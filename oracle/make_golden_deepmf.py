"""TEST TOOLING -- golden fixtures of DeepMatrixFactorization from the UNMODIFIED reference (``/root/reference``, imported under
``oracle/ref_shims.py``): run in the build container, commit ``tests/golden/dmf_*.npz``.

    python -m oracle.make_golden_deepmf
"""
import copy
import os
import sys

import numpy as np

from oracle import ref_shims
from oracle.make_golden import GOLDEN_DIR, PLAIN_FIXTURE_SHAPE, build_reference_datasets

CASES = {
    # conf/single/algorithms/dmf_ml1m_conf.yml shape: one middle layer per side, plain interaction vectors
    "dmf_plain": dict(
        corpus=dict(shape=PLAIN_FIXTURE_SHAPE, split_type="random", seed=61, scale=1.0),
        model=dict(u_mid_layers=[32], i_mid_layers=[24], final_dimension=16, mu=1e-6),
        batch=24, n_neg=4, lr=1e-3, wd=1e-6),
    # normalised interaction vectors and representations, ReLU on the towers' outputs, no middle layer on the item side
    "dmf_normalized_relu": dict(
        corpus=dict(shape=PLAIN_FIXTURE_SHAPE, split_type="random", seed=63, scale=1.0),
        model=dict(u_mid_layers=24, i_mid_layers=[], final_dimension=16, mu=1e-6, normalize_interactions=True,
                   normalize_representations=True, use_output_activation_fn=True),
        batch=24, n_neg=4, lr=1e-3, wd=1e-6),
}


def run_case(name, spec):
    import torch
    ref_shims.install()
    from sibrar_b200.synthetic import SynCorpus, sample_batch
    from algorithms.sgd_alg import DeepMatrixFactorization
    from train.rec_losses import RecommenderSystemLossesEnum
    from eval.eval import FullEvaluator, evaluate_recommender_algorithm
    from data.config_classes import EvalConfig
    from torch.utils.data import DataLoader

    torch.manual_seed(4321)
    torch.set_num_threads(1)
    corpus = SynCorpus(**spec["corpus"])
    dss = build_reference_datasets(corpus)
    for ds in dss.values():  # the item side of the train interactions (data/dataset.py:269-273)
        ds.model_requires_item_interactions = True
        ds.item_sampling_matrix_train = ds.user_sampling_matrix_train.T.tocsr()
    model = DeepMatrixFactorization.build_from_conf(copy.deepcopy(spec["model"]), dss["train"])
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
    rec.backward()
    out["s0/u"], out["s0/i"] = u, i
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
          f"{float(rec.item()):.5f} clamped {int((logits <= spec['model']['mu']).sum())} ndcg@5 {metrics.get('ndcg@5')}")


def main():
    import sibrar_b200  # noqa: F401
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if not only or name in only:
            run_case(name, spec)


if __name__ == "__main__":
    main()
