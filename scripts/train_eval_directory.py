"""End-to-end use of the B200 path on a preprocessed dataset DIRECTORY (the reference's on-disk format), i.e. what
``run_experiment.py`` -> ``experiment_helper.run_train_val`` does for ``alg: sbnet``, reduced to the hot path:

    DiskCorpus (CSV / NPZ / used_config.yaml)  ->  SingleBranchNet.build_from_conf
    epochs:  DeviceBatchFeeder (shuffled interactions + uniform_recbole negatives, on the device)
             -> FusedTrainer.step (one CUDA graph per batch shape)
             -> FullEvaluator(cuda_graph=True).evaluate on the validation split (model rebuilt on the val dataset,
                weights shared through the state dict, like experiment_helper.py:132)

    python scripts/train_eval_directory.py tests/golden/disk_cs_item --epochs 5 --batch 64
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import sibrar_b200  # noqa: E402,F401
from sibrar_b200.disk_dataset import DiskCorpus  # noqa: E402
from sibrar_b200.evaluator import FullEvaluator  # noqa: E402
from sibrar_b200.sbnet import SingleBranchNet  # noqa: E402
from sibrar_b200.trainer import DeviceBatchFeeder, FusedTrainer  # noqa: E402

USER_FEATURES = [dict(name="gender", type="categorical"), dict(name="age", type="discrete"),
                 dict(name="taste", type="vector")]
ITEM_FEATURES = [dict(name="genres", type="tag", tag_split_sep="|"), dict(name="year", type="continuous"),
                 dict(name="studio", type="categorical"), dict(name="plot", type="vector")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--dim", type=int, default=16)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    corpus = DiskCorpus(args.path, USER_FEATURES, ITEM_FEATURES, n_negative_samples=3,
                        negative_sampling_strategy="uniform_recbole")
    train, val = corpus.dataset("train"), corpus.dataset("val")
    ent = lambda feats, hidden: dict(features=[dict(feature_name=f, feature_hidden_layers=[]) for f in feats],  # noqa
                                     single_branch_hidden_layers=hidden, preference_hidden_layers=[],
                                     common_modality_dim=args.dim, activation_fn="relu")
    conf = dict(shared_common_dim=args.dim, user=ent(["interactions"] + [f["name"] for f in USER_FEATURES], []),
                item=ent(["interactions"] + [f["name"] for f in ITEM_FEATURES], [args.dim]))
    torch.manual_seed(0)
    model = SingleBranchNet.build_from_conf(conf, train).to(dev).train()
    val_model = SingleBranchNet.build_from_conf(conf, val).to(dev)
    trainer = FusedTrainer(model, dict(lr=1e-2, wd=1e-6, optimizer="adamw", rec_loss="bpr", loss_aggregator="mean"),
                           n_negative_samples=3, cuda_graph=True)
    feeder = DeviceBatchFeeder(train, batch_size=args.batch, device=dev, seed=0)
    evaluator = FullEvaluator(dict(top_k=[1, 5], metrics=["ndcg", "recall", "coverage"], calculate_std=False),
                              cuda_graph=True)
    print(f"{train.name}: {train.n_users} users x {train.n_items} items, {len(train)} train interactions, "
          f"{len(feeder)} batches / epoch; cold start item = {train.is_cold_start_item}")
    for epoch in range(args.epochs):
        t0 = time.perf_counter()
        for u, i in feeder.epoch():
            trainer.step(u, i)
        losses = trainer.read_losses()
        val_model.load_state_dict(model.state_dict())
        res = evaluator.evaluate(val_model, val)
        print(f"epoch {epoch}: loss {losses['train/loss']:.4f}  ndcg@5 {res['ndcg@5']:.4f}  recall@5 "
              f"{res['recall@5']:.4f}  coverage@5 {res['coverage@5']:.3f}  ({(time.perf_counter() - t0) * 1e3:.1f} ms)")
    model.check_errors()


if __name__ == "__main__":
    main()
