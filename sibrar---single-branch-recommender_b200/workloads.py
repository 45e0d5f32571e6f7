"""The BASELINE.json workloads as (synthetic corpus, model block, learn block) triples -- shared by ``bench.py``, the
parity tests at bench shape and the profiling scripts.  Model blocks restate the reference's YAML files key by key:

  * ``ml1m``          conf/single/algorithms/sbnet_ml1m_conf.yml:21-52            (BASELINE configs[0] / [1])
  * ``onion18_huge``  conf/single/algorithms/sbnet_onion18_huge_conf.yml:31-66 over the feature list of
                      sbnet_onion18_conf.yml:22-60                                 (BASELINE configs[2])
  * ``amazon_nouser`` conf/single/algorithms/sbnet_amazonvid2024_huge_no-user_conf.yml:30-63  (BASELINE configs[3])

Synthetic feature names follow ``synthetic.SHAPES`` (the real datasets are not reachable: no network).
"""
from __future__ import annotations

import copy

N_NEG = 10


def _features(names_hidden):
    return [dict(feature_name=f, feature_hidden_layers=list(h)) for f, h in names_hidden]


def ml1m_conf(D: int = 64, batch_norm: bool = True):
    ent = lambda feats, hidden, drop: dict(  # noqa: E731
        features=_features([(f, []) for f in feats]), single_branch_hidden_layers=hidden, preference_hidden_layers=[],
        common_modality_dim=D, activation_fn="relu", single_branch_input_dropout=drop,
        apply_batch_normalization=batch_norm)
    return dict(shared_common_dim=D, user=ent(["interactions", "gender", "occupation"], [], None),
                item=ent(["interactions", "genres", "plot_mpnet"], [D], 0.2))


def _huge_item(features, train_modalities, eval_modalities=None):
    d = dict(features=_features(features), single_branch_hidden_layers=[512, 512, 512, 256, 256],
             preference_hidden_layers=[], common_modality_dim=512, activation_fn="relu",
             single_branch_input_dropout=2e-2, normalize_single_branch_input=True, aggregation_fn="mean",
             embedding_regularization_type="pairwise_single", central_modality="interactions",
             train_modalities=list(train_modalities), apply_output_activation=True, apply_batch_normalization=True,
             apply_batch_norm_every=2)
    if eval_modalities is not None:
        d["eval_modalities"] = list(eval_modalities)
    return d


def onion18_huge_conf():
    user = dict(features=_features([("interactions", []), ("age", []), ("gender", []), ("country", []),
                                    ("mpnet", [128])]),
                single_branch_hidden_layers=[], preference_hidden_layers=[], common_modality_dim=128,
                activation_fn="relu", single_branch_input_dropout=None, normalize_single_branch_input=True,
                train_modalities=["interactions"], aggregation_fn="mean",
                embedding_regularization_type="no_regularization", apply_output_activation=True,
                apply_batch_normalization=True)
    item = _huge_item([("interactions", []), ("musicnn", []), ("lyrics_mpnet", []), ("jukebox", []), ("genres", [])],
                      ["interactions", "genres", "jukebox"])
    return dict(shared_common_dim=128, user=user, item=item)


def amazon_nouser_conf(eval_modalities=None):
    item = _huge_item([("interactions", []), ("title_mpnet", []), ("description_mpnet", []), ("image_resnet", [])],
                      ["interactions", "title_mpnet"], eval_modalities)
    return dict(shared_common_dim=128, user=dict(feature_name="user_embedding", embedding_dim=-1, activation_fn="relu"),
                item=item)


WORKLOADS = {
    "ml1m": dict(
        corpus=dict(shape="ml1m", split_type="cold_start_item", seed=42), conf=ml1m_conf,
        learn=dict(lr=1e-3, wd=1e-6, optimizer="adamw", rec_loss="bpr", loss_aggregator="mean"), batch=16384,
        text="SBNet train step, synthetic ML-1M shape (6040 users x 3706 items, 1,000,209 interactions; user: "
             "interactions/gender/occupation, item: interactions/genres(18 tags)/plot_mpnet(768)), cold_start_item "
             "split, sbnet_ml1m_conf model (C=D=64, item MLP [64], item input dropout 0.2, trailing BatchNorm), BPR, "
             "AdamW, n_neg=10"),
    "onion18_huge": dict(
        corpus=dict(shape="onion18", split_type="random", seed=42), conf=onion18_huge_conf,
        learn=dict(lr=5e-5, wd=1e-3, optimizer="adamw", rec_loss="bpr", loss_aggregator="mean"), batch=16384,
        text="SBNet train step, synthetic Onion18 shape (50,000 users x 100,000 items, 5,000,000 interactions; item: "
             "interactions(sparse)/genres(853 tags)/jukebox(4800), user: interactions(sparse)), random split, "
             "sbnet_onion18_huge_conf model (item C=512, MLP [512,512,512,256,256], BatchNorm every 2, L2-normalised "
             "input, dropout 0.02, pairwise InfoNCE k=2; user C=128; D=128), BPR, AdamW, n_neg=10"),
    "amazon_nouser": dict(
        corpus=dict(shape="amazonvid2024", split_type="random", seed=42), conf=amazon_nouser_conf,
        learn=dict(lr=5e-5, wd=1e-3, optimizer="adamw", rec_loss="bpr", loss_aggregator="mean"), batch=16384,
        text="SBNet train step, synthetic AmazonVideo2024 shape at paper size (11,454 users x 4,177 items, 87,098 "
             "interactions; item: interactions/title_mpnet(768), plain user embedding), random split, "
             "sbnet_amazonvid2024_huge_no-user_conf model (item C=512, MLP [512,512,512,256,256], BatchNorm every 2, "
             "pairwise InfoNCE k=2, D=128), BPR, AdamW, n_neg=10"),
}


def build(name: str, scale: float = 1.0, **conf_kw):
    """-> (SynCorpus, model conf dict, learn dict, default batch per GPU, description)"""
    from .synthetic import SynCorpus
    w = WORKLOADS[name]
    corpus = SynCorpus(**dict(w["corpus"], scale=scale))
    return corpus, copy.deepcopy(w["conf"](**conf_kw)), dict(w["learn"]), w["batch"], w["text"]
