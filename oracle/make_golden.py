"""TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/*.npz`` by running the UNMODIFIED reference on CPU.

Run in the build container (needs ``/root/reference``):  ``python -m oracle.make_golden``.

For every case in ``CASES`` the real reference ``SingleBranchNet`` (``algorithms/sgd_alg.py:2009-2144``) is built on
a tiny seeded synthetic corpus, then:
  * train: N steps of  forward -> rec loss (``train/rec_losses.py``) -> ``get_and_reset_other_loss`` ->
    ``backward`` -> ``torch.optim.{AdamW,Adam}`` (the body of ``train/trainer.py:204-223``), with the sampled
    modalities recorded (the reference's own sampling order depends on PYTHONHASHSEED) and dropout
    masks recorded through a hooked ``nn.Dropout``;
  * eval: ``evaluate_recommender_algorithm`` (``eval/eval.py:171-227``) with the restated ``rmet``.
Inputs, initial ``state_dict``, logits, losses, gradients, updated parameters, BN running stats, item/user
representations, top-k and per-user metrics go into one ``.npz`` per case.
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def entity(features, hidden, C, **kw):
    d = dict(features=[dict(feature_name=f, feature_hidden_layers=h) for f, h in features],
             single_branch_hidden_layers=hidden, preference_hidden_layers=[], common_modality_dim=C,
             activation_fn="relu")
    d.update(kw)
    return d


def _fixture_shapes():
    from sibrar_b200.synthetic import CATEGORICAL, CONTINUOUS, TAG, VECTOR, Shape
    onion = Shape(150, 320, 6000,
                  {"gender": (CATEGORICAL, 3), "country": (CATEGORICAL, 12), "age": (CONTINUOUS, None),
                   "mpnet": (VECTOR, 96)},
                  {"genres": (TAG, 24), "jukebox": (VECTOR, 1200), "musicnn": (VECTOR, 50),
                   "lyrics_mpnet": (VECTOR, 96)})
    amazon = Shape(137, 260, 3000, {},
                   {"title_mpnet": (VECTOR, 768), "description_mpnet": (VECTOR, 64), "image_resnet": (VECTOR, 64)})
    # users with a vector, a categorical and a TAG feature: the three forms of a plain (non single-branch) entity
    plain = Shape(90, 140, 2200,
                  {"country": (CATEGORICAL, 9), "mpnet": (VECTOR, 40), "languages": (TAG, 7)},
                  {"genres": (TAG, 11), "plot_mpnet": (VECTOR, 24)})
    return onion, amazon, plain


ONION_FIXTURE_SHAPE, AMAZON_FIXTURE_SHAPE, PLAIN_FIXTURE_SHAPE = _fixture_shapes()


def _plain_item():
    return entity([("interactions", []), ("genres", []), ("plot_mpnet", [])], [16], 16, single_branch_input_dropout=0.1)

# name -> dict(corpus kwargs, model conf, loss, optimizer, ...)
CASES = {
    # conf/single/algorithms/sbnet_ml1m_conf.yml model block (C = D = small), trailing BN, item dropout
    "ml1m_small": dict(
        corpus=dict(shape="ml1m", split_type="random", seed=3, scale=0.03, vector_dim_cap=24),
        model=dict(shared_common_dim=16,
                   user=entity([("interactions", []), ("gender", []), ("occupation", [])], [], 16,
                               single_branch_input_dropout=None),
                   item=entity([("interactions", []), ("genres", []), ("plot_mpnet", [])], [16], 16,
                               single_branch_input_dropout=0.2)),
        rec_loss="bpr", optimizer="adamw", lr=1e-3, wd=1e-6, batch=24, n_neg=4, steps=3),
    # sbnet_onion18_huge_conf.yml style: normalize, BN every 2, out activation, pairwise InfoNCE, k = 2
    "pairwise_bn2": dict(
        corpus=dict(shape="ml1m", split_type="cold_start_item", seed=5, scale=0.03, vector_dim_cap=24),
        model=dict(shared_common_dim=8,
                   user=entity([("interactions", []), ("age", [])], [], 16, train_modalities=["interactions"],
                               normalize_single_branch_input=True, apply_output_activation=True),
                   item=entity([("interactions", []), ("genres", []), ("plot_mpnet", [12]), ("item_embedding", [])],
                               [24, 24, 16, 16], 24, single_branch_input_dropout=0.0,
                               normalize_single_branch_input=True, embedding_regularization_type="pairwise_single",
                               regularization_temperature=0.5, regularization_weight=0.3,
                               apply_output_activation=True, apply_batch_norm_every=2)),
        rec_loss="bpr", optimizer="adamw", lr=5e-3, wd=1e-3, batch=16, n_neg=3, steps=3),
    # central modality on both entities (user-side in-batch InfoNCE), max aggregation, no BN, BCE + Adam
    "central_max_bce": dict(
        corpus=dict(shape="ml1m", split_type="random", seed=7, scale=0.03, vector_dim_cap=16),
        model=dict(shared_common_dim=12,
                   user=entity([("interactions", []), ("gender", []), ("occupation", []), ("age", [])], [20], 8,
                               embedding_regularization_type="central_modality", central_modality="interactions",
                               regularization_temperature=1.0, regularization_weight=0.5, aggregation_fn="max",
                               apply_batch_normalization=False),
                   item=entity([("interactions", []), ("genres", []), ("plot_mpnet", [])], [], 8,
                               embedding_regularization_type="central_modality", central_modality="plot_mpnet",
                               regularization_temperature=0.7, regularization_weight=1.0, aggregation_fn="max",
                               apply_batch_normalization=True, apply_batch_norm_every=-1)),
        rec_loss="bce", optimizer="adam", lr=2e-3, wd=1e-4, batch=20, n_neg=2, steps=2),
    # plain user embedding (sbnet_amazonvid2024_huge_no-user_conf.yml), sampled softmax
    "plain_user_ssm": dict(
        corpus=dict(shape="amazonvid2024", split_type="random", seed=9, scale=0.012, vector_dim_cap=20),
        model=dict(shared_common_dim=10,
                   user=dict(feature_name="user_embedding", embedding_dim=-1),
                   item=entity([("interactions", []), ("title_mpnet", []), ("image_resnet", [])], [12], 12,
                               eval_modalities=["title_mpnet", "image_resnet"],
                               embedding_regularization_type="pairwise_single", regularization_weight=0.1)),
        rec_loss="sampled_softmax", optimizer="adamw", lr=1e-3, wd=1e-2, batch=12, n_neg=5, steps=2),
    # ---- plain (non single-branch) USER entities on other than ID features (FeatureEmbedding as the entity module,
    # sgd_alg.py:1279-1396, 2043-2046): vector feature with pre- and post-embedding layers, categorical feature + post
    # layers, tag feature (EmbeddingBag) + post layers
    "plain_user_vector": dict(
        corpus=dict(shape=PLAIN_FIXTURE_SHAPE, split_type="random", seed=31, scale=1.0),
        model=dict(shared_common_dim=12,
                   user=dict(feature_name="mpnet", embedding_dim=20, pre_embedding_layers=[24],
                             post_embedding_layers=[16, 12], activation_fn="relu"),
                   item=_plain_item()),
        rec_loss="bpr", optimizer="adamw", lr=2e-3, wd=1e-4, batch=20, n_neg=3, steps=2),
    "plain_user_categorical_post": dict(
        corpus=dict(shape=PLAIN_FIXTURE_SHAPE, split_type="random", seed=33, scale=1.0),
        model=dict(shared_common_dim=12,
                   user=dict(feature_name="country", embedding_dim=10, post_embedding_layers=[12],
                             activation_fn="tanh"),
                   item=_plain_item()),
        rec_loss="bpr", optimizer="adam", lr=2e-3, wd=0.0, batch=20, n_neg=3, steps=2),
    "plain_user_tag_post": dict(
        corpus=dict(shape=PLAIN_FIXTURE_SHAPE, split_type="random", seed=35, scale=1.0),
        model=dict(shared_common_dim=12,
                   user=dict(feature_name="languages", embedding_dim=8, post_embedding_layers=[12]),
                   item=_plain_item()),
        rec_loss="bpr", optimizer="adamw", lr=2e-3, wd=1e-4, batch=20, n_neg=3, steps=2),
    # tanh everywhere (activation + activation-gradient epilogues other than ReLU), input dropout AND L2 normalisation on
    # both entities, a hidden layer inside a modality projection, an Embedding modality next to tags, Adam without decay
    "tanh_dropout_norm": dict(
        corpus=dict(shape="ml1m", split_type="cold_start_item", seed=13, scale=0.03, vector_dim_cap=16),
        model=dict(shared_common_dim=16,
                   user=entity([("interactions", [12]), ("gender", []), ("occupation", [])], [16], 16,
                               activation_fn="tanh", single_branch_input_dropout=0.3,
                               normalize_single_branch_input=True),
                   item=entity([("interactions", []), ("genres", []), ("plot_mpnet", []), ("item_embedding", [])],
                               [16, 16], 16, activation_fn="tanh", single_branch_input_dropout=0.1,
                               normalize_single_branch_input=True)),
        rec_loss="bce", optimizer="adam", lr=3e-3, wd=0.0, batch=24, n_neg=3, steps=3),
    # ---- REAL LAYER WIDTHS (catalogue scaled down, widths not): conf/single/algorithms/sbnet_onion18_huge_conf.yml:31-66
    # item C = 512, MLP [512, 512, 512, 256, 256] -> D = 128, BatchNorm after every 2nd Linear, output activation, input
    # dropout 0.02, L2 normalisation, pairwise InfoNCE (k = 2); user: train_modalities = [interactions], C = 128,
    # normalised, output activation + trailing BatchNorm.  Weights come from ``seeded_state`` (regenerated by the tests),
    # gradients are stored as scaled fp16 (see ``pack_f16``) to keep the fixture small.
    "onion_huge_widths": dict(
        corpus=dict(shape=ONION_FIXTURE_SHAPE, split_type="random", seed=21, scale=1.0),
        model=dict(shared_common_dim=128,
                   user=entity([("interactions", []), ("age", []), ("gender", []), ("country", []), ("mpnet", [128])],
                               [], 128, single_branch_input_dropout=None, normalize_single_branch_input=True,
                               train_modalities=["interactions"], aggregation_fn="mean",
                               embedding_regularization_type="no_regularization", apply_output_activation=True,
                               apply_batch_normalization=True),
                   item=entity([("interactions", []), ("musicnn", []), ("lyrics_mpnet", []), ("jukebox", []),
                                ("genres", [])],
                               [512, 512, 512, 256, 256], 512, single_branch_input_dropout=2e-2,
                               normalize_single_branch_input=True, aggregation_fn="mean",
                               embedding_regularization_type="pairwise_single", central_modality="interactions",
                               train_modalities=["interactions", "genres", "jukebox"], apply_output_activation=True,
                               apply_batch_normalization=True, apply_batch_norm_every=2)),
        rec_loss="bpr", optimizer="adamw", lr=5e-5, wd=1e-3, batch=64, n_neg=3, steps=1, seeded_init=7),
    # conf/single/algorithms/sbnet_amazonvid2024_huge_no-user_conf.yml:30-63: plain user embedding (embedding_dim -1 -> D),
    # the same huge item branch on [interactions, title], missing-modality evaluation (eval_modalities = [title])
    "amazon_nouser_widths": dict(
        corpus=dict(shape=AMAZON_FIXTURE_SHAPE, split_type="random", seed=23, scale=1.0),
        model=dict(shared_common_dim=128,
                   user=dict(feature_name="user_embedding", embedding_dim=-1, activation_fn="relu"),
                   item=entity([("interactions", []), ("title_mpnet", []), ("description_mpnet", []),
                                ("image_resnet", [])],
                               [512, 512, 512, 256, 256], 512, single_branch_input_dropout=2e-2,
                               normalize_single_branch_input=True, aggregation_fn="mean",
                               embedding_regularization_type="pairwise_single", central_modality="interactions",
                               train_modalities=["interactions", "title_mpnet"], eval_modalities=["title_mpnet"],
                               apply_output_activation=True, apply_batch_normalization=True,
                               apply_batch_norm_every=2)),
        rec_loss="bpr", optimizer="adamw", lr=5e-5, wd=1e-3, batch=48, n_neg=4, steps=1, seeded_init=11),
}


def seeded_state(shapes: dict, seed: int) -> dict:
    """deterministic fp32 ``state_dict`` values from names + shapes alone (numpy only, so the tests regenerate it
    instead of the fixture storing ~2 M floats): 2-D weights uniform with variance 1/fan_in, embedding tables N(0, 0.1),
    biases N(0, 0.05), BatchNorm gamma 1 + N(0, 0.1), running mean N(0, 0.05), running var 1 + U(0, 0.2)."""
    import zlib
    out = {}
    for name in sorted(shapes):
        shp = tuple(int(x) for x in shapes[name])
        rng = np.random.default_rng([int(seed), zlib.crc32(name.encode())])
        if name.endswith("num_batches_tracked"):
            v = np.zeros(shp, dtype=np.int64)
        elif name.endswith("running_var"):
            v = 1.0 + 0.2 * rng.random(shp)
        elif name.endswith("running_mean"):
            v = 0.05 * rng.standard_normal(shp)
        elif "embedding_layer" in name:
            v = 0.1 * rng.standard_normal(shp)
        elif len(shp) == 2:
            v = (rng.random(shp) * 2.0 - 1.0) * np.sqrt(3.0 / shp[1])
        elif "batch_norm" in name and name.endswith(".weight") or (len(shp) == 1 and name.endswith(".weight")):
            v = 1.0 + 0.1 * rng.standard_normal(shp)
        else:
            v = 0.05 * rng.standard_normal(shp)
        out[name] = v if v.dtype == np.int64 else v.astype(np.float32)
    return out


def pack_f16(a):
    """fp32 array -> (fp16 mantissas, fp32 scale): |a| / scale <= 1024, so entries down to 6e-8 of the largest one keep
    11 significant bits"""
    a = np.asarray(a, np.float32)
    m = float(np.abs(a).max())
    scale = np.float32(m / 1024.0 if m > 0 else 1.0)
    return (a / scale).astype(np.float16), scale


def unpack_f16(q, scale):
    return q.astype(np.float32) * np.float32(scale)


def build_reference_datasets(corpus):
    """Reference ``TrainRecDataset`` / ``FullEvalDataset`` instances without files."""
    ref_shims.install()
    from data.Feature import Feature
    from data.config_classes import FeatureDefinition, FeatureType
    from data.dataset import TrainRecDataset, FullEvalDataset

    def conv(feats):
        out = {}
        for name, f in feats.items():
            fd = FeatureDefinition(name, FeatureType(f.feature_definition.type),
                                   tag_split_sep=f.feature_definition.tag_split_sep)
            out[name] = Feature(fd, raw_values=f.raw_values, indices=f._indices)
        return out

    class _Csr:  # scipy >= 1.1x refuses csr[torch.Tensor] (eval/eval.py:219)
        def __init__(self, m):
            self.m = m

        def __getitem__(self, idx):
            return self.m[np.asarray(idx)]

    res = {}
    for split in ("train", "val", "test"):
        syn = corpus.dataset(split)
        cls = TrainRecDataset if split == "train" else FullEvalDataset
        ds = cls.__new__(cls)
        for k, v in syn.__dict__.items():
            if k != "name":
                setattr(ds, k, v)
        ds.user_features = conv(corpus.user_features)
        ds.item_features = conv(corpus.item_features)
        ds.user_feature_definitions = [f.feature_definition for f in ds.user_features.values()]
        if split != "train":
            ds.exclude_data = _Csr(syn.exclude_data)
        res[split] = ds
    return res


def run_case(name, spec):
    import torch
    ref_shims.install()
    from sibrar_b200.synthetic import SynCorpus, sample_batch
    from algorithms.sgd_alg import SingleBranchNet
    from train.rec_losses import RecommenderSystemLossesEnum
    from eval.eval import FullEvaluator, evaluate_recommender_algorithm
    from data.config_classes import EvalConfig

    torch.manual_seed(1234)
    torch.set_num_threads(1)
    corpus = SynCorpus(**spec["corpus"])
    dss = build_reference_datasets(corpus)
    train_ds = dss["train"]
    model = SingleBranchNet.build_from_conf(copy.deepcopy(spec["model"]), train_ds)
    out = {}
    seeded = spec.get("seeded_init")
    if seeded is not None:
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in seeded_state(shapes, seeded).items()})
        for k, shp in shapes.items():
            out[f"shape/{k}"] = np.asarray(shp, dtype=np.int64)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    if seeded is None:
        for k, v in sd0.items():
            out[f"sd0/{k}"] = v.numpy()

    # ---- hooks: record modalities and dropout masks
    rec = {"mods": {}, "drop": {}}
    for ent_name, ent in (("user", model.user_embedding_module), ("item", model.item_embedding_module)):
        if not hasattr(ent, "_sample_modalities"):
            continue
        orig = ent._sample_modalities

        def hooked(indices, _orig=orig, _n=ent_name):
            m = _orig(indices)
            rec["mods"][_n] = m
            return m
        ent._sample_modalities = hooked
        for mod in ent.sb_net:
            if isinstance(mod, torch.nn.Dropout):
                def drop_fwd(x, _m=mod, _n=ent_name):
                    if not _m.training or _m.p == 0:
                        rec["drop"][_n] = np.ones(x.shape, dtype=np.float32)
                        return x
                    keep = (torch.rand_like(x) >= _m.p).float()
                    rec["drop"][_n] = keep.numpy().copy()
                    return x * keep / (1 - _m.p)
                mod.forward = drop_fwd

    loss_fn = RecommenderSystemLossesEnum[spec["rec_loss"]].value(
        n_items=train_ds.n_items, aggregator=spec.get("aggregator", "mean"),
        train_neg_strategy=spec.get("neg_strategy", "uniform_recbole"), neg_train=spec["n_neg"])
    opt_cls = {"adam": torch.optim.Adam, "adamw": torch.optim.AdamW}[spec["optimizer"]]
    opt = opt_cls(model.parameters(), lr=spec["lr"], weight_decay=spec["wd"])
    rng = np.random.default_rng(99)
    model.train()
    for step in range(spec["steps"]):
        u, i = sample_batch(corpus.dataset("train"), spec["batch"], rng, spec["n_neg"])
        ut, it = torch.from_numpy(u), torch.from_numpy(i)
        labels = torch.zeros(i.shape, dtype=torch.float64)
        labels[:, 0] = 1.
        logits = model(ut, it)
        logits.retain_grad()
        out[f"s{step}/u"] = u
        out[f"s{step}/i"] = i
        out[f"s{step}/logits"] = logits.detach().numpy().copy()
        rec_loss = loss_fn.compute_loss(logits, labels)
        reg = model.get_and_reset_other_loss()
        total = rec_loss + reg["reg_loss"].to(rec_loss.device)
        out[f"s{step}/rec_loss"] = np.float64(rec_loss.item())
        for k, v in reg.items():
            out[f"s{step}/{k}"] = np.float64(v.item())
        out[f"s{step}/loss"] = np.float64(total.item())
        total.backward()
        for ent_name in ("user", "item"):
            if ent_name in rec["mods"]:
                m = rec["mods"][ent_name]
                names = sorted(set(m.reshape(-1).tolist()) | set(
                    getattr(model, f"{ent_name}_embedding_module").train_modalities))
                ids = np.vectorize({n: j for j, n in enumerate(names)}.__getitem__)(m).astype(np.int8)
                out[f"s{step}/mods_{ent_name}"] = ids
                out[f"s{step}/mod_names_{ent_name}"] = np.array(names)
            if ent_name in rec["drop"]:
                out[f"s{step}/drop_{ent_name}"] = rec["drop"][ent_name].astype(np.uint8)
        for k, p in model.named_parameters():
            gnp = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
            if seeded is None:
                out[f"s{step}/grad/{k}"] = gnp
            else:
                out[f"s{step}/grad16/{k}"], out[f"s{step}/gscale/{k}"] = pack_f16(gnp)
        if seeded is None:
            opt.step()
        opt.zero_grad()
        for k, v in model.state_dict().items():
            # seeded cases: the optimizer is not applied (Adam is pinned by the small cases); only what the step itself
            # changed -- the BatchNorm running statistics -- is stored, the tests overlay it on the seeded weights
            if seeded is None or k.endswith(("running_mean", "running_var", "num_batches_tracked")):
                out[f"s{step}/sd/{k}"] = v.detach().numpy().copy()
        rec["mods"].clear()
        rec["drop"].clear()

    # ---- eval on the val split with the trained weights (eval mode, BN running stats)
    from torch.utils.data import DataLoader
    val = dss["val"]
    ev_conf = EvalConfig(top_k=[1, 3, 5], metrics=["ndcg", "precision", "recall", "f_score", "hitrate", "coverage"],
                         calculate_std=False)
    evaluator = FullEvaluator(ev_conf, dataset=val)
    loader = DataLoader(val, batch_size=7, shuffle=False)
    model.eval()
    with torch.no_grad():
        i_idx = torch.tensor(val.items_in_split)
        i_repr = model.get_item_representations(i_idx)
        u_idx = torch.tensor(val.users_in_split)
        u_repr = model.get_user_representations(u_idx)
        scores = model.combine_user_item_representations(u_repr, i_repr)
        mask = torch.tensor(val.exclude_data[u_idx].toarray(), dtype=torch.bool)
        scores[mask] = -torch.inf
    out["eval/i_repr"] = i_repr.numpy()
    out["eval/u_repr"] = u_repr.numpy()
    out["eval/scores"] = scores.numpy()
    top = torch.topk(scores, 5, dim=-1)
    out["eval/topk_idx"] = top.indices.numpy()
    out["eval/topk_val"] = top.values.numpy()
    metrics, raw = evaluate_recommender_algorithm(model, loader, evaluator, device="cpu", return_raw=True)
    for k, v in metrics.items():
        out[f"eval/metric/{k}"] = np.float64(v)
    for k, v in raw.items():
        out[f"eval/raw/{k}"] = np.asarray(v)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"[golden] {name}: {len(out)} arrays -> {path} ({os.path.getsize(path) / 1024:.0f} KiB); "
          f"losses {[float(out[f's{s}/loss']) for s in range(spec['steps'])]}  ndcg@5 {metrics.get('ndcg@5')}")


def main():
    import sibrar_b200  # noqa: F401  (registers the package alias)
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        run_case(name, spec)


if __name__ == "__main__":
    main()
