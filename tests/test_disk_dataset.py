"""On-disk dataset loader against what the UNMODIFIED reference classes expose for the
same files (fixtures: ``oracle/make_disk_golden.py`` -> ``tests/golden/disk_<case>/`` + ``disk_<case>_expected.npz``),
plus -- on the GPU -- a fused train step and an evaluation on a model built from such a directory."""
import os

import numpy as np
import pytest

from sibrar_b200.disk_dataset import DiskCorpus, DiskDataset, DiskFeature, FeatureDefinition

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
USER_FEATURES = [dict(name="gender", type="categorical"), dict(name="age", type="discrete"),
                 dict(name="taste", type="vector")]
ITEM_FEATURES = [dict(name="genres", type="tag", tag_split_sep="|"), dict(name="year", type="continuous"),
                 dict(name="studio", type="categorical"), dict(name="plot", type="vector")]


def _csr_equal(m, g, prefix):
    m = m.tocsr()
    m.sort_indices()
    assert tuple(m.shape) == tuple(g[prefix + "shape"])
    assert np.array_equal(m.indptr, g[prefix + "indptr"])
    assert np.array_equal(m.indices, g[prefix + "indices"])
    assert np.array_equal(np.asarray(m.data).astype(np.int64), g[prefix + "data"])


@pytest.mark.parametrize("case", ["cs_item", "random", "cs_user"])
@pytest.mark.parametrize("split", ["train", "val", "test"])
def test_loader_matches_reference_dataset(case, split):
    g = np.load(os.path.join(GOLDEN, f"disk_{case}_expected.npz"))
    ds = DiskDataset(os.path.join(GOLDEN, f"disk_{case}"), split, USER_FEATURES, ITEM_FEATURES, n_negative_samples=3,
                     negative_sampling_strategy="uniform_recbole")
    p = f"{split}/"
    for k in ("n_users", "n_items", "n_interactions", "n_users_in_split", "n_items_in_split", "is_cold_start_user",
              "is_cold_start_item"):
        assert getattr(ds, k) == g[p + k].item(), k
    assert np.array_equal(ds.users_in_split, g[p + "users_in_split"])
    assert np.array_equal(ds.items_in_split, g[p + "items_in_split"])
    assert np.array_equal(ds.interaction_matrix.row, g[p + "coo_row"])  # COO keeps the history order (and duplicates)
    assert np.array_equal(ds.interaction_matrix.col, g[p + "coo_col"])
    _csr_equal(ds.user_sampling_matrix, g, p + "usm/")
    _csr_equal(ds.user_sampling_matrix_train, g, p + "usm_train/")
    _csr_equal(ds.item_sampling_matrix_train, g, p + "ism_train/")
    if split != "train":
        _csr_equal(ds.exclude_data, g, p + "exclude/")
    else:
        assert ds.exclude_data.nnz == 0 and ds.exclude_data.shape == (ds.n_users, ds.n_items_in_split)
        assert ds.user_sampling_matrix_train.data.max() == 2  # the duplicated history row is summed, like the reference
    assert len(ds) == (g[p + "n_interactions"].item() if split == "train" else g[p + "n_users_in_split"].item())
    for entity, feats in (("user", ds.user_features), ("item", ds.item_features)):
        for name, f in feats.items():
            q = f"{p}{entity}/{name}/"
            assert np.array_equal(np.asarray(f._indices), g[q + "indices"]), (entity, name)
            want = g[q + "values"]
            got = np.asarray(f.values)
            if f.feature_definition.type == "tag":
                # the reference orders the tags of a row by string hash (set iteration): compare as sorted rows
                assert got.shape == want.shape and np.array_equal(np.sort(got, axis=1), np.sort(want, axis=1))
            elif got.dtype.kind == "f":
                assert got.shape == want.shape and np.allclose(got, want, rtol=0, atol=0)
            else:
                assert np.array_equal(got, want), (entity, name)
            assert np.array_equal(np.asarray(f.dim), g[q + "dim"])
            if q + "unique_values" in g:
                assert list(map(str, f.unique_values)) == list(g[q + "unique_values"])


def test_train_split_sees_train_and_val_rows_and_full_vocabulary():
    ds = DiskDataset(os.path.join(GOLDEN, "disk_cs_item"), "train", USER_FEATURES, ITEM_FEATURES)
    val = DiskDataset(os.path.join(GOLDEN, "disk_cs_item"), "val", USER_FEATURES, ITEM_FEATURES)
    test = DiskDataset(os.path.join(GOLDEN, "disk_cs_item"), "test", USER_FEATURES, ITEM_FEATURES)
    items = set(ds.item_features["plot"]._indices.tolist())
    assert set(ds.items_in_split.tolist()) | set(val.items_in_split.tolist()) == items  # cold-start: disjoint item sets
    assert not items & set(test.items_in_split.tolist())
    assert np.all(np.diff(ds.item_features["plot"]._indices) > 0)  # sorted by entity index
    # vocabularies are those of all three splits, in every split view
    assert ds.item_features["genres"].unique_values == test.item_features["genres"].unique_values
    assert ds.item_features["studio"].n_unique_categories == test.item_features["studio"].n_unique_categories
    # host lookup by entity index; an entity without a row raises like data/Feature.py:146
    e = int(ds.item_features["plot"]._indices[3])
    assert np.array_equal(ds.item_features["plot"][np.array([e])][0], ds.item_features["plot"].values[3])
    with pytest.raises(KeyError):
        ds.item_features["plot"][np.array([int(test.items_in_split[0])])]


def test_feature_typing_rules_and_errors():
    fd = FeatureDefinition.from_dict
    f = DiskFeature(fd(dict(name="g", type="categorical")), ["b", "a", "b"], reference_values=["c", "a"])
    assert f.unique_values == ["a", "b", "c"] and f.values.tolist() == [1, 0, 1] and f.dim == 0
    assert f.n_unique_categories == 3
    f = DiskFeature(fd(dict(name="t", type="tag", tag_split_sep="|")), ["x|y", "z", "y"], reference_values=["w"])
    assert f.unique_values == ["w", "x", "y", "z"] and f.dim == 4 and f.values.tolist() == [[1, 2], [3, 4], [2, 4]]
    with pytest.raises(TypeError):
        f.n_unique_categories
    with pytest.raises(ValueError):
        DiskFeature(fd(dict(name="t", type="tag")), ["x|y"])
    f = DiskFeature(fd(dict(name="s", type="sequence")), ["[1, 2, 3]", "[4, 5, 6]"])
    assert f.dim == 3 and f.values.tolist() == [[1, 2, 3], [4, 5, 6]]
    f = DiskFeature(fd(dict(name="v", type="vector")), [np.ones(4), np.zeros(4)], indices=np.array([7, 2]))
    assert f.dim == 4 and f.values.shape == (2, 4) and f[7].tolist() == [1, 1, 1, 1]
    f = DiskFeature(fd(dict(name="m", type="matrix")), np.zeros((3, 2, 5)))
    assert f.dim == (2, 5)
    f = DiskFeature(fd(dict(name="g", type="categorical", preprocessing="one_hot")), ["b", "a"])
    assert f.dim == 2 and f.values.tolist() == [[0, 1], [1, 0]]
    f = DiskFeature(fd(dict(name="t", type="tag", tag_split_sep=",", preprocessing="multi_hot")), ["x,y", "y"])
    assert f.values.tolist() == [[1, 1], [0, 1]]
    with pytest.raises(ValueError):
        fd(dict(name="q", type="image"))
    with pytest.raises(ValueError):
        DiskFeature(fd(dict(name="a", type="discrete")), [1, 2, 3], indices=np.arange(2))
    with pytest.raises(FileNotFoundError):
        DiskDataset(os.path.join(GOLDEN, "disk_random"), "val", [dict(name="missing", type="vector")], [])
    with pytest.raises(ValueError):
        DiskDataset(os.path.join(GOLDEN, "disk_random"), "val", [dict(name="nope", type="discrete")], [])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["cs_item", "random"])
def test_model_trains_and_evaluates_on_a_dataset_directory(case):
    import torch
    from sibrar_b200.evaluator import FullEvaluator
    from sibrar_b200.sbnet import SingleBranchNet
    from sibrar_b200.trainer import FusedTrainer
    corpus = DiskCorpus(os.path.join(GOLDEN, f"disk_{case}"), USER_FEATURES, ITEM_FEATURES, n_negative_samples=3,
                        negative_sampling_strategy="uniform_recbole")
    train, val = corpus.dataset("train"), corpus.dataset("val")
    ent = lambda feats, hidden: dict(features=[dict(feature_name=f, feature_hidden_layers=[]) for f in feats],  # noqa
                                     single_branch_hidden_layers=hidden, preference_hidden_layers=[],
                                     common_modality_dim=16, activation_fn="relu")
    conf = dict(shared_common_dim=16, user=ent(["interactions", "gender", "age", "taste"], []),
                item=ent(["interactions", "genres", "year", "studio", "plot"], [16]))
    torch.manual_seed(0)
    model = SingleBranchNet.build_from_conf(conf, train).to("cuda").train()
    tr = FusedTrainer(model, dict(lr=3e-3, wd=0.0, optimizer="adam", rec_loss="bpr", loss_aggregator="mean"),
                      n_negative_samples=3)
    rng = np.random.default_rng(0)
    coo = train.interaction_matrix
    losses = []
    for _ in range(60):
        pick = rng.integers(0, coo.nnz, size=32)
        u = coo.row[pick].astype(np.int64)
        neg = rng.choice(train.items_in_split, size=(32, 3))
        i = np.concatenate([coo.col[pick].astype(np.int64)[:, None], neg], axis=1)
        tr.step(torch.from_numpy(u).cuda(), torch.from_numpy(i).cuda())
        losses.append(tr.read_losses()["train/loss"])
    model.check_errors()
    # (32-interaction batches of a ~100-interaction corpus: single steps are noisy, the trend is what is checked)
    assert np.isfinite(losses).all() and np.mean(losses[-15:]) < np.mean(losses[:3])
    # evaluation on the val split: a model for evaluation is built from the EVAL dataset (its feature tables), the
    # weights come from the trained one (experiment_helper.py:132)
    ev_model = SingleBranchNet.build_from_conf(conf, val).to("cuda")
    ev_model.load_state_dict(model.state_dict())
    res = FullEvaluator(dict(top_k=[1, 3], metrics=["ndcg", "recall", "coverage"], calculate_std=False)).evaluate(
        ev_model, val)
    assert set(res) == {"ndcg@1", "ndcg@3", "recall@1", "recall@3", "coverage@1", "coverage@3"}
    assert all(0.0 <= v <= 1.0 for v in res.values())
