"""TEST INFRASTRUCTURE ONLY -- golden fixtures for the on-disk dataset loader.

Run in the build container (needs ``/root/reference``):  ``python -m oracle.make_disk_golden``.

Writes three tiny preprocessed datasets (item cold start, random split, user cold start) in the reference's on-disk format (``data/data_preprocessing_utils.py:391-416``:
``user_idxs.csv``, ``item_idxs.csv``, ``listening_history_{split}.csv``, ``{entity}_features_{split}.csv``,
``{entity}_{feature}_{split}.npz``, ``used_config.yaml``) under ``tests/golden/disk_<case>/`` and loads every split
with the UNMODIFIED reference classes (``data/dataset.py``: ``TrainRecDataset`` for train, ``FullEvalDataset`` for
val / test).  What those objects expose -- index sets, the sparse matrices, the exclusion mask, every feature's
``_indices`` / ``values`` / ``dim`` / ``unique_values`` -- goes into ``tests/golden/disk_<case>_expected.npz``;
``tests/test_disk_dataset.py`` holds ``sibrar_b200.disk_dataset`` to exactly these.
"""
from __future__ import annotations

import os
import shutil
import sys

import numpy as np
import pandas as pd
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GENRES = ["action", "comedy", "drama", "horror", "romance", "scifi", "western"]
USER_FEATURES = [dict(name="gender", type="categorical"), dict(name="age", type="discrete"),
                 dict(name="taste", type="vector")]
ITEM_FEATURES = [dict(name="genres", type="tag", tag_split_sep="|"), dict(name="year", type="continuous"),
                 dict(name="studio", type="categorical"), dict(name="plot", type="vector")]
CASES = {"cs_item": dict(cold_start="item", seed=11, n_users=37, n_items=29),
         "random": dict(cold_start=None, seed=12, n_users=23, n_items=31),
         "cs_user": dict(cold_start="user", seed=14, n_users=41, n_items=19)}


def write_case(name: str, spec: dict) -> str:
    rng = np.random.default_rng(spec["seed"])
    U, I = spec["n_users"], spec["n_items"]
    out = os.path.join(GOLDEN_DIR, f"disk_{name}")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    pd.DataFrame(dict(user=[f"u{100 + u}" for u in range(U)], user_idx=np.arange(U))).to_csv(
        os.path.join(out, "user_idxs.csv"), index=False)
    pd.DataFrame(dict(item=[f"i{500 + i}" for i in range(I)], item_idx=np.arange(I))).to_csv(
        os.path.join(out, "item_idxs.csv"), index=False)
    # interactions: every user 4..9 distinct items (plus one DUPLICATED train row: the reference sums duplicates)
    pairs = [(u, int(i)) for u in range(U) for i in rng.choice(I, size=int(rng.integers(4, 10)), replace=False)]
    pairs = np.array(pairs)
    if spec["cold_start"] in ("item", "user"):
        n_ent, col = (I, 1) if spec["cold_start"] == "item" else (U, 0)
        perm = rng.permutation(n_ent)
        owner = np.empty(n_ent, dtype=int)
        owner[perm[:int(0.6 * n_ent)]] = 0
        owner[perm[int(0.6 * n_ent):int(0.8 * n_ent)]] = 1
        owner[perm[int(0.8 * n_ent):]] = 2
        part = owner[pairs[:, col]]
    else:
        part = rng.choice(3, size=len(pairs), p=[0.7, 0.15, 0.15])
    splits = {s: pairs[part == k] for k, s in enumerate(("train", "val", "test"))}
    splits["train"] = np.concatenate([splits["train"], splits["train"][:1]])  # the duplicate
    for s, p in splits.items():
        p = p[rng.permutation(len(p))]
        pd.DataFrame(dict(user_idx=p[:, 0], item_idx=p[:, 1], timestamp=rng.integers(0, 10 ** 6, len(p)))).to_csv(
            os.path.join(out, f"listening_history_{s}.csv"), index=False)
        splits[s] = p
    # entity features, stored per split for the entities of that split (rows deliberately NOT sorted by index)
    gender = rng.choice(["f", "m", "x"], size=U, p=[0.45, 0.45, 0.10])
    age = rng.integers(18, 70, size=U)
    taste = rng.normal(size=(U, 5)).astype(np.float32)
    genres = ["|".join(sorted(rng.choice(GENRES, size=int(rng.integers(1, 4)), replace=False))) for _ in range(I)]
    year = np.round(rng.uniform(1950, 2020, size=I), 1)
    studio = rng.choice(["a24", "mgm", "ufa", "toho"], size=I)
    plot = rng.normal(size=(I, 6)).astype(np.float32)
    for s, p in splits.items():
        us = rng.permutation(np.unique(p[:, 0]))
        its = rng.permutation(np.unique(p[:, 1]))
        pd.DataFrame(dict(user=[f"u{100 + u}" for u in us], user_idx=us, gender=gender[us], age=age[us])).to_csv(
            os.path.join(out, f"user_features_{s}.csv"), index=False)
        pd.DataFrame(dict(item=[f"i{500 + i}" for i in its], item_idx=its, genres=[genres[i] for i in its],
                          year=year[its], studio=studio[its])).to_csv(
            os.path.join(out, f"item_features_{s}.csv"), index=False)
        np.savez(os.path.join(out, f"user_taste_{s}.npz"), indices=us, values=taste[us])
        np.savez(os.path.join(out, f"item_plot_{s}.npz"), indices=its, values=plot[its])
    cfg = dict(split=dict(ratios=[0.6, 0.2, 0.2], split_type="coldstart" if spec["cold_start"] else "random",
                          cold_start_type=spec["cold_start"], seed=spec["seed"]),
               interactions=dict(k_core=1, min_n_interactions=1), user_features=[], item_features=[])
    with open(os.path.join(out, "used_config.yaml"), "w") as fh:
        yaml.safe_dump(cfg, fh)
    return out


def csr_parts(m, prefix, store):
    m = m.tocsr()
    m.sort_indices()
    store[prefix + "indptr"], store[prefix + "indices"] = m.indptr.astype(np.int64), m.indices.astype(np.int64)
    store[prefix + "data"] = np.asarray(m.data).astype(np.int64)
    store[prefix + "shape"] = np.asarray(m.shape)


def dump_features(feats: dict, prefix: str, store: dict):
    for fname, f in feats.items():
        p = f"{prefix}{fname}/"
        store[p + "indices"] = np.asarray(f._indices).astype(np.int64)
        store[p + "values"] = np.asarray(f.values)
        store[p + "dim"] = np.asarray(f.dim)
        if f._unique_values is not None:
            store[p + "unique_values"] = np.asarray(list(f._unique_values), dtype=str)


def load_with_reference(path: str) -> dict:
    from data.config_classes import FeatureDefinition, InteractionDatasetConfig, TrainDatasetConfig
    from data.dataset import FullEvalDataset, TrainRecDataset
    ufd = [FeatureDefinition.from_dict(d) for d in USER_FEATURES]
    ifd = [FeatureDefinition.from_dict(d) for d in ITEM_FEATURES]
    store = {}
    for split in ("train", "val", "test"):
        common = dict(split_set=split, dataset_path=path, user_feature_definitions=ufd, item_feature_definitions=ifd,
                      model_requires_train_interactions=True, model_requires_item_interactions=True)
        if split == "train":
            ds = TrainRecDataset(TrainDatasetConfig(n_negative_samples=3, negative_sampling_strategy="uniform_recbole",
                                                    use_dataset_negative_sampler=False, **common))
        else:
            ds = FullEvalDataset(InteractionDatasetConfig(**common))
        p = f"{split}/"
        for k in ("n_users", "n_items", "n_interactions", "n_users_in_split", "n_items_in_split", "is_cold_start_user",
                  "is_cold_start_item"):
            store[p + k] = np.asarray(getattr(ds, k))
        store[p + "users_in_split"] = np.asarray(ds.users_in_split).astype(np.int64)
        store[p + "items_in_split"] = np.asarray(ds.items_in_split).astype(np.int64)
        store[p + "coo_row"] = np.asarray(ds.interaction_matrix.row).astype(np.int64)
        store[p + "coo_col"] = np.asarray(ds.interaction_matrix.col).astype(np.int64)
        csr_parts(ds.user_sampling_matrix, p + "usm/", store)
        csr_parts(ds.user_sampling_matrix_train, p + "usm_train/", store)
        csr_parts(ds.item_sampling_matrix_train, p + "ism_train/", store)
        if split != "train":
            csr_parts(ds.exclude_data, p + "exclude/", store)
        dump_features(ds.user_features, p + "user/", store)
        dump_features(ds.item_features, p + "item/", store)
    return store


def main():
    ref_shims.install()
    for name, spec in CASES.items():
        path = write_case(name, spec)
        store = load_with_reference(path)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"disk_{name}_expected.npz"), **store)
        print(f"[disk golden] {name}: {len(store)} arrays, {sum(os.path.getsize(os.path.join(path, f)) for f in os.listdir(path))} B of dataset files")


if __name__ == "__main__":
    main()
