"""Seeded synthetic datasets shaped like ML-1M / Onion18 / AmazonVideo2024 (there is no network for the real ones).

The objects produced here expose exactly the attributes the reference's ``SingleBranchNet.__init__`` /
``evaluate_recommender_algorithm`` read from ``TrainRecDataset`` / ``FullEvalDataset``
(reference ``algorithms/sgd_alg.py:2017-2075``, ``eval/eval.py:205-222``, ``data/dataset.py:55-56,127-134,251-255,
412-453``): ``n_users, n_items, user_features, item_features, user_sampling_matrix_train,
item_sampling_matrix_train, is_cold_start_user/item, items_in_split, users_in_split, exclude_data,
user_sampling_matrix, n_items_in_split, n_users_in_split, interaction_matrix``.

Features mimic ``data/Feature.py:27-295``: ``.feature_definition.{name,type}``, ``.values``, ``.dim``, ``._indices``,
``.n_unique_categories``.  ``raw_values`` are kept so the test-only oracle can build real reference ``Feature``
objects from the same data.
"""
from __future__ import annotations

import dataclasses
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp

CATEGORICAL, TAG, DISCRETE, CONTINUOUS, VECTOR = "categorical", "tag", "discrete", "continuous", "vector"


class SynFeature:
    """Feature container with the attribute surface of the reference ``Feature`` (``data/Feature.py:93-114``)."""

    def __init__(self, name: str, ftype: str, raw_values, indices=None, tag_split_sep: str = None):
        self.feature_definition = SimpleNamespace(name=name, type=ftype, tag_split_sep=tag_split_sep)
        self.raw_values = raw_values
        n = raw_values.shape[0] if hasattr(raw_values, "shape") else len(raw_values)
        self._n_values = n
        self._indices = np.arange(n) if indices is None else np.asarray(indices)
        self._unique_values = None
        if ftype == CATEGORICAL:
            self._unique_values = sorted(set(np.asarray(raw_values).tolist()))
            vmap = {v: i for i, v in enumerate(self._unique_values)}
            self._values = np.array([vmap[v] for v in np.asarray(raw_values).tolist()], dtype=np.int64)
            self._dim = 0
        elif ftype == TAG:
            tags = [set(v.split(tag_split_sep)) for v in raw_values]
            self._unique_values = sorted(set().union(*tags))
            vmap = {v: i for i, v in enumerate(self._unique_values)}
            lists = [[vmap[t] for t in tg] for tg in tags]
            width = max(map(len, lists))
            pad = len(self._unique_values)
            self._values = np.array([li + [pad] * (width - len(li)) for li in lists], dtype=np.int64)
            self._dim = len(self._unique_values)
        elif ftype in (DISCRETE, CONTINUOUS):
            self._values = np.asarray(raw_values)
            self._dim = 1
        elif ftype == VECTOR:
            self._values = raw_values
            self._dim = raw_values.shape[1]
        else:
            raise ValueError(f"FeatureType '{ftype}' is not supported")

    values = property(lambda self: self._values)
    dim = property(lambda self: self._dim)
    n_values = property(lambda self: self._n_values)

    @property
    def n_unique_categories(self):
        if self.feature_definition.type != CATEGORICAL:
            raise TypeError("only categorical features support n_unique_categories")
        return len(self._unique_values)

    def __len__(self):
        return self._n_values


@dataclasses.dataclass
class Shape:
    n_users: int
    n_items: int
    n_interactions: int
    user_feats: dict  # name -> (type, param)
    item_feats: dict


SHAPES = {
    # BASELINE.json configs[0]/[1]: ML-1M (6040 x 3706, ~1M interactions, ID + genre + 768-d text)
    "ml1m": Shape(6040, 3706, 1_000_209,
                  {"gender": (CATEGORICAL, 2), "occupation": (CATEGORICAL, 21), "age": (CONTINUOUS, None)},
                  {"genres": (TAG, 18), "plot_mpnet": (VECTOR, 768)}),
    # configs[2]: Onion18 shape (music: ID + audio + text, ~1e5 items)
    "onion18": Shape(50_000, 100_000, 5_000_000,
                     {"gender": (CATEGORICAL, 3), "country": (CATEGORICAL, 152), "age": (CONTINUOUS, None),
                      "mpnet": (VECTOR, 768)},
                     {"genres": (TAG, 853), "jukebox": (VECTOR, 4800), "musicnn": (VECTOR, 50),
                      "lyrics_mpnet": (VECTOR, 768)}),
    # configs[3]: AmazonVideo2024 shape (image + text + ID), paper size
    "amazonvid2024": Shape(11_454, 4_177, 87_098,
                           {},
                           {"title_mpnet": (VECTOR, 768), "description_mpnet": (VECTOR, 768),
                            "image_resnet": (VECTOR, 2048)}),
}


def _make_feature(rng, name, ftype, param, n):
    if ftype == CATEGORICAL:
        vals = rng.integers(0, param, size=n)
        vals[:param] = np.arange(param)[:n]  # every category present
        return SynFeature(name, CATEGORICAL, vals)
    if ftype == TAG:
        n_tags = rng.integers(1, 7, size=n)
        raw = []
        for i in range(n):
            t = rng.choice(param, size=min(int(n_tags[i]), param), replace=False)
            raw.append("|".join(f"t{int(x):04d}" for x in t))
        # make sure all tags exist
        raw[0] = "|".join(f"t{int(x):04d}" for x in range(min(param, 6)))
        for j in range(6, param):
            raw[(j - 5) % n] += f"|t{j:04d}" if param <= 32 else ""
        if param > 32:
            for j in range(param):
                raw[(j * 7 + 1) % n] += f"|t{j:04d}"
        return SynFeature(name, TAG, raw, tag_split_sep="|")
    if ftype == CONTINUOUS:
        return SynFeature(name, CONTINUOUS, rng.normal(size=n).astype(np.float32))
    if ftype == VECTOR:
        v = rng.standard_normal(size=(n, param), dtype=np.float32)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        return SynFeature(name, VECTOR, v)
    raise ValueError(ftype)


def _interactions(rng, n_users, n_items, n_inter):
    """user activity ~ lognormal, item popularity ~ Zipf(1); deduplicated (u, i) pairs; every user >= 3 items."""
    pu = rng.lognormal(0., 1., size=n_users)
    pu /= pu.sum()
    pi = 1. / np.arange(1, n_items + 1) ** 0.9
    pi = pi[rng.permutation(n_items)]
    pi /= pi.sum()
    n_inter = min(n_inter, int(0.5 * n_users * n_items))
    keys = np.empty(0, dtype=np.int64)
    base_u = np.repeat(np.arange(n_users), 3)
    base_i = rng.integers(0, n_items, size=base_u.size)
    keys = np.unique(base_u.astype(np.int64) * n_items + base_i)
    while keys.size < n_inter:
        m = int((n_inter - keys.size) * 1.3) + 16
        u = rng.choice(n_users, size=m, p=pu)
        i = rng.choice(n_items, size=m, p=pi)
        keys = np.unique(np.concatenate([keys, u.astype(np.int64) * n_items + i]))
    if keys.size > n_inter:
        keys = np.sort(rng.choice(keys, size=n_inter, replace=False))
    return (keys // n_items).astype(np.int64), (keys % n_items).astype(np.int64)


class SynDataset:
    """One split view ('train' | 'val' | 'test') of a synthetic corpus."""

    def __init__(self, corpus: "SynCorpus", split_set: str):
        c = corpus
        self.name = f"{c.name}-{split_set}"
        self.split_set = split_set
        self.n_users, self.n_items = c.n_users, c.n_items
        self.is_cold_start_user = False
        self.is_cold_start_item = c.split_type == "cold_start_item"
        # fresh dicts: SingleBranchNet.__init__ mutates them (reference sgd_alg.py:2023-2032,2051-2059)
        self.user_features = dict(c.user_features)
        self.item_features = dict(c.item_features)
        self.user_feature_names = list(c.user_features)
        self.user_feature_definitions = [f.feature_definition for f in c.user_features.values()]
        self.user_sampling_matrix_train = c.train_csr
        self.item_sampling_matrix_train = c.train_csr_t
        u, i = c.split_pairs[split_set]
        self.interaction_matrix = sp.coo_matrix((np.ones(u.size, dtype=np.int8), (u, i)),
                                                shape=(c.n_users, c.n_items))
        self.user_sampling_matrix = self.interaction_matrix.tocsr()
        self.items_in_split = c.items_in_split[split_set]
        self.users_in_split = np.unique(u)
        self.n_items_in_split = len(self.items_in_split)
        self.n_users_in_split = len(self.users_in_split)
        self.n_negative_samples = c.n_negative_samples
        self.negative_sampling_strategy = "uniform_recbole"
        self.use_dataset_negative_sampler = False
        if split_set != "train":
            mask = c.train_csr.astype(bool)
            if split_set == "test":
                vu, vi = c.split_pairs["val"]
                mask = mask + sp.csr_matrix((np.ones(vu.size, dtype=bool), (vu, vi)), shape=mask.shape)
            self.exclude_data = mask[:, self.items_in_split].astype(bool).tocsr()
            self.exclude_data.sort_indices()
        else:
            self.exclude_data = sp.csr_matrix((c.n_users, self.n_items_in_split), dtype=bool)

    def __len__(self):
        return self.interaction_matrix.nnz if self.split_set == "train" else self.n_users_in_split


class SynCorpus:
    def __init__(self, shape: str | Shape = "ml1m", split_type: str = "random", seed: int = 42, scale: float = 1.0,
                 n_negative_samples: int = 10, vector_dim_cap: int = None):
        sh = SHAPES[shape] if isinstance(shape, str) else shape
        self.name = shape if isinstance(shape, str) else "custom"
        self.split_type = split_type
        rng = np.random.default_rng(seed)
        self.n_users = max(8, int(round(sh.n_users * scale)))
        self.n_items = max(8, int(round(sh.n_items * scale)))
        n_inter = max(self.n_users * 4, int(round(sh.n_interactions * scale * scale)))
        self.n_negative_samples = n_negative_samples

        def cap(ftype, p):
            return min(p, vector_dim_cap) if (ftype == VECTOR and vector_dim_cap) else p

        self.user_features = {k: _make_feature(rng, k, t, cap(t, p), self.n_users)
                              for k, (t, p) in sh.user_feats.items()}
        self.item_features = {k: _make_feature(rng, k, t, cap(t, p), self.n_items)
                              for k, (t, p) in sh.item_feats.items()}
        u, i = _interactions(rng, self.n_users, self.n_items, n_inter)
        self.n_interactions = u.size

        all_items = np.arange(self.n_items)
        if split_type == "random":
            r = rng.random(u.size)
            part = np.where(r < 0.8, 0, np.where(r < 0.9, 1, 2))
            self.items_in_split = {"train": all_items, "val": all_items, "test": all_items}
        elif split_type == "cold_start_item":
            # 80/10/10 of the *items* (reference data/data_preprocessing_utils.py:313-332)
            perm = rng.permutation(self.n_items)
            n_tr, n_va = int(0.8 * self.n_items), int(0.1 * self.n_items)
            item_part = np.empty(self.n_items, dtype=np.int64)
            item_part[perm[:n_tr]] = 0
            item_part[perm[n_tr:n_tr + n_va]] = 1
            item_part[perm[n_tr + n_va:]] = 2
            part = item_part[i]
            self.items_in_split = {s: np.sort(all_items[item_part == p]) for p, s in enumerate(("train", "val", "test"))}
        else:
            raise ValueError(f"split type {split_type} not supported by the synthetic generator")
        self.split_pairs = {s: (u[part == p], i[part == p]) for p, s in enumerate(("train", "val", "test"))}
        tu, ti = self.split_pairs["train"]
        self.train_csr = sp.csr_matrix((np.ones(tu.size, dtype=np.int8), (tu, ti)), shape=(self.n_users, self.n_items))
        self.train_csr.sort_indices()
        self.train_csr_t = self.train_csr.T.tocsr()
        self.train_csr_t.sort_indices()

    def dataset(self, split_set: str = "train") -> SynDataset:
        return SynDataset(self, split_set)


def sample_batch(ds: SynDataset, batch_size: int, rng: np.random.Generator, n_neg: int = None):
    """Host sampler with the reference collate contract (``data/dataloader.py:154-198``): ``u int64 [B]``,
    ``i int64 [B, 1+n_neg]`` (positive in column 0), negatives uniform with replacement from ``items_in_split``,
    re-drawn while they are train positives of that user ('uniform_recbole')."""
    n_neg = ds.n_negative_samples if n_neg is None else n_neg
    coo = ds.interaction_matrix
    sel = rng.integers(0, coo.nnz, size=batch_size)
    u = coo.row[sel].astype(np.int64)
    pos = coo.col[sel].astype(np.int64)
    neg = rng.choice(ds.items_in_split, size=(batch_size, n_neg), replace=True)
    csr = ds.user_sampling_matrix_train
    for _ in range(64):
        bad = np.asarray(csr[np.repeat(u, n_neg), neg.reshape(-1)]).reshape(batch_size, n_neg) != 0
        nb = int(bad.sum())
        if nb == 0:
            break
        neg[bad] = rng.choice(ds.items_in_split, size=nb, replace=True)
    items = np.concatenate([pos[:, None], neg], axis=1).astype(np.int64)
    return u, items
