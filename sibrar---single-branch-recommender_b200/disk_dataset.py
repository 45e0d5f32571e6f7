"""Loader for the reference's preprocessed on-disk dataset format.

A dataset directory written by the reference's preprocessing (``data/data_preprocessing_utils.py:391-416``) holds

    user_idxs.csv, item_idxs.csv                  at least the columns ``user_idx`` / ``item_idx``
    listening_history_{train,val,test}.csv        at least ``user_idx``, ``item_idx``
    {entity}_features_{split}.csv                 tabular features (categorical, tag, discrete, continuous, sequence)
    {entity}_{feature}_{split}.npz                ``indices`` + ``values`` of a vector / matrix feature
    used_config.yaml                              the preprocessing config (``split.cold_start_type`` is read)

``DiskDataset`` exposes one split of it with the attribute surface that ``SingleBranchNet``, ``FusedTrainer`` and
``FullEvaluator`` consume -- the one of the reference's ``TrainRecDataset`` / ``FullEvalDataset``
(``data/dataset.py:36-447``):  index sets, COO / CSR interaction matrices, the train matrices in both orientations, the
exclusion mask of the evaluation splits and typed ``Feature`` objects (``data/Feature.py:27-288``).  Host code only: the
device-resident copies are built once by ``feature_store.DeviceFeature`` from these objects.

Semantics kept from the reference (each checked against the unmodified reference classes on the fixtures of
``oracle/make_disk_golden.py``, ``tests/test_disk_dataset.py``):
  * cold-start datasets: ``users_in_split`` / ``items_in_split`` = sorted unique indices of the split's history; otherwise
    the index columns of ``user_idxs.csv`` / ``item_idxs.csv`` in file order (``data/dataset.py:127-130``);
  * matrices keep the full ``[n_users, n_items]`` shape; duplicate history rows stay separate COO entries and are summed
    by the CSR conversions (``:139, 250-254``);
  * the train split sees the features of train AND val rows, first occurrence wins, rows sorted by entity index
    (``:214-216``, ``data_preprocessing_utils.py:470-506``); categorical / tag vocabularies come from all three splits
    (``reference_values``, ``data/dataset.py:209-227``) and are sorted;
  * exclusion mask: val = train interactions, test = train + val, restricted to the columns ``items_in_split``
    (``:424-432``).
Deviation: the reference orders the tags of one row by Python's string-hash order (``data/Feature.py:252``: iteration
over a ``set``), i.e. differently from run to run; here they are sorted.  ``EmbeddingBag(mean)`` does not see the order.
"""
from __future__ import annotations

import os
from ast import literal_eval
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import scipy.sparse as sp

SPLITS = ("train", "val", "test")
TABULAR_TYPES = ("categorical", "tag", "discrete", "continuous", "sequence")
MULTI_D_TYPES = ("vector", "matrix")


@dataclass
class FeatureDefinition:
    """``data/config_classes.py:104-114``"""
    name: str
    type: str
    preprocessing: str = "none"
    tag_split_sep: Optional[str] = None

    def __post_init__(self):
        self.type = str(getattr(self.type, "value", self.type)).lower()
        self.preprocessing = str(getattr(self.preprocessing, "value", self.preprocessing) or "none").lower()
        if self.type not in TABULAR_TYPES + MULTI_D_TYPES:
            raise ValueError(f"FeatureType '{self.type}' is not supported")

    @classmethod
    def from_dict(cls, d):
        return d if isinstance(d, cls) else cls(**{k: d[k] for k in ("name", "type", "preprocessing", "tag_split_sep")
                                                   if k in d})


class DiskFeature:
    """Typed feature values of one entity kind (attribute surface of the reference ``Feature``)."""

    def __init__(self, definition: FeatureDefinition, raw_values, indices=None, reference_values=None):
        self.feature_definition = definition
        n = raw_values.shape[0] if hasattr(raw_values, "shape") else len(raw_values)
        self._n_values = int(n)
        self._indices = np.arange(n) if indices is None else np.asarray(indices)
        if len(self._indices) != n:
            raise ValueError(f"Provided indices must match size of supplied values ({n} != {len(self._indices)})")
        self._unique_values = self._value_map = None
        t = definition.type
        if t == "categorical":
            raw = list(np.asarray(raw_values).tolist())
            vocab = set(raw) | (set(np.asarray(reference_values).tolist()) if reference_values is not None else set())
            self._unique_values = sorted(vocab)
            self._value_map = {v: i for i, v in enumerate(self._unique_values)}
            codes = np.array([self._value_map[v] for v in raw], dtype=np.int64)
            if definition.preprocessing == "one_hot":
                self._values = np.eye(len(self._unique_values), dtype=np.float32)[codes]
                self._dim = len(self._unique_values)
            else:
                self._values, self._dim = codes, 0
        elif t == "tag":
            sep = definition.tag_split_sep
            if sep is None:
                raise ValueError(f'For tag feature "{definition.name}" a separator (tag_split_sep) for the individual '
                                 f'has to be provided. For genre tags "action|romance" this would be "|".')
            rows = [set(str(v).split(sep)) for v in raw_values]
            vocab = set().union(*rows) if rows else set()
            if reference_values is not None:
                vocab |= set().union(*[set(str(v).split(sep)) for v in reference_values])
            self._unique_values = sorted(vocab)
            self._value_map = {v: i for i, v in enumerate(self._unique_values)}
            pad = len(self._unique_values)
            lists = [sorted(self._value_map[tg] for tg in tags) for tags in rows]
            self._dim = pad
            if definition.preprocessing == "multi_hot":
                m = np.zeros((len(lists), pad), dtype=np.float32)
                for r, li in enumerate(lists):
                    m[r, li] = 1.0
                self._values = m
            else:
                width = max(map(len, lists)) if lists else 0
                self._values = np.array([li + [pad] * (width - len(li)) for li in lists], dtype=np.int64)
        elif t == "sequence":
            self._values = np.stack([np.asarray(literal_eval(v)) for v in raw_values], axis=0)
            self._dim = int(self._values.shape[1])
        elif t in ("discrete", "continuous"):
            self._values, self._dim = np.asarray(raw_values), 1
        else:  # vector | matrix
            self._values = np.stack(raw_values, axis=0) if isinstance(raw_values, list) else raw_values
            dim = tuple(self._values.shape[1:])
            self._dim = int(dim[0]) if len(dim) == 1 else dim

    values = property(lambda self: self._values)
    dim = property(lambda self: self._dim)
    n_values = property(lambda self: self._n_values)

    def _need(self, kinds, what):
        if self.feature_definition.type not in kinds:
            raise TypeError(f'Only features of type {list(kinds)} support "{what}"')

    @property
    def unique_values(self):
        self._need(("categorical", "tag"), "unique_values")
        return self._unique_values

    @property
    def n_unique_categories(self):
        self._need(("categorical",), "n_unique_categories")
        return len(self._unique_values)

    @property
    def value_map(self):
        self._need(("categorical", "tag"), "value_map")
        return self._value_map

    def __len__(self):
        return self._n_values

    def __getitem__(self, idx):
        """host lookup by ENTITY index (``data/Feature.py:140-162``; KeyError for an entity without a row)"""
        pos = {int(e): r for r, e in enumerate(self._indices)}
        if isinstance(idx, (int, np.integer)):
            return self._values[pos[int(idx)]]
        idx = np.asarray(idx)
        rows = np.array([pos[int(e)] for e in idx.reshape(-1)], dtype=np.int64)
        vals = self._values[rows]
        if sp.issparse(vals):
            vals = vals.toarray()
        return vals.reshape(idx.shape + (-1,)) if (isinstance(self._dim, tuple) or self._dim > 0) else vals

    def __repr__(self):
        d = self.feature_definition
        return f"Feature(name={d.name}, type={d.type}, number={self._n_values}, dim={self._dim})"


# ------------------------------------------------------------------------------------------------ files -> tables
def _read_entity_split(path: str, entity: str, defs: Sequence[FeatureDefinition], split: str):
    """(tabular DataFrame | None, {multi-d feature: (indices, values)}) of one split"""
    import pandas as pd
    tab_names = [d.name for d in defs if d.type in TABULAR_TYPES]
    table = None
    if tab_names:
        f = os.path.join(path, f"{entity}_features_{split}.csv")
        if not os.path.exists(f):
            raise FileNotFoundError(f'Feature file "{f}" does not exist')
        keep = [entity, f"{entity}_idx"] + tab_names
        table = pd.read_csv(f, usecols=lambda c: c in keep)
        missing = set(tab_names) - set(table.columns)
        if missing:
            raise ValueError(f"Column(s) for {entity} feature(s) {sorted(missing)} are missing.")
    multi = {}
    for d in defs:
        if d.type in MULTI_D_TYPES:
            f = os.path.join(path, f"{entity}_{d.name}_{split}.npz")
            if not os.path.exists(f):
                raise FileNotFoundError(f'Data file for {entity} feature "{d.name}" does not exist.')
            z = np.load(f, allow_pickle=True)
            ind, val = z["indices"], z["values"]
            if len(ind) != len(val):
                raise ValueError(f'Mismatch between number of {entity} indices and its "{d.name}" feature'
                                 f"({len(ind)} indices but {len(val)} feature values).")
            multi[d.name] = (ind, val)
    return table, multi


def _merge_splits(path: str, entity: str, defs: Sequence[FeatureDefinition], splits: Sequence[str]):
    """rows of the listed splits, first occurrence of an entity index wins, sorted by entity index"""
    import pandas as pd
    col = f"{entity}_idx"
    table, multi = None, {}
    for k, split in enumerate(splits):
        t, m = _read_entity_split(path, entity, defs, split)
        if k == 0:
            table, multi = t, dict(m)
            continue
        if table is not None and t is not None:
            table = pd.concat([table, t[~t[col].isin(table[col])]])
        for name, (ind, val) in m.items():
            have_i, have_v = multi[name]
            new = ~np.isin(ind, have_i)
            multi[name] = (np.concatenate([have_i, ind[new]], axis=0), np.concatenate([have_v, val[new]], axis=0))
    if table is not None:
        table = table.sort_values(by=col)
    for name, (ind, val) in multi.items():
        order = np.argsort(ind)
        multi[name] = (ind[order], val[order])
    return table, multi


def load_features(path: str, entity: str, defs: Sequence[FeatureDefinition], split: str) -> Dict[str, DiskFeature]:
    """``RecDataset._load_features`` (``data/dataset.py:190-231``)"""
    defs = [FeatureDefinition.from_dict(d) for d in (defs or [])]
    if not defs:
        return {}
    ref_table, _ = _merge_splits(path, entity, defs, SPLITS)
    table, multi = _merge_splits(path, entity, defs, (split, "val") if split == "train" else (split,))
    out = {}
    for d in defs:
        if d.type in TABULAR_TYPES:
            out[d.name] = DiskFeature(d, table[d.name].to_numpy(), indices=table[f"{entity}_idx"].to_numpy(),
                                      reference_values=ref_table[d.name].to_numpy())
        else:
            ind, val = multi[d.name]
            out[d.name] = DiskFeature(d, val, indices=ind)
    return out


# ------------------------------------------------------------------------------------------------ the dataset view
def _cold_start_type(path: str) -> Optional[str]:
    import yaml
    with open(os.path.join(path, "used_config.yaml")) as fh:
        cfg = yaml.safe_load(fh) or {}
    t = (cfg.get("split") or {}).get("cold_start_type")
    return None if t is None else str(t).lower()


def _matrix(lhs, n_users, n_items):
    return sp.coo_matrix((np.ones(len(lhs), dtype=np.int8), (lhs["user_idx"].to_numpy(), lhs["item_idx"].to_numpy())),
                         shape=(n_users, n_items))


class DiskDataset:
    """One split ('train' | 'val' | 'test') of a preprocessed dataset directory."""

    def __init__(self, dataset_path: str, split_set: str = "train", user_feature_definitions=None,
                 item_feature_definitions=None, n_negative_samples: int = 4,
                 negative_sampling_strategy: str = "uniform", use_dataset_negative_sampler: bool = False):
        import pandas as pd
        assert split_set in SPLITS, f"<{split_set}> is not a valid value for split set!"
        self.data_path, self.split_set = dataset_path, split_set
        self.name = f"{os.path.basename(os.path.normpath(dataset_path))}-{split_set}"
        self.is_train_split, self.is_eval_split = split_set == "train", split_set in ("val", "test")
        self.cold_start_type = _cold_start_type(dataset_path)
        self.is_cold_start_user = self.cold_start_type in ("user", "both")
        self.is_cold_start_item = self.cold_start_type in ("item", "both")
        self.is_cold_start_dataset = self.is_cold_start_user or self.is_cold_start_item

        user_idxs = pd.read_csv(os.path.join(dataset_path, "user_idxs.csv"))
        item_idxs = pd.read_csv(os.path.join(dataset_path, "item_idxs.csv"))
        self.n_users, self.n_items = len(user_idxs), len(item_idxs)
        read = lambda s: pd.read_csv(os.path.join(dataset_path, f"listening_history_{s}.csv"))  # noqa: E731
        lhs = read(split_set)
        if self.is_cold_start_dataset:
            self.users_in_split = np.array(sorted(lhs["user_idx"].unique()))
            self.items_in_split = np.array(sorted(lhs["item_idx"].unique()))
        else:
            self.users_in_split = user_idxs["user_idx"].to_numpy()
            self.items_in_split = item_idxs["item_idx"].to_numpy()
        self.n_interactions = len(lhs)
        self.n_users_in_split, self.n_items_in_split = len(self.users_in_split), len(self.items_in_split)
        self.interaction_matrix = _matrix(lhs, self.n_users, self.n_items)
        self.user_sampling_matrix = sp.csr_matrix(self.interaction_matrix)
        train = lhs if self.is_train_split else read("train")
        self.interaction_matrix_train = _matrix(train, self.n_users, self.n_items)
        self.user_sampling_matrix_train = sp.csr_matrix(self.interaction_matrix_train)
        self.item_sampling_matrix_train = sp.csr_matrix(self.interaction_matrix_train.T)
        mask = sp.csr_matrix((self.n_users, self.n_items), dtype=bool)
        if split_set != "train":
            mask = mask + self.user_sampling_matrix_train.astype(bool)
        if split_set == "test":
            mask = mask + sp.csr_matrix(_matrix(read("val"), self.n_users, self.n_items)).astype(bool)
        self.exclude_data = mask[:, self.items_in_split].astype(bool).tocsr()
        self.exclude_data.sort_indices()

        self.user_feature_definitions = [FeatureDefinition.from_dict(d) for d in (user_feature_definitions or [])]
        self.item_feature_definitions = [FeatureDefinition.from_dict(d) for d in (item_feature_definitions or [])]
        self.user_feature_names = [d.name for d in self.user_feature_definitions]
        self.item_feature_names = [d.name for d in self.item_feature_definitions]
        self.user_features = load_features(dataset_path, "user", self.user_feature_definitions, split_set)
        self.item_features = load_features(dataset_path, "item", self.item_feature_definitions, split_set)
        self.features = {"user": self.user_features, "item": self.item_features}
        self.n_negative_samples = n_negative_samples
        self.negative_sampling_strategy = negative_sampling_strategy
        self.use_dataset_negative_sampler = use_dataset_negative_sampler

    def __len__(self):
        return self.interaction_matrix.nnz if self.is_train_split else self.n_users_in_split


class DiskCorpus:
    """the three split views of one directory (``dataset(split)`` like ``synthetic.SynCorpus``)"""

    def __init__(self, dataset_path: str, user_feature_definitions: List = None, item_feature_definitions: List = None,
                 **train_kw):
        self.path, self.ufd, self.ifd, self.train_kw = dataset_path, user_feature_definitions, item_feature_definitions, train_kw
        self._views = {}

    def dataset(self, split: str) -> DiskDataset:
        if split not in self._views:
            self._views[split] = DiskDataset(self.path, split, self.ufd, self.ifd, **self.train_kw)
        return self._views[split]
