"""Model / learning / eval configuration with the reference's keys (``data/module_config_classes.py:19-25,45-127``,
``data/config_classes.py:172-198``).  Plain dataclasses with a small ``from_dict`` -- no third-party config library.
"""
from __future__ import annotations

import dataclasses
import enum
from dataclasses import dataclass, field
from typing import List, Optional, Set, Union


class EmbeddingRegularizationType(enum.Enum):
    NoRegularization = "no_regularization"
    PairwiseSingle = "pairwise_single"
    CentralModality = "central_modality"


class MissingField(KeyError):
    pass


def _build(cls, d: dict):
    known = {f.name: f for f in dataclasses.fields(cls)}
    d = {k: v for k, v in d.items() if k in known}  # unknown keys are ignored, as the reference's from_dict does
    for f in known.values():
        if f.name not in d and f.default is dataclasses.MISSING and f.default_factory is dataclasses.MISSING:
            raise MissingField(f'{cls.__name__}: required key "{f.name}" is missing')
    return cls(**d)


@dataclass
class FeatureModuleConfig:
    feature_name: str
    embedding_dim: int
    pre_embedding_layers: Optional[List[int]] = None
    post_embedding_layers: Optional[List[int]] = None
    activation_fn: str = "relu"

    @classmethod
    def from_dict(cls, d: dict):
        return _build(cls, dict(d))


@dataclass
class SingleBranchFeatureConfig:
    feature_name: str
    feature_hidden_layers: Optional[List[int]] = None

    @classmethod
    def from_dict(cls, d: dict):
        return _build(cls, dict(d))


@dataclass
class SingleBranchNetEntityConfig:
    features: List[SingleBranchFeatureConfig]
    single_branch_hidden_layers: List[int]
    preference_hidden_layers: List[int]  # required by the reference, unused by it
    common_modality_dim: int
    activation_fn: str = "relu"
    train_modalities: Optional[Set[str]] = None
    eval_modalities: Optional[Set[str]] = None
    sampling_seed: int = 42
    single_branch_input_dropout: Optional[float] = None
    aggregation_fn: str = "mean"
    normalize_single_branch_input: bool = False
    embedding_regularization_type: EmbeddingRegularizationType = EmbeddingRegularizationType.NoRegularization
    central_modality: Optional[str] = None
    regularization_temperature: float = 1.0
    regularization_weight: float = 1.0
    apply_output_activation: bool = False
    apply_batch_normalization: bool = True
    apply_batch_norm_every: int = 0

    @classmethod
    def from_dict(cls, d: dict):
        d = dict(d)
        if "features" in d:
            d["features"] = [f if isinstance(f, SingleBranchFeatureConfig) else SingleBranchFeatureConfig.from_dict(f)
                             for f in d["features"]]
        for key in ("train_modalities", "eval_modalities"):
            if d.get(key) is not None:
                d[key] = set(d[key])
        if "embedding_regularization_type" in d and not isinstance(d["embedding_regularization_type"],
                                                                   EmbeddingRegularizationType):
            d["embedding_regularization_type"] = EmbeddingRegularizationType(d["embedding_regularization_type"])
        return _build(cls, d)


@dataclass
class SingleBranchNetConfig:
    user: Union[SingleBranchNetEntityConfig, FeatureModuleConfig]
    item: Union[SingleBranchNetEntityConfig, FeatureModuleConfig]
    shared_common_dim: int

    @staticmethod
    def _entity(d):
        """plain FeatureModuleConfig first, single-branch entity when a required key is missing
        (``data/module_config_classes.py:114-119``)"""
        if not isinstance(d, dict):
            return d
        try:
            return FeatureModuleConfig.from_dict(d)
        except (MissingField, TypeError):
            return SingleBranchNetEntityConfig.from_dict(d)

    @classmethod
    def from_dict(cls, d: dict):
        d = dict(d)
        for k in ("user", "item", "shared_common_dim"):
            if k not in d:
                raise MissingField(f'SingleBranchNetConfig: required key "{k}" is missing')
        return cls(user=cls._entity(d["user"]), item=cls._entity(d["item"]), shared_common_dim=d["shared_common_dim"])

    @property
    def is_user_sb_module(self) -> bool:
        return isinstance(self.user, SingleBranchNetEntityConfig)

    @property
    def is_item_sb_module(self) -> bool:
        return isinstance(self.item, SingleBranchNetEntityConfig)


@dataclass
class LearningConfig:
    n_epochs: int = 50
    max_batches_per_epoch: Optional[int] = None
    lr: float = 1e-3
    wd: float = 0.0
    optimizer: str = "adam"          # adam | adamw | adagrad  (train/trainer.py:62-66)
    optimizing_metric: str = "ndcg@10"
    rec_loss: str = "bce"            # bce | bpr | sampled_softmax
    loss_aggregator: str = "mean"    # mean | sum
    max_patience: int = 2 ** 62

    @classmethod
    def from_dict(cls, d: dict):
        names = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in d.items() if k in names})


@dataclass
class EvalConfig:
    top_k: List[int] = field(default_factory=lambda: [1, 3, 5, 10, 20, 50, 100])
    metrics: List[str] = field(default_factory=lambda: ["ndcg", "precision", "recall", "f_score", "hitrate",
                                                        "coverage"])
    calculate_std: bool = True
    calculate_group_metrics: bool = False
    user_group_features: Optional[List[str]] = None

    @classmethod
    def from_dict(cls, d: dict):
        names = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in d.items() if k in names})
