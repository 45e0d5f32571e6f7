"""DropoutNet on the sibrar_b200 kernels (reference ``DropoutNetEntity`` / ``DropoutNet``,
``algorithms/sgd_alg.py:1617-1761``, config classes ``data/module_config_classes.py:10-43``).

Per entity:  ``net(cat([content_1(idx), ..., content_m(idx), pref_net(preferences)]))`` with ``preferences`` = the
entity's row of the train interaction matrix, or a zero vector for rows whose sampled training strategy is
``NoPreference`` (the cold-start simulation of the paper; in evaluation every row keeps its preferences).

Same constructor ``(config, dataset)``, ``build_from_conf``, ``forward / predict / get_{user,item}_representations /
combine_user_item_representations / get_and_reset_other_loss / sample_training_strategy`` and ``state_dict()`` keys as the
reference (``{user,item}_net.{pref_net, net}.layers.linear_i.*``, ``{user,item}_net.cont_modules.j.*``).  The strategy draws
use the reference's generator (``np.random.default_rng(sampling_seed).choice``, users first, then items, one draw per
leading index), so a run with the same seed drops the same preferences.

How it runs: every tower is a function of the ENTITY ROW alone, so each step builds one table per tower over all rows
(``PlainEntity``: tcgen05 GEMMs over the bit-packed / CSR interaction matrix resp. the feature table) and the batch gathers
rows of them straight into the column blocks of the concatenated input; ``NoPreference`` rows gather the extra all-zero
row the preference table carries (= the bias path of ``pref_net``, as in the reference).  The common net is a ``Chain``
(bf16 operands, fp32 accumulation); the backward runs the chain, then every tower's sorted-run gather backward and
table-level backward.  ``model(u, i)`` in training mode is differentiable (hand-written backward behind an
``autograd.Function``), so the reference's loop -- torch loss on the logits, ``backward()``, ``torch.optim`` -- runs
unchanged.
"""
from __future__ import annotations

import enum
import os
import weakref
from dataclasses import dataclass
from typing import Dict, List

import numpy as np
import scipy.sparse as sp
import torch
from torch import nn

from . import ops
from .config import FeatureModuleConfig, MissingField, _build
from .sbnet import Chain, FeatureEmbedding, PlainEntity, PolyLinear, _Runtime, build_stages

F32 = torch.float32


class DropoutNetSamplingStrategy(enum.Enum):
    """data/module_config_classes.py:10-16 (``auto()`` values: 1, 2)"""
    Normal = 1
    NoPreference = 2

    @classmethod
    def list(cls):
        return [c.value for c in cls]


@dataclass
class DropoutNetEntityConfig:
    features: List[FeatureModuleConfig]
    preference_layers: List[int]      # the number of items / users is prepended
    common_hidden_layers: List[int]   # feature + preference width is prepended, shared_common_dim appended
    activation_fn: str = "relu"

    @classmethod
    def from_dict(cls, d: dict):
        d = dict(d)
        if "features" in d:
            d["features"] = [f if isinstance(f, FeatureModuleConfig) else FeatureModuleConfig.from_dict(f)
                             for f in d["features"]]
        return _build(cls, d)


@dataclass
class DropoutNetConfig:
    user: DropoutNetEntityConfig
    item: DropoutNetEntityConfig
    shared_common_dim: int
    sampling_seed: int = 42

    @classmethod
    def from_dict(cls, d: dict):
        d = dict(d)
        for k in ("user", "item", "shared_common_dim"):
            if k not in d:
                raise MissingField(f'DropoutNetConfig: required key "{k}" is missing')
        for k in ("user", "item"):
            if isinstance(d[k], dict):
                d[k] = DropoutNetEntityConfig.from_dict(d[k])
        return _build(cls, d)

    def to_dict(self):
        import dataclasses
        return dataclasses.asdict(self)


class DropoutNetEntity(nn.Module):
    """reference ``DropoutNetEntity`` (sgd_alg.py:1617-1655): parameter holders under the reference's names + the engines
    that run them (unregistered: the parameters are registered once, through the holders)"""

    def __init__(self, entity_name: str, preference_dim: int, features: Dict, entity_config: DropoutNetEntityConfig,
                 shared_common_dim: int, preferences: sp.spmatrix):
        super().__init__()
        from .synthetic import SynFeature
        self.entity_name, self.entity_config, self.shared_common_dim = entity_name, entity_config, shared_common_dim
        layers = list(entity_config.preference_layers)
        m = sp.csr_matrix(preferences, dtype=np.float32)
        if m.shape[1] != preference_dim:
            raise ValueError(f"{entity_name}: preference matrix has {m.shape[1]} columns, expected {preference_dim}")
        self.n_entities = int(m.shape[0])
        # the rows of the train interaction matrix + ONE all-zero row (index n_entities): what a NoPreference row reads
        table = sp.vstack([m, sp.csr_matrix((1, preference_dim), dtype=np.float32)]).tocsr()
        pref_feature = SynFeature(f"{entity_name}_preferences", "vector", table)
        # PolyLinear([preference_dim] + preference_layers) with its defaults (ReLU between the layers AND on the output,
        # modules/polylinear.py:18-19) is what a FeatureEmbedding of a vector feature builds
        pref_fe = FeatureEmbedding(pref_feature, embedding_dim=layers[-1], pre_embedding_layers=layers[:-1],
                                   activation_fn="relu")
        self.pref_net = pref_fe.pre_embedding_layers
        self.pref_dim = layers[-1]
        self.cont_modules = nn.ModuleList(FeatureEmbedding.build_from_conf(f, features[f.feature_name])
                                          for f in entity_config.features)
        self.cont_dim = sum(int(mod.output_dim) for mod in self.cont_modules)
        self._net_shape = [self.pref_dim + self.cont_dim] + list(entity_config.common_hidden_layers) + [shared_common_dim]
        self.net = PolyLinear(self._net_shape, activation_fn=entity_config.activation_fn)  # (output_fn: default ReLU)
        cont = [PlainEntity(features[f.feature_name], f, self.n_entities, fe=mod)
                for f, mod in zip(entity_config.features, self.cont_modules)]
        pref = PlainEntity(pref_feature, FeatureModuleConfig(feature_name=pref_feature.feature_definition.name,
                                                             embedding_dim=layers[-1], pre_embedding_layers=layers[:-1]),
                           self.n_entities + 1, fe=pref_fe)
        object.__setattr__(self, "_cont", cont)
        object.__setattr__(self, "_pref", pref)
        object.__setattr__(self, "_chain", None)
        object.__setattr__(self, "_saved", None)

    def engines(self):
        return list(self._cont) + [self._pref]

    def chain(self) -> Chain:
        if self._chain is None:
            object.__setattr__(self, "_chain", Chain(build_stages(self.net)))
        return self._chain

    def forward(self, *a, **k):
        raise RuntimeError("DropoutNetEntity is a parameter container; arithmetic runs in the sibrar_b200 kernels")

    # ---- the kernels
    def embed(self, idx: torch.Tensor, pref_idx: torch.Tensor, training: bool, rt: _Runtime) -> torch.Tensor:
        """idx / pref_idx: int64 [rows] (pref_idx = idx, or n_entities for a row without preferences) -> fp32 [rows, D]"""
        rows = idx.numel()
        width = self.cont_dim + self.pref_dim
        X = torch.empty((rows, width), dtype=F32, device=idx.device)
        col = 0
        for eng in self._cont:
            eng.embed(idx, training, out=X[:, col:col + eng.output_dim])
            col += eng.output_dim
        self._pref.embed(pref_idx, training, out=X[:, col:])
        x16 = ops.cast_bf16(X)
        out = self.chain().forward(x16, rows, training, rt.arena, keep_for_backward=training)
        object.__setattr__(self, "_saved", (rows, width))
        return out

    def backward(self, d_out: torch.Tensor, grads, rt: _Runtime):
        rows, width = self._saved
        dX = self.chain().backward(d_out.contiguous(), grads, need_dx=True, arena=rt.arena)
        col = 0
        for eng in self._cont:
            eng.backward(dX[:, col:col + eng.output_dim], grads)
            col += eng.output_dim
        self._pref.backward(dX[:, col:width], grads)


class DropoutNet(nn.Module):
    """reference ``DropoutNet`` (sgd_alg.py:1658-1761)"""

    def __init__(self, config: DropoutNetConfig, dataset):
        super().__init__()
        self.config, self.dataset = config, dataset
        u_pref = dataset.user_sampling_matrix_train
        i_pref = getattr(dataset, "item_sampling_matrix_train", None)
        if i_pref is None:
            i_pref = sp.csr_matrix(u_pref).T.tocsr()
        self.user_net = DropoutNetEntity("user", dataset.n_items, dataset.user_features, config.user,
                                         config.shared_common_dim, u_pref)
        self.item_net = DropoutNetEntity("item", dataset.n_users, dataset.item_features, config.item,
                                         config.shared_common_dim, i_pref)
        self._rng = np.random.default_rng(config.sampling_seed)
        self.name = "DropoutNet"
        self._runtime = None
        ref = weakref.ref(self)
        for ent in (self.user_net, self.item_net):
            for eng in ent.engines():
                eng._owner = lambda r=ref: r()._rt()

    @staticmethod
    def build_from_conf(conf: dict, dataset):
        return DropoutNet(DropoutNetConfig.from_dict(conf), dataset)

    # ---- runtime
    @property
    def device(self):
        return next(iter(self.parameters())).device

    def _rt(self) -> _Runtime:
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("sibrar_b200 models run on a CUDA device (sm_100a) only -- there is no CPU fallback")
        if self._runtime is None or self._runtime.device != dev:
            self._runtime = _Runtime(dev)
        return self._runtime

    def refresh_shadows(self):
        for ent in (self.user_net, self.item_net):
            for st in ent.chain().stages:
                st.refresh(False)
            for eng in ent.engines():
                for chain in eng.chains():
                    for st in chain.stages:
                        st.refresh(False)

    def check_errors(self):
        if self._runtime is not None and int(self._runtime.err_flag.item()) != 0:
            self._runtime.err_flag.zero_()
            raise KeyError("an entity index without a feature row was requested")

    # ---- reference API
    def sample_training_strategy(self, n_samples):
        if self.training:
            return self._rng.choice(DropoutNetSamplingStrategy.list(), size=n_samples, replace=True)
        return np.full(n_samples, fill_value=DropoutNetSamplingStrategy.Normal.value)  # evaluation: all information

    def _pref_index(self, idx: torch.Tensor, n_entities: int, strategy=None) -> torch.Tensor:
        """entity index of every row, or ``n_entities`` (the zero row) where the LEADING index drew NoPreference"""
        if strategy is None:
            strategy = self.sample_training_strategy(len(idx))
        strategy = np.asarray(strategy)
        if np.all(strategy == DropoutNetSamplingStrategy.Normal.value):
            return idx
        keep = torch.from_numpy(strategy == DropoutNetSamplingStrategy.Normal.value).to(idx.device)
        keep = keep.view(-1, *([1] * (idx.dim() - 1))).expand_as(idx)
        return torch.where(keep, idx, torch.full_like(idx, n_entities))

    def _represent(self, ent: DropoutNetEntity, idx: torch.Tensor, strategy=None):
        pref_idx = self._pref_index(idx, ent.n_entities, strategy)
        out = ent.embed(idx.reshape(-1).contiguous(), pref_idx.reshape(-1).contiguous(), self.training, self._rt())
        return out.view(*idx.shape, -1)

    @torch.no_grad()
    def get_user_representations(self, u_idxs: torch.Tensor, strategy=None):
        self._rt().arena.reset()
        return self._represent(self.user_net, u_idxs, strategy)

    @torch.no_grad()
    def get_item_representations(self, i_idxs: torch.Tensor, strategy=None):
        self._rt().arena.reset()
        return self._represent(self.item_net, i_idxs, strategy)

    @torch.no_grad()
    def combine_user_item_representations(self, u_repr, i_repr):
        from .autograd import _combine_nograd
        return _combine_nograd(u_repr, i_repr)

    def forward(self, u_idxs, i_idxs, strategies=None):
        """``strategies`` (optional, parity tests): (user strategies [B], item strategies [B]) instead of the draws"""
        if self.training and torch.is_grad_enabled():
            anchor = next(iter(self.parameters()))
            return _DNTrainForward.apply(anchor, self, u_idxs, i_idxs, strategies)
        return self._forward_nograd(u_idxs, i_idxs, strategies)[0]

    @torch.no_grad()
    def predict(self, u_idxs, i_idxs):
        self.eval()
        return self(u_idxs, i_idxs)

    def get_and_reset_other_loss(self) -> Dict:
        return {"reg_loss": torch.zeros(1, device=self.device)}  # algorithms/base_classes.py:136-145

    def save_model_to_path(self, path: str):
        torch.save(self.state_dict(), os.path.join(path, "model.pth"))
        print("Model Saved")

    def load_model_from_path(self, path: str):
        self.load_state_dict(torch.load(os.path.join(path, "model.pth"), map_location=self.device))
        print("Model Loaded")

    # ---- the step on the kernels
    def _forward_nograd(self, u_idxs, i_idxs, strategies=None):
        rt = self._rt()
        ops.tick(rt.step_dev)
        rt.arena.reset()
        su, si = strategies if strategies is not None else (None, None)
        Eu = self._represent(self.user_net, u_idxs, su)          # (users first: the order of the reference's draws)
        Ei = self._represent(self.item_net, i_idxs, si)
        B, D = Eu.shape
        if i_idxs.dim() == 1:
            from .autograd import _combine_nograd
            return _combine_nograd(Eu, Ei), Eu, Ei
        n = i_idxs.shape[1]
        logits = torch.empty((B, n), dtype=F32, device=Eu.device)
        ops.score_loss(Eu.contiguous(), Ei.contiguous(), B, n, 1, 1, D, 0, 0, "bce", 0, 0.0, logits, None)
        return logits, Eu, Ei


class _DNTrainForward(torch.autograd.Function):
    """logits = model(u, i) in training mode; backward = the hand-written backward kernels (accumulates into
    ``param.grad``)"""

    @staticmethod
    def forward(ctx, anchor, model, u_idxs, i_idxs, strategies):
        logits, Eu, Ei = model._forward_nograd(u_idxs, i_idxs, strategies)
        ctx.model, ctx.saved = model, (i_idxs.shape, Eu, Ei)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        (B, n), Eu, Ei = ctx.saved
        D = Eu.shape[1]
        rt = model._rt()
        grads = {}
        for p in model.parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads[id(p)] = p.grad
        dl = dlogits.contiguous().to(F32)
        su, si = Eu.contiguous().view(B, 1, D), Ei.contiguous().view(B * n, 1, D)
        dEu, dEi = torch.empty_like(su), torch.empty_like(si)
        ops.score_bwd(su, si, B, n, 1, 1, D, 0, 0, dl, dEu, dEi)
        # the item entity ran last in the forward: its chain / engines still hold that pass; the user entity's state is
        # its own (separate modules), so the order of the two backward passes is free
        model.item_net.backward(dEi.view(B * n, D), grads, rt)
        model.user_net.backward(dEu.view(B, D), grads, rt)
        return None, None, None, None, None
