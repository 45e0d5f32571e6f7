"""Sibling models of SingleBranchNet on the same kernels: the matrix-factorisation
family of the reference -- ``SGDMatrixFactorization`` (``algorithms/sgd_alg.py:126-200``),
``ItemFeatureMatrixFactorization`` (``:1399-1505``) and ``UserFeatureMatrixFactorization`` (``:1508-1614``; MF + a
``FeatureEmbedding`` content tower + symmetric InfoNCE between the profile and the content embedding, "CLCRec").

Same constructor arguments, ``build_from_conf`` keys, ``forward / predict / get_{user,item}_representations /
combine_user_item_representations / get_and_reset_other_loss`` and ``state_dict()`` keys as the reference classes
(``user_embeddings.weight``, ``item_embeddings.weight``, ``user_bias.weight``, ``item_bias.weight``, ``global_bias``,
``embedding_net.{pre_embedding_layers.layers.linear_i.*, embedding_layer.weight}``).  Every arithmetic step runs in the
sibrar_b200 kernels: the embedding gathers and their sorted-run backward (``PlainEntity`` engines), the content tower
as a per-step table of all feature rows, the k-slot score kernel (profile / content = two slots of one row; their mean
is ``aggregate_for_rec``), ``sbr_infonce`` on the same two slots, the bias kernels ``sbr_logit_bias_fwd/bwd``.

``model(u, i)`` in training mode is differentiable (the reference's loop: torch loss on the logits, ``loss.backward()``,
``torch.optim``); the InfoNCE gradient enters with coefficient 1 (``total = rec_loss + reg_loss``,
``train/trainer.py:215``).  Evaluation goes through ``FullEvaluator`` via ``eval_factors`` (biases folded into two extra
factor columns, so the fused score + mask + top-k kernel sees plain dot products).
"""
from __future__ import annotations

import os
import weakref
from typing import Dict

import numpy as np
import torch
from torch import nn

from . import ops
from .config import FeatureModuleConfig
from .sbnet import FeatureEmbedding, PlainEntity, _Runtime

F32 = torch.float32


def _id_feature(name, n):
    from .synthetic import SynFeature
    return SynFeature(name, "categorical", np.arange(n))


class SGDMatrixFactorization(nn.Module):
    """reference ``SGDMatrixFactorization`` (sgd_alg.py:126-200)"""
    content_side = None  # 'item' | 'user' in the feature-extended subclasses

    def __init__(self, n_users: int, n_items: int, embedding_dim: int = 100, use_user_bias: bool = False,
                 use_item_bias: bool = False, use_global_bias: bool = False):
        super().__init__()
        self.n_users, self.n_items, self.embedding_dim = n_users, n_items, embedding_dim
        self.use_user_bias, self.use_item_bias, self.use_global_bias = use_user_bias, use_item_bias, use_global_bias
        # parameter holders under the reference's names; nn.Embedding init = general_weight_init (train/utils.py:11-13)
        fe_u = FeatureEmbedding(_id_feature("user_embedding", n_users), embedding_dim=embedding_dim)
        fe_i = FeatureEmbedding(_id_feature("item_embedding", n_items), embedding_dim=embedding_dim)
        self.user_embeddings, self.item_embeddings = fe_u.embedding_layer, fe_i.embedding_layer
        if use_user_bias:
            self.user_bias = nn.Embedding(n_users, 1)
            nn.init.normal_(self.user_bias.weight, std=.1)
        if use_item_bias:
            self.item_bias = nn.Embedding(n_items, 1)
            nn.init.normal_(self.item_bias.weight, std=.1)
        if use_global_bias:
            self.global_bias = nn.Parameter(torch.zeros(1), requires_grad=True)
        self.name = "SGDMatrixFactorization"
        self._runtime = None
        ref = weakref.ref(self)
        cfg = FeatureModuleConfig(feature_name="id", embedding_dim=embedding_dim)
        engines = {"user": PlainEntity(fe_u._feature, cfg, n_users, fe=fe_u),
                   "item": PlainEntity(fe_i._feature, cfg, n_items, fe=fe_i)}
        object.__setattr__(self, "_engines", engines)  # (not registered: the parameters are registered above)
        for e in engines.values():
            e._owner = lambda r=ref: r()._rt()

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
        for e in self._engines.values():
            for chain in e.chains():
                for st in chain.stages:
                    st.refresh(False)

    def check_errors(self):
        if self._runtime is not None and int(self._runtime.err_flag.item()) != 0:
            self._runtime.err_flag.zero_()
            raise KeyError("an entity index without a feature row was requested")

    def _bias(self, which):
        if which == "user":
            return self.user_bias.weight.detach().view(-1) if self.use_user_bias else None
        if which == "item":
            return self.item_bias.weight.detach().view(-1) if self.use_item_bias else None
        return self.global_bias.detach() if self.use_global_bias else None

    # ---- slots: [profile] or [profile | content] per entity row
    def _slots(self, side):
        return 2 if self.content_side == side else 1

    def _embed_side(self, side, idx, training):
        """fp32 [numel, k, D]: slot 0 = the MF embedding, slot 1 = the content tower (feature-extended side only)"""
        k, D = self._slots(side), self.embedding_dim
        flat = idx.reshape(-1)
        E = torch.empty((flat.numel(), k, D), dtype=F32, device=flat.device)
        wide = E.view(flat.numel(), k * D)
        self._engines[side].embed(flat, training, out=wide[:, :D])
        if k == 2:
            self._engines["content"].embed(flat, training, out=wide[:, D:])
        return E

    # ---- reference API
    def get_user_representations(self, u_idxs):
        with torch.no_grad():
            self._rt().arena.reset()
            E = self._embed_side("user", u_idxs, self.training)
        out = tuple(E[:, s].reshape(*u_idxs.shape, -1) for s in range(E.shape[1]))
        if self.use_user_bias:
            out += (self.user_bias.weight.detach()[u_idxs] if self.content_side != "user"
                    else self.user_bias.weight.detach()[u_idxs].squeeze(),)
        return out if len(out) > 1 else out[0]

    def get_item_representations(self, i_idxs):
        with torch.no_grad():
            self._rt().arena.reset()
            E = self._embed_side("item", i_idxs, self.training)
        out = tuple(E[:, s].reshape(*i_idxs.shape, -1) for s in range(E.shape[1]))
        if self.use_item_bias:
            out += (self.item_bias.weight.detach()[i_idxs].squeeze(-1),)
        return out if len(out) > 1 else out[0]

    def _unpack(self, repr_, side, use_bias):
        """-> (embedding used for the score, bias or None) like the reference's combine methods"""
        if not isinstance(repr_, tuple):
            return repr_, None
        bias = repr_[-1] if use_bias else None
        embs = repr_[:-1] if use_bias else repr_
        if self.content_side == side and len(embs) >= 2:
            emb = torch.stack([embs[0], embs[1]], dim=-2) if self.aggregate_for_rec else embs[0]
        else:
            emb = embs[0]
        return emb, bias

    def _factors(self, emb, bias, side):
        """[rows, k, D] slots (+ bias [rows]) -> fp32 [rows, D + 3] whose dot product with the other side's factors is
        the reference score (sgd_alg.py:187-195): user = [mean_k u, u_bias, 1, global], item = [mean_k i, 1, i_bias, 1]"""
        D = self.embedding_dim
        E = emb if emb.dim() == 3 else emb.reshape(-1, 1, D)
        F_ = torch.zeros((E.shape[0], D + 3), dtype=F32, device=E.device)
        if E.shape[1] == 1:
            F_[:, :D] = E[:, 0]
        else:  # (sbr_aggregate writes a dense [rows, D] block)
            mean = torch.empty((E.shape[0], D), dtype=F32, device=E.device)
            ops.aggregate(E.contiguous(), E.shape[0], E.shape[1], D, 0, out_f32=mean)
            F_[:, :D] = mean
        if side == "user":
            F_[:, D + 1] = 1.0
            if bias is not None:
                F_[:, D] = bias.reshape(-1)
            if self.use_global_bias:
                F_[:, D + 2] = self.global_bias.detach()
        else:
            F_[:, D] = 1.0
            F_[:, D + 2] = 1.0
            if bias is not None:
                F_[:, D + 1] = bias.reshape(-1)
        return F_

    @torch.no_grad()
    def combine_user_item_representations(self, u_repr, i_repr):
        """scores from representations (no autograd graph; gradients flow through ``forward``): [B, n] for per-interaction
        item representations [B, n, D], [B, I] for a catalogue [I, D]"""
        from .autograd import _combine_nograd
        u_emb, u_bias = self._unpack(u_repr, "user", self.use_user_bias)
        i_emb, i_bias = self._unpack(i_repr, "item", self.use_item_bias)
        D = self.embedding_dim
        stacked_i = self.content_side == "item" and self.aggregate_for_rec and isinstance(i_repr, tuple)
        stacked_u = self.content_side == "user" and self.aggregate_for_rec and isinstance(u_repr, tuple)
        fu = self._factors(u_emb if stacked_u else u_emb.reshape(-1, 1, D), u_bias, "user")
        B = fu.shape[0]
        lead = i_emb.shape[:-2] if stacked_i else i_emb.shape[:-1]
        fi = self._factors(i_emb.reshape(-1, 2, D) if stacked_i else i_emb.reshape(-1, 1, D), i_bias, "item")
        if len(lead) == 1:  # catalogue
            return _combine_nograd(fu, fi)
        n = lead[1]
        logits = torch.empty((B, n), dtype=F32, device=fu.device)
        ops.score_loss(fu, fi, B, n, 1, 1, D + 3, 0, 0, "bce", 0, 0.0, logits, None)
        return logits

    def forward(self, u_idxs, i_idxs):
        if self.training and torch.is_grad_enabled():
            anchor = next(iter(self.parameters()))
            return _MFTrainForward.apply(anchor, self, u_idxs, i_idxs)
        return self._forward_nograd(u_idxs, i_idxs)[0]

    @torch.no_grad()
    def predict(self, u_idxs, i_idxs):
        self.eval()
        return self(u_idxs, i_idxs)

    def get_and_reset_other_loss(self) -> Dict:
        return {"reg_loss": torch.zeros(1, device=self.device)}

    def save_model_to_path(self, path: str):
        torch.save(self.state_dict(), os.path.join(path, "model.pth"))
        print("Model Saved")

    def load_model_from_path(self, path: str):
        self.load_state_dict(torch.load(os.path.join(path, "model.pth"), map_location=self.device))
        print("Model Loaded")

    @staticmethod
    def build_from_conf(conf: dict, dataset):
        return SGDMatrixFactorization(dataset.n_users, dataset.n_items, conf["embedding_dim"], conf["use_user_bias"],
                                      conf["use_item_bias"], conf["use_global_bias"])

    # ---- the step on the kernels
    def _score_operands(self, Eu, Ei):
        """the slots that enter the score: both (their mean) with aggregate_for_rec, else slot 0 (a contiguous copy)"""
        agg = getattr(self, "aggregate_for_rec", False)
        su = Eu if (Eu.shape[1] == 1 or agg) else Eu[:, :1].contiguous()
        si = Ei if (Ei.shape[1] == 1 or agg) else Ei[:, :1].contiguous()
        return su, si

    def _forward_nograd(self, u_idxs, i_idxs):
        rt = self._rt()
        B, n = i_idxs.shape
        D = self.embedding_dim
        ops.tick(rt.step_dev)
        rt.arena.reset()
        Eu = self._embed_side("user", u_idxs, self.training)
        Ei = self._embed_side("item", i_idxs, self.training)
        su, si = self._score_operands(Eu, Ei)
        logits = torch.empty((B, n), dtype=F32, device=Eu.device)
        ops.score_loss(su, si, B, n, su.shape[1], si.shape[1], D, 0, 0, "bce", 0, 0.0, logits, None)
        if self.use_user_bias or self.use_item_bias or self.use_global_bias:
            ops.logit_bias_fwd(logits, u_idxs.contiguous(), i_idxs.contiguous(), self._bias("user"), self._bias("item"),
                               self._bias("global"))
        return logits, Eu, Ei, su, si

    def _reg(self, Eu, Ei, B, n):
        """-> (loss tensor [1] or None, dEu_reg, dEi_reg): subclasses with a content tower"""
        return None, None, None

    # ---- evaluation: factor matrices whose plain dot product is the model's score
    @torch.no_grad()
    def eval_factors(self, users, items):
        """fp32 [U, D + 3], [I, D + 3] factor matrices (see ``_factors``): ``FullEvaluator``'s fused score + mask + top-k
        kernel computes the reference's scores as plain dot products"""
        was_training = self.training
        self.eval()
        self._rt().arena.reset()
        Eu = self._embed_side("user", users, False)
        Ei = self._embed_side("item", items, False)
        if was_training:
            self.train()
        su, si = self._score_operands(Eu, Ei)
        ub = self.user_bias.weight.detach()[users] if self.use_user_bias else None
        ib = self.item_bias.weight.detach()[items] if self.use_item_bias else None
        return self._factors(su, ub, "user"), self._factors(si, ib, "item")


class _MFTrainForward(torch.autograd.Function):
    """logits = model(u, i) in training mode; backward = the hand-written backward kernels (accumulates into
    ``param.grad``)"""

    @staticmethod
    def forward(ctx, anchor, model, u_idxs, i_idxs):
        B, n = i_idxs.shape
        logits, Eu, Ei, su, si = model._forward_nograd(u_idxs, i_idxs)
        loss, dEu_reg, dEi_reg = model._reg(Eu, Ei, B, n)
        model.emb_loss = loss if loss is not None else 0.
        ctx.model, ctx.saved = model, (u_idxs, i_idxs, Eu, Ei, su, si, dEu_reg, dEi_reg)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        u_idxs, i_idxs, Eu, Ei, su, si, dEu_reg, dEi_reg = ctx.saved
        B, n = i_idxs.shape
        D = model.embedding_dim
        grads = {}
        for p in model.parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads[id(p)] = p.grad
        dl = dlogits.contiguous().to(F32)
        dsu, dsi = torch.empty_like(su), torch.empty_like(si)
        ops.score_bwd(su, si, B, n, su.shape[1], si.shape[1], D, 0, 0, dl, dsu, dsi)

        def full(dscore, E, dreg):
            """gradient w.r.t. every slot of E: score gradient (slot 0 only unless both slots were scored) + InfoNCE"""
            if dscore.shape == E.shape:
                d = dscore
            else:
                d = torch.zeros_like(E)
                d[:, :1] = dscore
            if dreg is not None:
                d = d + dreg
            return d
        dEu, dEi = full(dsu, Eu, dEu_reg), full(dsi, Ei, dEi_reg)
        eng = model._engines
        for side, dE, idx in (("item", dEi, i_idxs), ("user", dEu, u_idxs)):
            k = dE.shape[1]
            wide = dE.view(dE.shape[0], k * D)
            eng[side]._ctx = (idx.reshape(-1).contiguous(),)
            eng[side].backward(wide[:, :D], grads)
            if k == 2:
                eng["content"]._ctx = (idx.reshape(-1).contiguous(),)
                eng["content"].backward(wide[:, D:], grads)
        if model.use_user_bias or model.use_item_bias or model.use_global_bias:
            ops.logit_bias_bwd(dl, u_idxs.contiguous(), i_idxs.contiguous(),
                               grads[id(model.user_bias.weight)].view(-1) if model.use_user_bias else None,
                               grads[id(model.item_bias.weight)].view(-1) if model.use_item_bias else None,
                               grads[id(model.global_bias)] if model.use_global_bias else None)
        return None, None, None, None


class _FeatureMF(SGDMatrixFactorization):
    """shared body of Item/UserFeatureMatrixFactorization"""

    def __init__(self, dataset, feature_name: str, aggregate_for_rec: bool = False, lambda_content: float = 0.0001,
                 temperature: float = 0.1, embedding_loss_aggregator: str = "mean", intermediate_layers=None,
                 embedding_dim: int = 100, use_user_bias: bool = False, use_item_bias: bool = False,
                 use_global_bias: bool = False):
        super().__init__(dataset.n_users, dataset.n_items, embedding_dim, use_user_bias, use_item_bias, use_global_bias)
        if embedding_loss_aggregator not in ("mean", "sum"):
            raise ValueError(f'{embedding_loss_aggregator} is not a valid value for reduction')  # F.cross_entropy
        self.dataset, self.feature_name = dataset, feature_name
        self.aggregate_for_rec, self.lambda_content = aggregate_for_rec, lambda_content
        self.temperature, self.embedding_loss_aggregator = temperature, embedding_loss_aggregator
        feats = dataset.item_features if self.content_side == "item" else dataset.user_features
        n = dataset.n_items if self.content_side == "item" else dataset.n_users
        self.embedding_net = FeatureEmbedding(feature=feats[feature_name], pre_embedding_layers=intermediate_layers,
                                              embedding_dim=embedding_dim)
        if self.embedding_net.output_dim != embedding_dim:
            raise ValueError(f'the content tower of "{feature_name}" produces {self.embedding_net.output_dim}-d vectors, '
                             f'embedding_dim is {embedding_dim}')
        eng = PlainEntity(feats[feature_name], FeatureModuleConfig(feature_name=feature_name, embedding_dim=embedding_dim),
                          n, fe=self.embedding_net)
        ref = weakref.ref(self)
        eng._owner = lambda r=ref: r()._rt()
        self._engines["content"] = eng
        self.emb_loss = 0.

    def _reg(self, Eu, Ei, B, n):
        """symmetric InfoNCE between the profile and the content slot (train/regularization_losses.py:14-43): item side
        contrasts the n items of an interaction (G = B groups of n); user side is called with [B, 1, D] operands in the
        reference (sgd_alg.py:1563-1564), i.e. B groups of ONE row -- a constant zero, kept as it is"""
        if self.content_side == "item":
            E, G, nn_ = Ei, B, n
        else:
            E, G, nn_ = Eu, B, 1
        acc = torch.zeros(1, dtype=torch.float64, device=E.device)
        dE = torch.zeros_like(E)
        weight = float(G * nn_) if self.embedding_loss_aggregator == "sum" else 1.0
        ops.infonce(E, G, nn_, self.embedding_dim, self.temperature, weight, acc, dE, 1)
        loss = acc.to(F32)
        return (loss, None, dE) if self.content_side == "item" else (loss, dE, None)

    def get_and_reset_other_loss(self) -> Dict:
        emb_loss = self.emb_loss
        self.emb_loss = 0
        if not torch.is_tensor(emb_loss):
            emb_loss = torch.zeros(1, device=self.device)
        return {"reg_loss": emb_loss}  # (the reference does not apply lambda_content either: sgd_alg.py:1492-1497)

    @classmethod
    def build_from_conf(cls, conf: dict, dataset):
        return cls(dataset, conf["feature_name"], conf["aggregate_for_rec"], conf["lambda_content"], conf["temperature"],
                   conf["embedding_loss_aggregator"], conf["intermediate_layers"], conf["embedding_dim"],
                   conf["use_user_bias"], conf["use_item_bias"], conf["use_global_bias"])


class ItemFeatureMatrixFactorization(_FeatureMF):
    """reference ``ItemFeatureMatrixFactorization`` (sgd_alg.py:1399-1505)"""
    content_side = "item"


class UserFeatureMatrixFactorization(_FeatureMF):
    """reference ``UserFeatureMatrixFactorization`` (sgd_alg.py:1508-1614)"""
    content_side = "user"
