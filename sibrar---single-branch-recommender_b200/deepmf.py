"""DeepMatrixFactorization on the sibrar_b200 kernels (reference
``DeepMatrixFactorization``, ``algorithms/sgd_alg.py:1141-1276``; Xue et al., IJCAI 2017).

``user_nn`` maps a user's row of the train interaction matrix, ``item_nn`` an item's column, to ``final_dimension``
(PolyLinear with ReLU between the layers, optional ReLU on the output); the score is the cosine similarity of the two,
clamped from below at ``mu``.  Same constructor arguments, ``build_from_conf`` keys, ``forward / predict /
get_{user,item}_representations / combine_user_item_representations`` and ``state_dict()`` keys
(``{user,item}_nn.layers.linear_i.{weight,bias}``) as the reference.

How it runs: both towers are functions of the ENTITY ROW alone, so each step projects the whole (bit-packed / CSR,
optionally row-normalised) interaction matrix once per tower (``PlainEntity``: tcgen05 GEMMs, table-level backward) and the
batch gathers L2-normalised rows of the two tables (the gather kernels' normalise + its backward); the cosine is then the
plain dot product of the score kernel, followed by the lower clamp (``sbr_clamp_min_fwd / bwd``).  ``model(u, i)`` in
training mode is differentiable (hand-written backward behind an ``autograd.Function``).

One documented deviation: ``get_*_representations`` return L2-normalised vectors also with
``normalize_representations=False`` (their dot product is the reference's cosine either way; the evaluator's fused
score + top-k kernel takes dot products).  The clamp only matters for the ORDER of items whose cosine is below ``mu``
(ties at ``mu`` in the reference); the evaluator ranks those by their raw cosine.
"""
from __future__ import annotations

import os
import weakref
from typing import Dict, List, Union

import numpy as np
import scipy.sparse as sp
import torch
from torch import nn

from . import ops
from .config import FeatureModuleConfig
from .sbnet import FeatureEmbedding, PlainEntity, PolyLinear, _Runtime

F32 = torch.float32


def _row_normalised(m: sp.csr_matrix) -> sp.csr_matrix:
    """rows / max(||row||_2, 1e-8)  (sgd_alg.py:1209-1211)"""
    m = sp.csr_matrix(m, dtype=np.float64)
    norm = np.sqrt(np.asarray(m.multiply(m).sum(axis=1)).reshape(-1))
    return sp.diags(1.0 / np.maximum(norm, 1e-8)).dot(m).tocsr().astype(np.float32)


class DeepMatrixFactorization(nn.Module):
    def __init__(self, dataset, u_mid_layers: Union[List[int], int], i_mid_layers: Union[List[int], int],
                 final_dimension: int, mu: float = 1.e-6, normalize_interactions: bool = False,
                 normalize_representations: bool = False, use_output_activation_fn: bool = False):
        super().__init__()
        from .synthetic import SynFeature
        self.dataset = dataset
        self.normalize_interactions, self.normalize_representations = normalize_interactions, normalize_representations
        self.mu, self.final_dimension = mu, final_dimension
        u_mid = [u_mid_layers] if isinstance(u_mid_layers, int) else list(u_mid_layers)
        i_mid = [i_mid_layers] if isinstance(i_mid_layers, int) else list(i_mid_layers)
        self.u_layers = [dataset.n_items] + u_mid + [final_dimension]
        self.i_layers = [dataset.n_users] + i_mid + [final_dimension]
        output_fn = "relu" if use_output_activation_fn else None
        u_mat = sp.csr_matrix(dataset.user_sampling_matrix_train, dtype=np.float32)
        i_mat = getattr(dataset, "item_sampling_matrix_train", None)
        i_mat = sp.csr_matrix(i_mat, dtype=np.float32) if i_mat is not None else u_mat.T.tocsr()
        if normalize_interactions:
            u_mat, i_mat = _row_normalised(u_mat), _row_normalised(i_mat)
        engines = {}
        for side, mat, layers, mid in (("user", u_mat, self.u_layers, u_mid), ("item", i_mat, self.i_layers, i_mid)):
            feat = SynFeature(f"{side}_interactions", "vector", mat)
            fe = FeatureEmbedding(feat, embedding_dim=final_dimension, pre_embedding_layers=mid, activation_fn="relu")
            # (FeatureEmbedding puts the activation on its output as well; DeepMF's towers make that optional)
            fe.pre_embedding_layers = PolyLinear(layers, activation_fn="relu", output_fn=output_fn)
            setattr(self, f"{side}_nn", fe.pre_embedding_layers)
            eng = PlainEntity(feat, FeatureModuleConfig(feature_name=feat.feature_definition.name,
                                                        embedding_dim=final_dimension, pre_embedding_layers=mid),
                              mat.shape[0], fe=fe)
            eng.normalize_output = True
            engines[side] = eng
        object.__setattr__(self, "_engines", engines)
        self.name = "DeepMatrixFactorization"
        self._runtime = None
        ref = weakref.ref(self)
        for e in engines.values():
            e._owner = lambda r=ref: r()._rt()

    @staticmethod
    def build_from_conf(conf: dict, train_dataset):
        return DeepMatrixFactorization(dataset=train_dataset, u_mid_layers=conf.get("u_mid_layers", []),
                                       i_mid_layers=conf.get("i_mid_layers", []),
                                       final_dimension=conf["final_dimension"], mu=conf.get("mu", 1e-6),
                                       normalize_interactions=conf.get("normalize_interactions", False),
                                       normalize_representations=conf.get("normalize_representations", False),
                                       use_output_activation_fn=conf.get("use_output_activation_fn", False))

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

    # ---- reference API
    def _represent(self, side, idx):
        out = self._engines[side].embed(idx.reshape(-1).contiguous(), self.training)
        return out.view(*idx.shape, -1)

    @torch.no_grad()
    def get_user_representations(self, u_idxs: torch.Tensor):
        self._rt().arena.reset()
        return self._represent("user", u_idxs)

    @torch.no_grad()
    def get_item_representations(self, i_idxs: torch.Tensor):
        self._rt().arena.reset()
        return self._represent("item", i_idxs)

    @torch.no_grad()
    def combine_user_item_representations(self, u_repr, i_repr):
        from .autograd import _combine_nograd
        sim = _combine_nograd(u_repr, i_repr)
        ops.clamp_min_fwd(sim, self.mu)
        return sim

    def forward(self, u_idxs, i_idxs):
        if self.training and torch.is_grad_enabled():
            anchor = next(iter(self.parameters()))
            return _DMFTrainForward.apply(anchor, self, u_idxs, i_idxs)
        return self._forward_nograd(u_idxs, i_idxs, None)[0]

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
        sd = torch.load(os.path.join(path, "model.pth"), map_location=self.device)
        sd.pop("user_vectors.weight", None)  # (checkpoints of the reference's previous version, sgd_alg.py:1248-1252)
        sd.pop("item_vectors.weight", None)
        self.load_state_dict(sd)
        print("Model Loaded")

    # ---- the step on the kernels
    def _forward_nograd(self, u_idxs, i_idxs, clamped):
        rt = self._rt()
        ops.tick(rt.step_dev)
        rt.arena.reset()
        Eu = self._represent("user", u_idxs)   # L2-normalised rows
        Ei = self._represent("item", i_idxs)
        B, D = Eu.shape
        if i_idxs.dim() == 1:
            from .autograd import _combine_nograd
            sim = _combine_nograd(Eu, Ei)
        else:
            n = i_idxs.shape[1]
            sim = torch.empty((B, n), dtype=F32, device=Eu.device)
            ops.score_loss(Eu.contiguous(), Ei.contiguous(), B, n, 1, 1, D, 0, 0, "bce", 0, 0.0, sim, None)
        ops.clamp_min_fwd(sim, self.mu, clamped)
        return sim, Eu, Ei


class _DMFTrainForward(torch.autograd.Function):
    """sim = model(u, i) in training mode; backward = the hand-written backward kernels (accumulates into ``param.grad``)"""

    @staticmethod
    def forward(ctx, anchor, model, u_idxs, i_idxs):
        clamped = torch.empty(i_idxs.shape, dtype=torch.uint8, device=i_idxs.device)
        sim, Eu, Ei = model._forward_nograd(u_idxs, i_idxs, clamped)
        ctx.model, ctx.saved = model, (i_idxs.shape, Eu, Ei, clamped)
        return sim

    @staticmethod
    def backward(ctx, dsim):
        model = ctx.model
        (B, n), Eu, Ei, clamped = ctx.saved
        D = Eu.shape[1]
        grads = {}
        for p in model.parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads[id(p)] = p.grad
        dl = dsim.contiguous().to(F32).clone()
        ops.clamp_min_bwd(dl, clamped)
        su, si = Eu.contiguous().view(B, 1, D), Ei.contiguous().view(B * n, 1, D)
        dEu, dEi = torch.empty_like(su), torch.empty_like(si)
        ops.score_bwd(su, si, B, n, 1, 1, D, 0, 0, dl, dEu, dEi)
        eng = model._engines
        eng["item"].backward(dEi.view(B * n, D), grads)   # (gather backward incl. the L2-normalise backward)
        eng["user"].backward(dEu.view(B, D), grads)
        return None, None, None, None
