"""Reference-API glue: ``combine_user_item_representations`` and the autograd-facing training forward used when a
caller drives the model exactly like the reference's ``Trainer`` (``train/trainer.py:209-223``):

    logits = model(u_idxs, i_idxs)                       # training mode
    loss = rec_loss.compute_loss(logits, labels) + model.get_and_reset_other_loss()['reg_loss']
    loss.backward(); optimizer.step(); optimizer.zero_grad()

The forward runs the same kernels as ``FusedTrainer``; ``backward`` receives d loss / d logits, runs the hand-written
backward kernels and ACCUMULATES into ``param.grad`` (created on demand), so any ``torch.optim`` optimizer works.
The fast path is ``sibrar_b200.trainer.FusedTrainer`` (loss fused into the scoring kernel, fused Adam, CUDA graph).

Regularisation losses (InfoNCE) are computed -- value and gradient -- during the forward; their gradient enters the
backward with coefficient 1, which is how the reference combines them (``total_loss = rec_loss + reg_loss``,
``train/trainer.py:215``).
"""
from __future__ import annotations

import torch

from . import ops

F32 = torch.float32


def combine(u_repr: torch.Tensor, i_repr: torch.Tensor) -> torch.Tensor:
    """``einsum('be,ce->bc')`` for 2-D item representations (full catalog) and ``einsum('be,bce->bc')`` for 3-D ones
    (sgd_alg.py:2093-2114), computed by the score kernels (bf16 tcgen05 GEMM / fused score kernel)."""
    return _combine_nograd(u_repr.detach(), i_repr.detach())


def _combine_nograd(u_repr, i_repr):
    B, D = u_repr.shape
    if i_repr.dim() == 2:
        I = i_repr.shape[0]
        out = torch.empty((B, I), dtype=F32, device=u_repr.device)
        u16, i16 = ops.cast_bf16(u_repr.contiguous()), ops.cast_bf16(i_repr.contiguous())
        ops.gemm(u16, i16, B, I, D, out_f32=out)
        return out
    n = i_repr.shape[1]
    out = torch.empty((B, n), dtype=F32, device=u_repr.device)
    ops.score_loss(u_repr.contiguous(), i_repr.contiguous(), B, n, 1, 1, D, 0, 0, "bce", 0, 0.0, out, None)
    return out


class _TrainForward(torch.autograd.Function):
    """logits = model(u, i) in training mode; backward = the hand-written backward kernels"""

    @staticmethod
    def forward(ctx, anchor, model, u_idxs, i_idxs):
        rt = model._rt()
        user, item = model.user_embedding_module, model.item_embedding_module
        B, n = i_idxs.shape
        D = model.config.shared_common_dim
        ops.tick(rt.step_dev)
        rt.arena.reset()
        inj_mods, inj_keep = getattr(model, "_injected_inputs", None) or ({}, {})  # tests: the reference's draws
        Eu = user.embed(u_idxs, True, inj_mods.get("user"), inj_keep.get("user"))
        Ei = item.embed(i_idxs, True, inj_mods.get("item"), inj_keep.get("item"))
        ku, ki = user.k_train, item.k_train
        logits = torch.empty((B, n), dtype=F32, device=Eu.device)
        ops.score_loss(Eu, Ei, B, n, ku, ki, D, user.agg_max, item.agg_max, "bce", 0, 0.0, logits, None)
        # regularisation: loss values now, gradients kept for the backward
        reg = torch.zeros(4, dtype=torch.float64, device=Eu.device)
        dEu_reg = dEi_reg = None
        if user.reg_enabled:
            c = user.entity_config
            dEu_reg = torch.zeros_like(Eu)
            ops.infonce(Eu, 1, B, D, c.regularization_temperature, c.regularization_weight, reg[1:2], dEu_reg, 1)
            user.regularization_loss = (reg[1:2] / c.regularization_weight).to(F32) if c.regularization_weight else \
                torch.zeros(1, device=Eu.device)
        if item.reg_enabled:
            c = item.entity_config
            dEi_reg = torch.zeros_like(Ei)
            ops.infonce(Ei, B, n, D, c.regularization_temperature, c.regularization_weight, reg[2:3], dEi_reg, 1)
            item.regularization_loss = (reg[2:3] / c.regularization_weight).to(F32) if c.regularization_weight else \
                torch.zeros(1, device=Eu.device)
        ctx.model, ctx.saved = model, (Eu, Ei, dEu_reg, dEi_reg, B, n, D)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        Eu, Ei, dEu_reg, dEi_reg, B, n, D = ctx.saved
        user, item = model.user_embedding_module, model.item_embedding_module
        grads = {}
        for p in model.parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads[id(p)] = p.grad
        dEu, dEi = torch.empty_like(Eu), torch.empty_like(Ei)
        ops.score_bwd(Eu, Ei, B, n, user.k_train, item.k_train, D, user.agg_max, item.agg_max,
                      dlogits.contiguous().to(F32), dEu, dEi)
        if dEu_reg is not None:
            dEu += dEu_reg
        if dEi_reg is not None:
            dEi += dEi_reg
        item.backward(dEi, grads)
        user.backward(dEu, grads)
        return None, None, None, None


def train_forward(model, u_idxs, i_idxs):
    anchor = next(iter(model.parameters()))  # makes the output require grad
    return _TrainForward.apply(anchor, model, u_idxs, i_idxs)
