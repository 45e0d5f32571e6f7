"""Reference-API glue: ``combine_user_item_representations`` and the autograd-facing training forward used when a
caller drives the model like the reference's ``Trainer`` (``model(u, i)`` -> ``loss.backward()`` ->
``torch.optim``).  The fast path is ``sibrar_b200.trainer.FusedTrainer``; both run the same kernels."""
from __future__ import annotations

import torch

from . import ops

F32 = torch.float32


def combine(u_repr: torch.Tensor, i_repr: torch.Tensor) -> torch.Tensor:
    """``einsum('be,ce->bc')`` for 2-D item representations (full catalog) and ``einsum('be,bce->bc')`` for 3-D ones
    (sgd_alg.py:2093-2114), computed by the score kernels (bf16 tcgen05 GEMM / fused score kernel)."""
    if u_repr.requires_grad or i_repr.requires_grad:
        return _ScoreFn.apply(u_repr, i_repr)
    return _combine_nograd(u_repr, i_repr)


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


class _ScoreFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, i):
        ctx.save_for_backward(u, i)
        return _combine_nograd(u.detach(), i.detach())

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError("autograd through combine_user_item_representations: use FusedTrainer.step")


def entity_forward_with_grad(model, ent, idx):
    raise NotImplementedError("training through the autograd API is not wired up yet; use "
                              "sibrar_b200.trainer.FusedTrainer.step(u_idxs, i_idxs) (same kernels, no host syncs)")
