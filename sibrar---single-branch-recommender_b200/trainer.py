"""Fused training step: the body of the reference's ``Trainer._train`` loop (``train/trainer.py:204-223``: forward,
rec loss, reg losses, backward, optimizer step, zero_grad) as one stream of hand-written kernels with no host
synchronisation -- the reference syncs >= 3 times per step through ``.item()`` (``train/trainer.py:217-219``).

Losses are accumulated on the device (fp64) and read on demand; the optimizer is the multi-tensor Adam/AdamW kernel
(``torch.optim.Adam`` / ``AdamW`` semantics, selected like ``train/trainer.py:62-68``).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch

from . import ops
from .config import LearningConfig
from .sbnet import SingleBranchNet, SingleBranchNetEntity

F32 = torch.float32


class FusedTrainer:
    def __init__(self, model: SingleBranchNet, learn, n_negative_samples: int,
                 negative_sampling_strategy: str = "uniform_recbole", betas=(0.9, 0.999), eps: float = 1e-8,
                 grad_scale: float = 1.0, cuda_graph: bool = False):
        if isinstance(learn, dict):
            learn = LearningConfig.from_dict(learn)
        if learn.optimizer not in ("adam", "adamw", "adagrad"):  # train/trainer.py:62-66
            raise KeyError(learn.optimizer)
        if learn.rec_loss not in ("bpr", "bce", "sampled_softmax"):
            raise KeyError(learn.rec_loss)
        assert learn.loss_aggregator in ("mean", "sum"), "Type of Aggregator not yet defined"
        assert negative_sampling_strategy in ("uniform", "uniform_recbole"), \
            "Type of Negative Strategy not currently supported"
        self.model, self.learn = model, learn
        self.n_neg = n_negative_samples
        self.betas, self.eps, self.grad_scale = betas, eps, grad_scale
        self.ssm_shift = 0.0
        if learn.rec_loss == "sampled_softmax" and negative_sampling_strategy == "uniform":
            self.ssm_shift = math.log(model.n_items / n_negative_samples)  # train/rec_losses.py:104-105
        self.rt = model._rt()
        dev = model.device
        self.user, self.item = model.user_embedding_module, model.item_embedding_module
        for ent in (self.user, self.item):
            ent._materialize()
        # persistent fp32 gradient accumulators (= p.grad), Adam moments, bf16 shadows maintained by the optimizer
        self.grads: Dict[int, torch.Tensor] = {}
        shadows = {}
        for ent in (self.user, self.item):
            for chain in ent.chains():
                for st in chain.stages:
                    st.refresh(False)
                    shadows[id(st.linear.weight)] = st.w16
        # one flat fp32 gradient buffer, item-entity parameters first (they finish first in the backward pass, so a
        # data-parallel run can all-reduce that bucket while the user entity is still back-propagating)
        ordered = list(self.item.parameters()) + list(self.user.parameters())
        assert len(ordered) == len(list(model.parameters()))
        offs, total = [], 0
        for p in ordered:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4  # keep every view 16-byte aligned
        self.flat_grads = self._alloc_flat_grads(total, dev)
        n_item = len(list(self.item.parameters()))
        self.bucket_bounds = (0, offs[n_item] if n_item < len(offs) else total, total)  # [item | user]
        entries = []
        self.grad_offsets = {id(p): (o, p.numel(), tuple(p.shape)) for p, o in zip(ordered, offs)}
        for p, o in zip(ordered, offs):
            g = self.flat_grads[o:o + p.numel()].view(p.shape)
            p.grad = g
            self.grads[id(p)] = g
            entries.append(dict(param=p.data, grad=g, exp_avg=torch.zeros_like(g), exp_avg_sq=torch.zeros_like(g),
                                shadow=shadows.get(id(p))))
        self.adam = ops.AdamPlan(entries, dev)
        self.opt_step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        # [rec, user_reg, item_reg] sums since the last reset (fp64), + number of accumulated steps on the host
        self.loss_acc = torch.zeros(4, dtype=torch.float64, device=dev)
        self.steps_accumulated = 0
        self.logits = None
        # CUDA-graph replay of the whole step (the ~40 launches of a step cost more host time than GPU time at
        # paper-sized batches).  Per batch shape: two eager steps first (lazy caches, function attributes), then the
        # step is captured once and replayed; the step counter, the Philox streams and the loss sums live on the device.
        self.cuda_graph = bool(cuda_graph)
        self._graphs = {}
        # user side / item side of the step on two streams (SBR_BRANCHES=0: one stream)
        # 2 (default): + weight-gradient GEMMs and per-modality table chains on further streams
        self.branches = int(os.environ.get("SBR_BRANCHES", "2"))
        if self.branches:
            from . import sbnet as _sbnet
            _sbnet.SPLITK_FILL = 1  # the two entities' split-K GEMMs run side by side (sbnet.SPLITK_FILL)
        self._side = None
        # debugging / parity tests: a copy of the flat gradient buffer taken inside the step right before the optimizer
        # consumes (and clears) it -- also inside the captured graph
        self.snapshot_grads = False
        self.grads_snapshot = None

    # ------------------------------------------------------------------------------------------------ one step
    def step(self, u_idxs: torch.Tensor, i_idxs: torch.Tensor, mods: Optional[dict] = None,
             keep_masks: Optional[dict] = None, apply_optimizer: bool = True):
        """u_idxs int64 [B], i_idxs int64 [B, 1 + n_neg] on the device (positive item in column 0)."""
        if self.cuda_graph and not mods and not keep_masks and apply_optimizer:
            return self._graph_step(u_idxs, i_idxs)
        return self._eager_step(u_idxs, i_idxs, mods, keep_masks, apply_optimizer)

    def _graph_step(self, u_idxs, i_idxs):
        key = (tuple(u_idxs.shape), tuple(i_idxs.shape))
        st = self._graphs.setdefault(key, {"eager": 0})
        if "graph" not in st:
            if st["eager"] < 2:
                st["eager"] += 1
                return self._eager_step(u_idxs, i_idxs, None, None, True)
            st["u"], st["i"] = u_idxs.clone(), i_idxs.clone()
            torch.cuda.synchronize()
            from . import _lib
            before = _lib.launch_counter()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._eager_step(st["u"], st["i"], None, None, True)
            self.steps_accumulated -= 1  # the capture only records
            st["launches"] = _lib.launch_counter() - before
            _lib._launches[0] = before
            st["graph"] = g
        st["u"].copy_(u_idxs, non_blocking=True)
        st["i"].copy_(i_idxs, non_blocking=True)
        st["graph"].replay()
        from . import _lib
        _lib._launches[0] += st["launches"]  # kernels of ours inside the replayed graph
        self.steps_accumulated += 1

    def _eager_step(self, u_idxs, i_idxs, mods, keep_masks, apply_optimizer):
        m, rt = self.model, self.rt
        mods, keep_masks = mods or {}, keep_masks or {}
        # data parallel: the gradient buffers are all-reduced on the step that applies the optimizer only (accumulation
        # steps keep rank-local sums: reducing the persistent buffer every step would count earlier micro-batches
        # world-size times)
        self._reduce_now = bool(apply_optimizer)
        B, n = i_idxs.shape
        D = m.config.shared_common_dim
        # one launch: model step counter (+ Adam's t when this step updates) += 1, accumulator arena cleared
        rt.arena.begin_step(rt.step_dev, self.opt_step_dev if apply_optimizer else None)
        ku, ki = self.user.k_train, self.item.k_train
        # entities that end in a BatchNorm leave it to the score/loss kernel (no normalised copy, no separate
        # BatchNorm-backward reduction) when there is one modality slot per entity and D fits the fused kernel
        fuse = ku == 1 and ki == 1 and D in (16, 32, 64, 128) and not self.user.reg_enabled and not self.item.reg_enabled
        # The user and the item side of a step only meet in the score/loss kernel and in Adam.  Their kernels are small
        # (tens of CTAs for the table projections, one CTA for the BatchNorm finalisers), so the user side runs on a
        # side stream -- a parallel branch of the captured graph -- next to the (longer) item side.
        side = self._side_stream(u_idxs.device, small_batch=B * n <= 16384) if (self.branches and u_idxs.is_cuda) else None
        rt.branches = side is not None and self.branches >= 2
        main = torch.cuda.current_stream() if side is not None else None
        sb_u, sb_i = isinstance(self.user, SingleBranchNetEntity), isinstance(self.item, SingleBranchNetEntity)
        if side is not None:
            imods = mods.get("item")
            if imods is None and sb_i:
                imods = self.item.sample_modalities(i_idxs.numel())
            if sb_i:
                # (the item-side gather plan is built on the side stream: its row keys need the compact positions of the
                # referenced-rows route, so the marking pass runs here, before the fork)
                self.item._materialize()
                self.item.begin_call(i_idxs.reshape(-1).contiguous(), imods, ki)
                self.item._premarked = True
            side.wait_stream(main)
            with torch.cuda.stream(side):
                Eu = self.user.embed(u_idxs, True, mods.get("user"), keep_masks.get("user"), defer_final_bn=fuse)
                if sb_u:
                    self.user.build_plan(u_idxs, self.user._ctx[1], ku, self.grads)
                if sb_i:
                    self.item._materialize()
                    self.item.build_plan(i_idxs, imods, ki, self.grads)
                fwd_done = side.record_event()
            Ei = self.item.embed(i_idxs, True, imods, keep_masks.get("item"), defer_final_bn=fuse)
            main.wait_event(fwd_done)
        else:
            Eu = self.user.embed(u_idxs, True, mods.get("user"), keep_masks.get("user"), defer_final_bn=fuse)
            Ei = self.item.embed(i_idxs, True, mods.get("item"), keep_masks.get("item"), defer_final_bn=fuse)
        if self.logits is None or self.logits.shape != (B, n):
            self.logits = torch.empty((B, n), dtype=F32, device=u_idxs.device)
        bn_u = bn_i = None
        if Eu is None or Ei is None:
            def inline(ent):
                d = dict(ent.sb_chain.deferred)
                d["sums"] = rt.arena.take(ops.BN_SUM_REPLICAS * 2 * D)
                return d
            bn_u = inline(self.user) if Eu is None else None
            bn_i = inline(self.item) if Ei is None else None
            dEu = torch.empty((B, D), dtype=F32, device=u_idxs.device)
            dEi = torch.empty((B * n, D), dtype=F32, device=u_idxs.device)
            ops.score_loss_bn(Eu, bn_u, Ei, bn_i, B, n, D, self.learn.rec_loss, self.learn.loss_aggregator == "sum",
                              self.ssm_shift, self.logits, self.loss_acc[0:1], dEu, dEi)
        else:
            dEu, dEi = torch.empty_like(Eu), torch.empty_like(Ei)
            ops.score_loss(Eu, Ei, B, n, ku, ki, D, self.user.agg_max, self.item.agg_max, self.learn.rec_loss,
                           self.learn.loss_aggregator == "sum", self.ssm_shift, self.logits, self.loss_acc[0:1], dEu,
                           dEi)
        if self.user.reg_enabled:
            c = self.user.entity_config
            ops.infonce(Eu, 1, B, D, c.regularization_temperature, c.regularization_weight, self.loss_acc[1:2], dEu, 1)
        if self.item.reg_enabled:
            c = self.item.entity_config
            ops.infonce(Ei, B, n, D, c.regularization_temperature, c.regularization_weight, self.loss_acc[2:3], dEi, 1)
        if side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.user.backward(dEu, self.grads, final_bn_sums=bn_u["sums"] if bn_u else None)
            self.item.backward(dEi, self.grads, final_bn_sums=bn_i["sums"] if bn_i else None)
            self._after_item_backward()
            main.wait_stream(side)
            self._after_user_backward()
        else:
            self.item.backward(dEi, self.grads, final_bn_sums=bn_i["sums"] if bn_i else None)
            self._after_item_backward()
            self.user.backward(dEu, self.grads, final_bn_sums=bn_u["sums"] if bn_u else None)
            self._after_user_backward()
        if self.snapshot_grads:
            if self.grads_snapshot is None:
                self.grads_snapshot = torch.empty_like(self.flat_grads)
            self.grads_snapshot.copy_(self.flat_grads)
        if apply_optimizer:
            self.optimizer_step(ticked=True)
        self.steps_accumulated += 1

    def _alloc_flat_grads(self, total: int, dev):
        return torch.zeros(total, dtype=F32, device=dev)

    def _side_stream(self, device, small_batch: bool = False):
        """the user branch's stream.  Small batches (fixed-cost kernels everywhere): high priority, so the short user chain
        is never queued behind the item chain's table projections (0.167 vs 0.187 ms per step at B = 256; at B = 16 384 the
        same priority costs 14 %, so large batches keep the default)"""
        if small_batch and os.environ.get("SBR_SIDE_PRIORITY", "auto") != "0":
            if getattr(self, "_side_hi", None) is None:
                self._side_hi = torch.cuda.Stream(device=device, priority=-1)
            return self._side_hi
        if self._side is None:
            self._side = torch.cuda.Stream(device=device)
        return self._side

    # hooks for the data-parallel subclass (gradient all-reduce overlapped with the rest of the backward pass)
    def _after_item_backward(self):
        pass

    def _after_user_backward(self):
        pass

    def optimizer_step(self, ticked: bool = False):
        b1, b2 = self.betas
        if not ticked:
            ops.tick(self.opt_step_dev)  # Adam's t counts the updates of THIS optimizer state (device-side: graph-safe)
        mode = {"adam": 0, "adamw": 1, "adagrad": 2}[self.learn.optimizer]
        eps = 1e-10 if mode == 2 and self.eps == 1e-8 else self.eps  # torch.optim.Adagrad's default eps
        self.adam.step(self.learn.lr, b1, b2, eps, self.learn.wd, mode, self.opt_step_dev, self.grad_scale)

    # ------------------------------------------------------------------------------------------------ logging
    def read_losses(self, reset: bool = True) -> Dict[str, float]:
        """host sync.  Mean per-step losses since the last reset, keyed like the reference's epoch dict
        (``train/trainer.py:217-219,233``)."""
        acc = self.loss_acc.cpu().tolist()
        n = max(1, self.steps_accumulated)
        out = {"train/rec_loss": acc[0] / n, "train/reg_loss": (acc[1] + acc[2]) / n,
               "train/loss": (acc[0] + acc[1] + acc[2]) / n}
        if self.model.is_user_sb_module:
            out["train/user_reg_loss"] = acc[1] / n
        if self.model.is_item_sb_module:
            out["train/item_reg_loss"] = acc[2] / n
        if reset:
            self.loss_acc.zero_()
            self.steps_accumulated = 0
        return out


class DeviceBatchFeeder:
    """Device-side replacement of ``NegativeSamplingDataLoader`` over ``TrainRecDataset`` (``data/dataloader.py:128-198``,
    ``data/dataset.py:380-396``): every epoch visits each train interaction exactly once, in a fresh random order when
    ``shuffle`` (the reference's ``DataLoader(shuffle=True)``), in batches of ``batch_size`` (the last one shorter);
    each slot gets ``n_negative_samples`` 'uniform_recbole' negatives drawn by ``sbr_sample_epoch_batch``.  No host work
    per batch: the COO arrays, the sorted train CSR and the epoch permutation live on the device.

        feeder = DeviceBatchFeeder(train_dataset, batch_size=16384, device="cuda")
        for epoch in range(n_epochs):
            for u_idxs, i_idxs in feeder.epoch():        # int64 [b], int64 [b, 1 + n_neg], positive in column 0
                trainer.step(u_idxs, i_idxs)
    """

    def __init__(self, dataset, batch_size: int, device, shuffle: bool = True, seed: int = 0,
                 n_negative_samples: Optional[int] = None):
        import numpy as np
        strategy = getattr(dataset, "negative_sampling_strategy", "uniform_recbole")
        # 'uniform_recbole' is the dataloader-level sampler (data/dataloader.py:145-147); 'uniform' is the reference's
        # dataset-level sampler (use_dataset_negative_sampler, data/dataset.py:361-375, data/sampling.py:7-32)
        if strategy not in {"uniform_recbole", "uniform"}:
            raise ValueError(f"sampling strategy {strategy} not supported for dataloader sampling!")
        self.strategy = strategy
        self.device = torch.device(device)
        self.batch_size, self.shuffle, self.seed = int(batch_size), shuffle, int(seed)
        self.n_neg = int(dataset.n_negative_samples if n_negative_samples is None else n_negative_samples)
        coo = dataset.interaction_matrix
        csr = dataset.user_sampling_matrix.tocsr()
        csr.sort_indices()
        n_choices, row_len = len(dataset.items_in_split), np.diff(csr.indptr)
        if n_choices - int(row_len.max(initial=0)) < self.n_neg:
            raise ValueError(f'Not enough values in the range to sample "{self.n_neg}" unique values.')
        dev = lambda a, t: torch.from_numpy(np.ascontiguousarray(a).astype(t)).to(self.device)  # noqa: E731
        self.coo_u, self.coo_i = dev(coo.row, np.int32), dev(coo.col, np.int32)
        self.indptr, self.indices = dev(csr.indptr, np.int64), dev(csr.indices, np.int32)
        self.items = dev(dataset.items_in_split, np.int32)
        self.item_pos = None
        if strategy == "uniform":
            if n_choices - int(row_len.max(initial=0)) < self.n_neg:
                raise ValueError(f'Not enough values in the range to sample "{self.n_neg}" unique values.')
            pos = np.full(csr.shape[1], -1, dtype=np.int32)
            pos[np.asarray(dataset.items_in_split)] = np.arange(n_choices, dtype=np.int32)
            self.item_pos = dev(pos, np.int32)
        self.nnz = int(coo.nnz)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.gen = torch.Generator(device=self.device).manual_seed(self.seed)
        self.epochs_done = 0

    def __len__(self):
        return -(-self.nnz // self.batch_size)

    def epoch(self):
        order = torch.randperm(self.nnz, device=self.device, generator=self.gen) if self.shuffle else \
            torch.arange(self.nnz, device=self.device)
        for off in range(0, self.nnz, self.batch_size):
            b = min(self.batch_size, self.nnz - off)
            u = torch.empty(b, dtype=torch.int64, device=self.device)
            i = torch.empty((b, 1 + self.n_neg), dtype=torch.int64, device=self.device)
            ops.tick(self.step_dev)
            ops.sample_negatives(self.coo_u, self.coo_i, order, off, self.indptr, self.indices, self.items, self.item_pos,
                                 b, self.n_neg, self.strategy, self.seed, self.step_dev, u, i)
            yield u, i
        self.epochs_done += 1
