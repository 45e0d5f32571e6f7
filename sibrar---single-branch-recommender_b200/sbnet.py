"""SingleBranchNet on B200: drop-in for the reference's algorithm API (``algorithms/sgd_alg.py:2009-2144``,
``algorithms/base_classes.py:87-170``) with every arithmetic step executed by the hand-written sm_100a kernels.

Same constructor ``(config, dataset)``, ``build_from_conf(conf: dict, dataset)``, ``forward / predict /
get_user_representations / get_item_representations / combine_user_item_representations /
get_and_reset_other_loss / save_model_to_path / load_model_from_path`` and the same ``state_dict()`` keys
(the modules below only hold parameters under the reference's names; they never run torch arithmetic).

B200-first design:
  * feature store resident in HBM (``feature_store.py``); no host work per step;
  * "entity-table projection": each step projects ALL rows of a dense/sparse modality once
    (``T_m = act(X_m W^T + b)``, a tcgen05 GEMM or a CSR SpMM), batch rows then gather 1 row of ``T_m`` each
    -- the reference re-projects every batch row (``sgd_alg.py:1960-1974``); results are identical;
  * SB-net = chain of fused stages ``Linear -> [act] -> [BatchNorm -> [act]]``; BN statistics come out of the GEMM
    epilogue; dgrad epilogues apply the previous activation's derivative and emit the bias gradient;
  * gradients accumulate into persistent fp32 buffers; one multi-tensor Adam(W) launch updates fp32 masters and
    their bf16 shadows.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch
from torch import nn

from . import ops
from ._lib import SRC_CATEGORICAL, SRC_TABLE, SRC_TAG
from .config import (EmbeddingRegularizationType, FeatureModuleConfig, SingleBranchNetConfig,
                     SingleBranchNetEntityConfig)
from .feature_store import DeviceFeature, feature_type

BF16, F32 = torch.bfloat16, torch.float32
ACTIVATIONS = ("relu", "tanh", "sigmoid", "selu")


# ================================================================================================ parameter holders
def _kaiming_relu_(linear: nn.Linear):
    """``general_weight_init`` of the reference for nn.Linear (train/utils.py:5-10)"""
    nn.init.kaiming_uniform_(linear.weight, nonlinearity="relu")
    if linear.bias is not None:
        nn.init.constant_(linear.bias, 0)


class PolyLinear(nn.Module):
    """Holds the parameters of the reference's PolyLinear under the same names (modules/polylinear.py:50-71):
    ``layers.linear_{i}``, ``layers.batch_norm_{i}``, ``layers.batch_norm``.  ``spec`` lists the ops in order."""

    def __init__(self, layer_config, activation_fn="relu", output_fn="relu", apply_batch_norm_every: int = 0):
        super().__init__()
        assert len(layer_config) > 1, "For a linear network, we at least need one input and one output dimension"
        for a in (activation_fn, output_fn):
            if a is not None and a not in ACTIVATIONS:
                raise ValueError(f'activation "{a}" is not supported (choose from {ACTIVATIONS})')
        self.layer_config = list(layer_config)
        mods, spec = OrderedDict(), []
        n_layers = len(layer_config) - 1
        for i, (d1, d2) in enumerate(zip(layer_config[:-1], layer_config[1:])):
            mods[f"linear_{i}"] = nn.Linear(d1, d2)
            spec.append(("linear", f"linear_{i}"))
            if apply_batch_norm_every > 0 and (i + 1) % apply_batch_norm_every == 0:
                mods[f"batch_norm_{i}"] = nn.BatchNorm1d(d2)
                spec.append(("bn", f"batch_norm_{i}"))
            if i < n_layers - 1:
                spec.append(("act", activation_fn))
        if apply_batch_norm_every == -1:
            mods["batch_norm"] = nn.BatchNorm1d(layer_config[-1])
            spec.append(("bn", "batch_norm"))
        if output_fn is not None:
            spec.append(("act", output_fn))
        self.layers = nn.ModuleDict(mods)
        self.spec = spec

    def forward(self, *a, **k):
        raise RuntimeError("PolyLinear is a parameter container; arithmetic runs in the sibrar_b200 kernels")


class _Identity(nn.Module):
    """placeholder that keeps the reference's nn.Sequential numbering (a Dropout module shifts sb_net indices)"""

    def forward(self, x):
        return x


# ================================================================================================ fused stages
class LinearStage:
    """Linear -> [act1] -> [BatchNorm -> [act2]]  run by kernels; holds bf16 shadows of the fp32 master weight."""

    def __init__(self, linear: nn.Linear):
        self.linear = linear
        self.act1: Optional[str] = None
        self.bn: Optional[nn.BatchNorm1d] = None
        self.act2: Optional[str] = None
        self.in_f, self.out_f = linear.in_features, linear.out_features
        self.w16 = None     # bf16 [out, pad8(in)]: K-major B of the forward GEMM, MN-major B of the dgrad GEMM
        self.wt16 = None    # bf16 [in, pad8(out)]: the transposed shadow, gathered row by row on the CSR route
        self._ver = -1
        # saved for backward
        self.x = self.y16 = self.y32 = self.a32 = self.mi = None

    def refresh(self, need_wt16: bool):
        w = self.linear.weight
        if self.w16 is None or self.w16.device != w.device:
            self.w16 = torch.zeros((self.out_f, ops.pad8(self.in_f)), dtype=BF16, device=w.device)
            self._ver = -1
        if w._version != self._ver:
            ops.cast_bf16(w.detach(), self.w16)
            self._ver = w._version
        if need_wt16:  # refreshed every forward: the fused optimizer only maintains the row-major bf16 shadow
            if self.wt16 is None or self.wt16.device != w.device:
                self.wt16 = torch.zeros((self.in_f, ops.pad8(self.out_f)), dtype=BF16, device=w.device)
            ops.transpose_bf16(w.detach(), self.wt16)


def build_stages(poly: PolyLinear, trailing_bn: Optional[nn.BatchNorm1d] = None):
    stages = []
    for kind, arg in poly.spec:
        if kind == "linear":
            stages.append(LinearStage(poly.layers[arg]))
        elif kind == "bn":
            assert stages[-1].bn is None
            stages[-1].bn = poly.layers[arg]
        else:
            st = stages[-1]
            if st.bn is None:
                assert st.act1 is None
                st.act1 = arg
            else:
                assert st.act2 is None
                st.act2 = arg
    if trailing_bn is not None:
        assert stages[-1].bn is None
        stages[-1].bn = trailing_bn
    return stages


class ZeroArena:
    """small accumulators of one step (BN sums, loss sums, ...) carved out of one buffer cleared by one memset"""

    def __init__(self, device, nbytes: int = 1 << 20):
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.off = 0
        self.high = 0  # everything ever handed out (= everything that can be non-zero) lies in [0, high)

    def reset(self):
        self.buf.zero_()
        self.off = 0

    def begin_step(self, counter0=None, counter1=None):
        """``reset`` of the used part fused with the step-counter increments (one launch, sbr_step_begin)"""
        nbytes = min(self.buf.numel(), max(4096, (self.high + 255) // 256 * 256))
        ops.step_begin(counter0, counter1, self.buf, nbytes)
        self.off = 0

    def take(self, n: int, dtype=F32) -> torch.Tensor:
        size = torch.empty((), dtype=dtype).element_size() * n
        start = (self.off + 15) // 16 * 16
        if start + size > self.buf.numel():
            raise RuntimeError("ZeroArena exhausted")
        self.off = start + size
        self.high = max(self.high, self.off)
        return self.buf[start:start + size].view(dtype)


# CTAs per SM a split-K projection / wgrad GEMM aims at.  2 fills the machine with one GEMM; under the two-branch step
# (FusedTrainer sets 1) the user-side and the item-side GEMM are launched together, and two 145-CTA grids that run SIDE BY
# SIDE (23 us) beat two 290-CTA grids that run one after the other (2 x 16 us): scripts/step_timeline.py
SPLITK_FILL = 2


def splitk_fill() -> int:
    e = os.environ.get("SBR_SPLITK_FILL")
    return int(e) if e else SPLITK_FILL


def run_branches(thunks, streams):
    """fork/join: ``thunks[0]`` runs on the current stream, ``thunks[j]`` on ``streams[j - 1]``, all of them after
    what the current stream holds so far; the current stream continues after all of them.  Under CUDA-graph capture
    these become parallel branches of the graph.  ``streams`` None / too short: plain sequential execution."""
    if len(thunks) <= 1 or not streams or len(streams) < len(thunks) - 1:
        for t in thunks:
            t()
        return
    cur = torch.cuda.current_stream()
    for t, s in zip(thunks[1:], streams):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            t()
    thunks[0]()
    for s in streams[:len(thunks) - 1]:
        cur.wait_stream(s)


class BnSync:
    """BatchNorm statistics over the GLOBAL batch under data parallelism (optional: ``DataParallelTrainer(sync_bn=True)``).
    Forward: the per-CTA rows of partial column sums of every rank are all-gathered and added in a fixed order
    (deterministic, identical on every rank) -- mean / variance / running statistics are those of the reference's single
    process at the same global batch (``nn.BatchNorm1d`` over all rows, modules/polylinear.py:58-61).  Backward: the
    column sums of (dy, dy * xhat) are all-reduced, pre-multiplied by 1 / world: the kernels divide by the LOCAL row
    count, so that they see sum_global / N_global."""

    def __init__(self, world: int):
        self.world = int(world)

    def gather_stats(self, stats: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist
        out = torch.empty((self.world * stats.shape[0], stats.shape[1]), dtype=stats.dtype, device=stats.device)
        dist.all_gather_into_tensor(out, stats.contiguous())
        return out

    def reduce_sums(self, sums: torch.Tensor):
        import torch.distributed as dist
        dist.all_reduce(sums, op=dist._make_nccl_premul_sum(1.0 / self.world))


class Chain:
    """A stack of LinearStages.  forward: bf16 rows (or a CSR feature) -> fp32 output; backward: hand-written."""
    bn_sync: Optional[BnSync] = None

    def __init__(self, stages, feature: Optional[DeviceFeature] = None):
        self.stages = stages
        self.feature = feature  # first-stage input when the chain projects a feature table
        self.rows = 0

    @property
    def csr_input(self):
        return self.feature is not None and self.feature.kind == "csr"

    @property
    def bits_input(self):
        return self.feature is not None and self.feature.kind == "bits"

    @staticmethod
    def _fwd_split_k(rows, st, n_sms: int = 148) -> int:
        bn = 64 if st.out_f <= 64 else (128 if st.out_f <= 128 else 256)
        tiles = -(-rows // 128) * -(-st.out_f // bn)
        num_kb = -(-st.in_f // 64)
        if tiles * 2 > n_sms or num_kb < 16:
            return 1
        return max(1, min(num_kb // 4, (splitk_fill() * n_sms) // tiles))

    def forward(self, x16, rows, training, arena: ZeroArena, keep_for_backward=True, out32=None,
                defer_final_bn=False, ref=None):
        """``defer_final_bn``: when the chain ends in BatchNorm (no activation after it) the normalised output is not
        written; ``self.deferred`` = dict(z, mean_invstd, gamma, beta) lets the score/loss kernel apply it inline."""
        dev = self.stages[0].linear.weight.device
        self.rows = rows
        self.deferred = None
        # ref: the "referenced rows" route -- the chain runs on the compact list of feature rows this step touches
        # (dict(list, count, pos, units, n_units, max_units, gT)); x16 is then the gathered copy of those rows
        self._ref = ref
        self._first_x = x16 if ref is not None else None
        last = len(self.stages) - 1
        for si, st in enumerate(self.stages):
            final = si == last
            first_csr = si == 0 and self.csr_input
            st.refresh(first_csr)
            bias = st.linear.bias.detach() if st.linear.bias is not None else None
            out_pad = ops.pad8(st.out_f)
            y16 = y32 = a32 = mi = None
            if st.bn is None:
                if final:
                    y32 = out32 if out32 is not None else torch.empty((rows, st.out_f), dtype=F32, device=dev)
                if not final:
                    y16 = torch.empty((rows, out_pad), dtype=BF16, device=dev)
                if first_csr:
                    ip, ix = self.feature.csr
                    # gather-sum of bf16 weight rows, fp32 accumulation: the rounding points of the dense route
                    seg = self.feature.csr_seg
                    if (seg is not None or ref is not None) and y32 is None:  # (partial sums need an fp32 home)
                        y32 = torch.empty((rows, st.out_f), dtype=F32, device=dev)
                    if ref is None:
                        ops.spmm_csr(ip, ix, rows, st.wt16, st.out_f, bias, st.act1, y32, vals=self.feature.csr_vals,
                                     out_bf16=y16, segments=seg)
                    else:
                        # only the listed rows (or their segments): raw sums into the cleared compact rows, then one
                        # pass applies bias + activation (+ the bf16 copy)
                        y32.zero_()
                        ops.spmm_csr(ip, ix, ref["max_units"], st.wt16, st.out_f, None, None, y32,
                                     vals=self.feature.csr_vals, row_list=ref["units"], n_rows_dev=ref["n_units"],
                                     segments=seg, out_pos=ref["pos"])
                        ops.splitk_reduce(y32, 1, rows, st.out_f, bias=bias, act=st.act1, out_f32=y32, out_bf16=y16)
                else:
                    first_bits = si == 0 and self.bits_input
                    mm = (lambda *a, **k: ops.gemm_bits(self.feature.bits, *a[1:], **k)) if first_bits else ops.gemm
                    split = ops.effective_splits(st.in_f, self._fwd_split_k(rows, st))
                    if split > 1:
                        # few output tiles, long contraction (an 'interactions' table): every K partition writes its
                        # own fp32 slice (no atomics, deterministic), one pass sums them and applies bias + activation
                        part = torch.empty((split, rows, st.out_f), dtype=F32, device=dev)
                        mm(x16, st.w16, rows, st.out_f, st.in_f, out_f32=part.view(split * rows, st.out_f),
                           split_k=split, split_stride=rows * st.out_f)
                        ops.splitk_reduce(part, split, rows, st.out_f, bias=bias, act=st.act1, out_f32=y32,
                                          out_bf16=y16)
                    else:
                        mm(x16, st.w16, rows, st.out_f, st.in_f, bias=bias, act=st.act1, out_bf16=y16, out_f32=y32)
            else:
                bn = st.bn
                a32 = torch.empty((rows, st.out_f), dtype=F32, device=dev)
                mi = torch.empty(2 * st.out_f, dtype=F32, device=dev)
                # deterministic batch statistics: one row of partial sums per epilogue warp, added in a fixed order
                n_part = ops.gemm_colstats_rows(rows, st.out_f) if training else 0
                stats = torch.empty((n_part, 2 * st.out_f), dtype=F32, device=dev) if training else None
                if first_csr:
                    raise NotImplementedError("BatchNorm directly on a sparse feature projection")
                if si == 0 and self.bits_input:
                    ops.gemm_bits(self.feature.bits, st.w16, rows, st.out_f, st.in_f, bias=bias, act=st.act1,
                                  out_f32=a32, colstats=stats, colstats_rows=n_part)
                else:
                    ops.gemm(x16, st.w16, rows, st.out_f, st.in_f, bias=bias, act=st.act1, out_f32=a32,
                             colstats=stats, colstats_rows=n_part)
                if training:
                    rows_g = rows
                    if self.bn_sync is not None:  # statistics of the global batch (equal rows per rank)
                        stats, n_part, rows_g = self.bn_sync.gather_stats(stats), n_part * self.bn_sync.world, \
                            rows * self.bn_sync.world
                    ops.bn_finalize(stats, rows_g, st.out_f, mi, bn.running_mean, bn.running_var,
                                    bn.num_batches_tracked, eps=bn.eps, momentum=bn.momentum, n_partials=n_part)
                else:
                    ops.bn_eval_coeffs(bn.running_mean, bn.running_var, st.out_f, mi, eps=bn.eps)
                if final and defer_final_bn and training and st.act2 is None:
                    self.deferred = dict(z=a32, mean_invstd=mi, gamma=bn.weight.detach(), beta=bn.bias.detach())
                else:
                    if final:
                        y32 = out32 if out32 is not None else torch.empty((rows, st.out_f), dtype=F32, device=dev)
                    else:
                        y16 = torch.empty((rows, out_pad), dtype=BF16, device=dev)
                    ops.bn_apply(a32, mi, bn.weight.detach(), bn.bias.detach(), st.act2, rows, st.out_f, out_bf16=y16,
                                 out_f32=y32)
            if keep_for_backward:
                st.x, st.y16, st.y32, st.a32, st.mi = x16, y16, y32, a32, mi
            x16 = y16
        return y32

    def backward(self, dy32, grads: Dict[int, torch.Tensor], need_dx: bool, arena: ZeroArena, zero_dy=False,
                 final_bn_sums=None, wgrad_stream=None):
        """dy32: fp32 [rows, out] gradient w.r.t. the chain output.  Accumulates parameter gradients into
        ``grads[id(param)]``; returns fp32 [rows, in] gradient w.r.t. the chain input if ``need_dx``.
        ``wgrad_stream``: the weight-gradient GEMMs (needed by the optimizer only) run there, next to the dz -> dx
        chain on the current stream; joined before returning."""
        if wgrad_stream is None or len(self.stages) < 2:
            return self._backward(dy32, grads, need_dx, arena, zero_dy, final_bn_sums, None, None)
        keep = []  # operands of the side-stream GEMMs stay allocated until the join
        try:
            return self._backward(dy32, grads, need_dx, arena, zero_dy, final_bn_sums, wgrad_stream, keep)
        finally:
            torch.cuda.current_stream().wait_stream(wgrad_stream)
            keep.clear()

    def _backward(self, dy32, grads, need_dx, arena, zero_dy, final_bn_sums, wstream, keep):
        rows = self.rows
        dev = dy32.device
        n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
        fused = None  # (dz16, dz32) of the current stage when the following stage's dgrad produced it
        for si in reversed(range(len(self.stages))):
            st = self.stages[si]
            first_csr = si == 0 and self.csr_input
            out_pad = ops.pad8(st.out_f)
            g_w = grads[id(st.linear.weight)]
            g_b = grads[id(st.linear.bias)] if st.linear.bias is not None else None
            if fused is not None:
                dz16, dz32 = fused
                fused = None
            else:
                y = st.y32 if st.y32 is not None else st.y16
                dz16 = torch.zeros((rows, out_pad), dtype=BF16, device=dev) if (first_csr and out_pad != st.out_f) \
                    else torch.empty((rows, out_pad), dtype=BF16, device=dev)
                dz32 = None
                if st.bn is None:
                    ops.actgrad_colsum(dy32, y, st.act1, rows, st.out_f, out_bf16=dz16, out_f32=dz32, colsum=g_b,
                                       zero_dy=zero_dy)
                else:
                    bn = st.bn
                    if final_bn_sums is not None and si == len(self.stages) - 1:
                        sums, reps = final_bn_sums, ops.BN_SUM_REPLICAS  # produced by the fused score/loss kernel
                    else:
                        sums, reps = arena.take(2 * st.out_f), 1
                        ops.bn_bwd_reduce(dy32, y, st.act2, st.a32, st.mi, rows, st.out_f, sums)
                    if self.bn_sync is not None:
                        self.bn_sync.reduce_sums(sums)
                    g_gamma, g_beta = grads[id(bn.weight)], grads[id(bn.bias)]
                    if st.act1 is None:
                        # the Linear bias in front of a BatchNorm has an exactly-zero gradient: not computed
                        ops.bn_bwd_apply(dy32, y, st.act2, st.a32, st.mi, bn.weight.detach(), sums, rows, st.out_f,
                                         dz_bf16=dz16, dgamma=g_gamma, dbeta=g_beta, n_replicas=reps)
                    else:
                        da32 = torch.empty((rows, st.out_f), dtype=F32, device=dev)
                        ops.bn_bwd_apply(dy32, y, st.act2, st.a32, st.mi, bn.weight.detach(), sums, rows, st.out_f,
                                         dz_f32=da32, dgamma=g_gamma, dbeta=g_beta, n_replicas=reps)
                        ops.actgrad_colsum(da32, st.a32, st.act1, rows, st.out_f, out_bf16=dz16, colsum=g_b)
            # ---- wgrad: dW[out, in] += dz^T x   (contraction over the rows)
            if first_csr and self._ref is not None:
                # referenced rows only: every stored entry (r, j) adds dz[pos(r)] into row j of the transposed gradient,
                # which is then added into dW [out, in] (and cleared)
                ref = self._ref
                ip, ix = self.feature.csr
                seg = self.feature.csr_seg
                ops.spmm_scatter_wgrad(seg[0] if seg is not None else ip, ix, self.feature.csr_vals, ref["units"],
                                       ref["n_units"], ref["max_units"], seg[1] if seg is not None else None,
                                       ref["pos"], dz16, st.out_f, ref["gT"])
                ops.transpose_add_f32(ref["gT"], g_w)
            elif first_csr:
                ip_t, ix_t = self.feature.csr_t
                # accumulates like every other wgrad of the path (gradient accumulation over micro-batches)
                ops.spmm_csr(ip_t, ix_t, st.in_f, dz16, st.out_f, None, None, g_w, transpose_out=True,
                             vals=self.feature.csr_t_vals, accumulate=True, segments=self.feature.csr_t_seg)
            else:
                x16 = st.x if (si > 0 or self.feature is None or self._first_x is not None) else self.feature.x16
                tiles = -(-st.in_f // 128) * -(-st.out_f // (64 if st.out_f <= 64 else 128 if st.out_f <= 128 else 256))
                split = ops.effective_splits(rows, max(1, min(-(-rows // 64), (splitk_fill() * n_sms) // max(1, tiles))))
                # wgrad partitions accumulate with fp32 atomics (measured faster than private slices + a reduce pass at
                # every shape of the ML-1M step; SBR_SLICED_MIN_SPLIT selects the deterministic sliced variant)
                sliced = split >= int(os.environ.get("SBR_SLICED_MIN_SPLIT", 1 << 30)) and \
                    split * st.out_f * st.in_f * 4 <= ops.SPLITK_MAX_PARTIAL_BYTES
                if sliced:
                    part = torch.empty((split, st.out_f, st.in_f), dtype=F32, device=dev)
                    kw = dict(out_f32=part.view(split * st.out_f, st.in_f), transpose_out=True, split_k=split,
                              split_stride=st.out_f * st.in_f)
                else:
                    kw = dict(out_f32=g_w, transpose_out=True, atomic_out=True, split_k=split)
                def wgrad(st=st, si=si, x16=x16, dz16=dz16, kw=kw, sliced=sliced, split=split, g_w=g_w,
                          part=part if sliced else None):
                    if si == 0 and self.bits_input:
                        # dW^T [in, out] = X^T dZ with X^T = the transposed bit matrix as the K-major A operand
                        ops.gemm_bits(self.feature.bits_t, dz16, st.in_f, st.out_f, rows, b_mn=True, **kw)
                    else:
                        ops.gemm(x16, dz16, st.in_f, st.out_f, rows, a_mn=True, b_mn=True, **kw)
                    if sliced:
                        ops.splitk_reduce(part, split, st.out_f, st.in_f, out_f32=g_w, accumulate=True)
                if wstream is not None and (si > 0 or need_dx):  # (the last GEMM of the chain has nothing to overlap)
                    keep.extend((dz16, x16, part if sliced else None))
                    wstream.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(wstream):
                        wgrad()
                else:
                    wgrad()
            # ---- dgrad
            if si > 0:
                prev = self.stages[si - 1]
                if prev.bn is None:
                    prev_csr = si - 1 == 0 and self.csr_input
                    pad_in = ops.pad8(st.in_f)
                    p16 = torch.zeros((rows, pad_in), dtype=BF16, device=dev) if (prev_csr and pad_in != st.in_f) \
                        else torch.empty((rows, pad_in), dtype=BF16, device=dev)
                    p32 = None
                    g_pb = grads[id(prev.linear.bias)] if prev.linear.bias is not None else None
                    ops.gemm(dz16, st.w16, rows, st.in_f, st.out_f, b_mn=True, out_bf16=p16, out_f32=p32,
                             actgrad_y=prev.y16, actgrad_act=prev.act1, colstats=g_pb, colstats_sum_only=True)
                    fused = (p16, p32)
                else:
                    dy32 = torch.empty((rows, st.in_f), dtype=F32, device=dev)
                    ops.gemm(dz16, st.w16, rows, st.in_f, st.out_f, b_mn=True, out_f32=dy32)
                    zero_dy = False
            elif need_dx:
                dx32 = torch.empty((rows, st.in_f), dtype=F32, device=dev)
                ops.gemm(dz16, st.w16, rows, st.in_f, st.out_f, b_mn=True, out_f32=dx32)
                return dx32
        return None

    def release(self):
        for st in self.stages:
            st.x = st.y16 = st.y32 = st.a32 = st.mi = None


# ================================================================================================ FeatureEmbedding
class FeatureEmbedding(nn.Module):
    """Parameters of one modality (reference ``FeatureEmbedding``, sgd_alg.py:1279-1396): ``pre_embedding_layers``
    (PolyLinear with the activation also on its output) for vector-like features, ``embedding_layer``
    (nn.Embedding / nn.EmbeddingBag(mean, padding_idx=-1)) for categorical / tag features."""

    def __init__(self, feature, embedding_dim: int = None, pre_embedding_layers=None, post_embedding_layers=None,
                 activation_fn: str = "relu"):
        super().__init__()
        self._feature = feature
        self._feature_type = feature_type(feature)
        name = feature.feature_definition.name
        if embedding_dim is None and self._feature_type in ("categorical", "tag"):
            raise ValueError(f'For {self._feature_type} feature "{name}", the size of its embeddings have to be '
                             f'specified with "embedding_dim"')
        if pre_embedding_layers and self._feature_type in ("categorical", "tag"):
            raise ValueError(f'For {self._feature_type} feature "{name}", using pre-embedding layers would not make '
                             f'any sense (as the input are simple indices).')
        self.pre_embedding_layers = None
        self.embedding_layer = None
        self.post_embedding_layers = None
        self.output_dim = embedding_dim
        if self._feature_type == "categorical":
            self.embedding_layer = nn.Embedding(feature.n_unique_categories, embedding_dim)
            nn.init.normal_(self.embedding_layer.weight, std=.1 / embedding_dim)  # train/utils.py:11-13
        elif self._feature_type == "tag":
            # default torch init (general_weight_init skips EmbeddingBag), last row = padding
            self.embedding_layer = nn.EmbeddingBag(feature.dim + 1, embedding_dim, padding_idx=-1)
        else:
            dim = feature.dim if not isinstance(feature.dim, tuple) else int(np.prod(feature.dim))
            cfg = [int(dim)] + list(pre_embedding_layers or []) + ([embedding_dim] if embedding_dim is not None else [])
            self.output_dim = cfg[-1]
            if len(cfg) > 1:
                self.pre_embedding_layers = PolyLinear(cfg, activation_fn=activation_fn, output_fn=activation_fn)
                for m in self.pre_embedding_layers.layers.values():
                    _kaiming_relu_(m)
        if post_embedding_layers:
            cfg = [self.output_dim] + list(post_embedding_layers)
            self.output_dim = cfg[-1]
            self.post_embedding_layers = PolyLinear(cfg, activation_fn=activation_fn, output_fn=activation_fn)
            for m in self.post_embedding_layers.layers.values():
                _kaiming_relu_(m)

    @classmethod
    def build_from_conf(cls, config: FeatureModuleConfig, feature):
        return cls(feature, embedding_dim=config.embedding_dim, pre_embedding_layers=config.pre_embedding_layers,
                   post_embedding_layers=config.post_embedding_layers, activation_fn=config.activation_fn)

    def forward(self, *a, **k):
        raise RuntimeError("FeatureEmbedding is a parameter container; arithmetic runs in the sibrar_b200 kernels")


# ================================================================================================ entity engines
class _EntityBase(nn.Module):
    """shared device-side plumbing of the two entity kinds"""

    def _init_runtime(self, n_entities: int):
        self.n_entities = n_entities
        self._dev_ready = None
        self._owner = None  # set by SingleBranchNet: access to step counter / arena / grads

    def _device(self):
        return next(iter(self.parameters())).device

    def _rt(self):
        if self._owner is None:
            raise RuntimeError("entity module used outside of a SingleBranchNet")
        return self._owner()


def _gather_plan(ent, n_keys: int, N: int, device) -> "ops.GatherPlan":
    """scratch of the sorted-run gather backward, cached per (n_keys, N)"""
    cache = ent.__dict__.setdefault("_plans", {})
    plan = cache.get((n_keys, N))
    if plan is None:
        plan = cache[(n_keys, N)] = ops.GatherPlan(n_keys, N, device)
    return plan


class PlainEntity(_EntityBase):
    """Entity that is a single FeatureEmbedding (reference: ``FeatureEmbedding`` used directly as
    ``{user,item}_embedding_module`` when the config parses as FeatureModuleConfig, sgd_alg.py:2043-2046; forward
    sgd_alg.py:1373-1389): any feature type, optional ``pre_embedding_layers`` (vector-like features) and
    ``post_embedding_layers``.

    A categorical feature without post layers (the shipped configs: ``user_embedding``) is gathered directly.  Every
    other form is a function of the FEATURE ROW alone, so each step computes the table of all rows once
    (``T = post(embed | pre(x))``: the entity-table projection of the single-branch entities) and the batch gathers rows
    of it; the table gradient flows back through the same chain."""

    def __init__(self, feature, config: FeatureModuleConfig, n_entities: int, fe: "FeatureEmbedding" = None):
        """``fe``: an existing FeatureEmbedding (parameter holder) to run instead of building one from ``config`` -- the
        sibling models keep it registered under the reference's own attribute names and run this engine unregistered"""
        nn.Module.__init__(self)
        if fe is None:
            fe = FeatureEmbedding.build_from_conf(config, feature)
        # flatten: the reference's state_dict keys are '<entity>_embedding_module.embedding_layer.weight',
        # '<entity>_embedding_module.{pre,post}_embedding_layers.layers.linear_{i}.{weight,bias}'
        self.embedding_layer = fe.embedding_layer
        self.pre_embedding_layers = fe.pre_embedding_layers
        self.post_embedding_layers = fe.post_embedding_layers
        self._feature = feature
        self._feature_type = fe._feature_type
        self.output_dim = fe.output_dim
        self.k_train = self.k_eval = 1
        self.agg_max = 0
        self.reg_enabled = False
        self.direct = self._feature_type == "categorical" and self.post_embedding_layers is None
        self.normalize_output = False  # L2-normalise the gathered rows (DeepMatrixFactorization's cosine scores)
        self._init_runtime(n_entities)

    def _materialize(self):
        dev = self._device()
        if self._dev_ready == dev:
            return
        self.df = DeviceFeature("plain", self._feature, self.n_entities, dev,
                                dense_min_density=float(os.environ.get("SBR_DENSE_MIN_DENSITY", 0.004)))
        self._dev_ready = dev
        self._srcs = self._srcs_grad = None
        if self.direct:
            return
        df = self.df
        stages = []
        if self.pre_embedding_layers is not None:
            stages += build_stages(self.pre_embedding_layers)
        if self.post_embedding_layers is not None:
            stages += build_stages(self.post_embedding_layers)
        vector = self._feature_type not in ("categorical", "tag")
        self.chain = Chain(stages, feature=df if vector else None) if stages else None
        self.emb_dim = self.embedding_layer.weight.shape[1] if self.embedding_layer is not None else None
        self.table = torch.zeros((df.n_rows, self.output_dim), dtype=F32, device=dev)
        self.table_grad = torch.zeros((df.n_rows, self.output_dim), dtype=F32, device=dev)
        self._rows = torch.arange(df.n_rows, dtype=torch.int64, device=dev)
        if not vector and self.chain is not None:
            self._x32 = torch.zeros((df.n_rows, self.emb_dim), dtype=F32, device=dev)
            self._x16 = torch.zeros((df.n_rows, ops.pad8(self.emb_dim)), dtype=BF16, device=dev)
        self._emb_srcs = {}

    # ---- descriptors
    def _src_blob(self, grads):
        if self.direct:
            w = self.embedding_layer.weight
            g = grads[id(w)] if grads is not None else None
            self.n_keys = int(w.shape[0])
            return ops.make_modality_srcs([dict(kind=SRC_CATEGORICAL, remap=self.df.remap, table=w.detach(), grad=g,
                                                codes=self.df.codes, key_base=0)], w.device)
        self.n_keys = int(self.df.n_rows)
        return ops.make_modality_srcs([dict(kind=SRC_TABLE, remap=self.df.remap, table=self.table, key_base=0,
                                            grad=self.table_grad if grads is not None else None)], self._device())

    def _embedding_src(self, grads):
        """the categorical Embedding addressed by FEATURE ROW (identity remap): the input of the post layers"""
        key = id(grads) if grads is not None else 0
        hit = self._emb_srcs.get(key)
        if hit is None or hit[0] is not grads:
            w = self.embedding_layer.weight
            blob = ops.make_modality_srcs([dict(kind=SRC_CATEGORICAL, remap=None, table=w.detach(), codes=self.df.codes,
                                                grad=grads[id(w)] if grads is not None else None, key_base=0)],
                                          w.device)
            self._emb_srcs[key] = hit = (grads, blob)
        return hit[1]

    # ---- the per-step table of all feature rows
    def _build_table(self, training):
        rt, df = self._rt(), self.df
        if self._feature_type == "tag":
            w = self.embedding_layer.weight.detach()
            bag = self._x32 if self.chain is not None else self.table
            ops.tag_bag_fwd(df.codes, df.max_tags, df.pad_id, w, bag)
            if self.chain is not None:
                ops.cast_bf16(bag, self._x16)
                self.chain.forward(self._x16, df.n_rows, training, rt.arena, keep_for_backward=training, out32=self.table)
        elif self._feature_type == "categorical":
            ops.row_gather_fwd(self._embedding_src(None), 1, self._rows, None, 1, self.emb_dim, False, 0.0, 0,
                               rt.step_dev, None, out_bf16=self._x16, err_flag=rt.err_flag)
            self.chain.forward(self._x16, df.n_rows, training, rt.arena, keep_for_backward=training, out32=self.table)
        else:
            self.chain.forward(df.x16, df.n_rows, training, rt.arena, keep_for_backward=training, out32=self.table)

    def embed(self, idx, training, mods=None, keep_mask=None, defer_final_bn=False, out=None):
        """``out``: optional fp32 [numel, D] destination (may be a column block of a wider buffer: row pitch = stride)"""
        self._materialize()
        rt = self._rt()
        flat = idx.reshape(-1).contiguous()
        D = self.output_dim
        if not self.direct:
            self._build_table(training)
        if out is None:
            out = torch.empty((flat.numel(), D), dtype=F32, device=flat.device)
        if self._srcs is None:
            self._srcs = self._src_blob(None)
        ops.row_gather_fwd(self._srcs, 1, flat, None, 1, D, self.normalize_output, 0.0, 0, rt.step_dev, None,
                           out_f32=out, err_flag=rt.err_flag)
        self._ctx = (flat,)
        return out

    def build_plan(self, idx, mods, k, grads):
        pass  # (single source: the plan is built in backward)

    def chains(self):
        self._materialize()
        return [] if self.direct or self.chain is None else [self.chain]

    def backward(self, dE, grads, final_bn_sums=None):
        rt = self._rt()
        (flat,) = self._ctx
        if self._srcs_grad is None or self._srcs_grad[0] is not grads:
            self._srcs_grad = (grads, self._src_blob(grads))
        plan = _gather_plan(self, self.n_keys, flat.numel(), flat.device)
        plan.build(self._srcs_grad[1], 1, flat, None, 1)
        plan.backward(self._srcs_grad[1], 1, self.output_dim, self.normalize_output, 0.0, 0, rt.step_dev, None, dE)
        if self.direct:
            return
        # table-level backward: G = d loss / d T  (the consumers clear it)
        df = self.df
        if self.chain is None:  # tag feature without post layers: T is the bag table itself
            self._tag_backward(self.table_grad, grads)
            return
        vector = self._feature_type not in ("categorical", "tag")
        dx = self.chain.backward(self.table_grad, grads, need_dx=not vector, arena=rt.arena, zero_dy=True)
        if vector:
            return
        if self._feature_type == "tag":
            self._tag_backward(dx, grads)
        else:
            p2 = _gather_plan(self, int(self.embedding_layer.weight.shape[0]), df.n_rows, dx.device)
            src = self._embedding_src(grads)
            p2.build(src, 1, self._rows, None, 1)
            p2.backward(src, 1, self.emb_dim, False, 0.0, 0, rt.step_dev, None, dx)

    def _tag_backward(self, bag_grad, grads):
        df, w = self.df, self.embedding_layer.weight
        C_ = int(w.shape[1])
        if w.numel() <= ops.TAG_BAG_SMEM_FLOATS or C_ % 4 != 0:
            ops.tag_bag_bwd(df.codes, df.max_tags, df.pad_id, bag_grad, grads[id(w)])
        else:
            seg_ptr, seg_rows, seg_vals, seg_tag = df.tag_segments
            ops.spmm_csr(seg_ptr, seg_rows, seg_tag.numel(), bag_grad, C_, None, None, grads[id(w)], vals=seg_vals,
                         row_map=seg_tag, atomic=True)
            bag_grad.zero_()


class SingleBranchNetEntity(_EntityBase):
    """reference ``SingleBranchNetEntity`` (sgd_alg.py:1764-2006)"""

    def __init__(self, entity_name: str, features: dict, entity_config: SingleBranchNetEntityConfig,
                 shared_common_dim: int, val_interactions_available: bool = True, n_entities: int = None):
        nn.Module.__init__(self)
        self.features = features
        self.entity_name = entity_name
        self.entity_config = entity_config
        self.output_dim = shared_common_dim
        self.val_interactions_available = val_interactions_available
        cfg = entity_config
        if len(cfg.features) == 0:
            raise ValueError("SingleBranchEntity requires at least one feature.")
        self.train_modalities = self._get_modalities(train=True)
        self.eval_modalities = self._get_modalities(train=False)
        missing = self.train_modalities - set(features.keys())
        if missing:
            raise ValueError(f"Features for modalities {missing} are not available!")
        missing = self.train_modalities - {f.feature_name for f in cfg.features}
        if missing:
            raise ValueError(f"Network definitions for modalities {missing} are not available!")
        if cfg.aggregation_fn not in ("mean", "max"):
            raise ValueError(f'Aggregation function "{cfg.aggregation_fn}" is not supported.')
        if not isinstance(cfg.embedding_regularization_type, EmbeddingRegularizationType):
            raise ValueError(f'Embedding regularization "{cfg.embedding_regularization_type}" is not yet supported.')

        self.modality_modules = nn.ModuleDict()
        for f in cfg.features:
            if f.feature_name not in self.train_modalities:
                continue
            self.modality_modules[f.feature_name] = FeatureEmbedding(
                features[f.feature_name], embedding_dim=cfg.common_modality_dim,
                pre_embedding_layers=f.feature_hidden_layers, activation_fn=cfg.activation_fn)
        self.mod_names = list(self.modality_modules.keys())  # canonical modality ids of this entity

        layers = []
        if cfg.single_branch_input_dropout is not None:
            layers.append(_Identity())  # nn.Dropout in the reference: no parameters, shifts the numbering
        every = cfg.apply_batch_norm_every if cfg.apply_batch_normalization else 0
        poly = PolyLinear([cfg.common_modality_dim] + list(cfg.single_branch_hidden_layers) + [self.output_dim],
                          activation_fn=cfg.activation_fn,
                          output_fn=cfg.activation_fn if cfg.apply_output_activation else None,
                          apply_batch_norm_every=every)
        layers.append(poly)
        trailing_bn = None
        if cfg.apply_batch_normalization and cfg.apply_batch_norm_every == 0:
            trailing_bn = nn.BatchNorm1d(self.output_dim)
            layers.append(trailing_bn)
        self.sb_net = nn.Sequential(*layers)
        # plain references (not registered a second time as sub-modules)
        object.__setattr__(self, "_poly", poly)
        object.__setattr__(self, "_trailing_bn", trailing_bn)

        self.reg_type = cfg.embedding_regularization_type
        self.reg_enabled = self.reg_type != EmbeddingRegularizationType.NoRegularization
        if self.reg_type == EmbeddingRegularizationType.CentralModality and \
                cfg.central_modality not in self.train_modalities:
            raise ValueError(f'central item "{cfg.central_modality}" must be contained in "a"')
        if self.reg_enabled and len(self.mod_names) < 2:
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")
        self.k_train = 2 if self.reg_enabled else 1
        self.k_eval = len(self.eval_modalities)
        self.agg_max = 1 if cfg.aggregation_fn == "max" else 0
        self.regularization_loss = None
        self._init_runtime(n_entities)

    # ---- modality sets (sgd_alg.py:1879-1902)
    def _get_modalities(self, train=True):
        cfg = self.entity_config
        available = {f.feature_name for f in cfg.features}
        if train:
            mods = set(cfg.train_modalities or available)
        else:
            train_mods = self._get_modalities(train=True)
            if cfg.eval_modalities is not None:
                for m in cfg.eval_modalities:
                    if m not in train_mods:
                        raise ValueError(f'Cannot use modality "{m}" during evaluation, if it is not used during '
                                         f'training.')
            mods = set(cfg.eval_modalities or train_mods)
            if not self.val_interactions_available:
                mods.discard("interactions")
        if len(mods) == 0:
            raise ValueError(f'No single modality is available during {"training" if train else "evaluation"}: '
                             f'There are either no modalities specified or no interactions are available)')
        return mods

    # ---- device state
    def _materialize(self):
        dev = self._device()
        if self._dev_ready == dev:
            return
        cfg = self.entity_config
        self.dfeat, self.proj = {}, {}
        for name, fe in self.modality_modules.items():
            df = DeviceFeature(name, self.features[name], self.n_entities, dev,
                               dense_min_density=float(os.environ.get("SBR_DENSE_MIN_DENSITY", 0.004)))
            self.dfeat[name] = df
            if fe.pre_embedding_layers is not None:
                self.proj[name] = Chain(build_stages(fe.pre_embedding_layers), feature=df)
        self.sb_chain = Chain(build_stages(self._poly, self._trailing_bn))
        # tag features: the EmbeddingBag means of all feature rows are a per-step table as well (sbr_tag_bag_fwd/bwd),
        # so the batch-sized gathers see one source row per lookup
        self.bags = [n for n in self.mod_names if self.dfeat[n].kind == "tag"]
        # fp32 [n_rows, C] projected modality tables at stable addresses (the gather descriptors point at them)
        self.tables = {n: torch.zeros((self.dfeat[n].n_rows, cfg.common_modality_dim), dtype=F32, device=dev)
                       for n in list(self.proj) + self.bags}
        self.table_grads = {n: torch.zeros((self.dfeat[n].n_rows, cfg.common_modality_dim), dtype=F32, device=dev)
                            for n in list(self.proj) + self.bags}
        self._srcs_cache = {}
        # "referenced rows" route (csrc/refrows.cu): per modality projection, the per-step subset of feature rows
        self._ref_state = {}    # name -> dict(stamp, pos, gT)
        self._ref_bufs = {}     # (name, capacity) -> per-shape buffers
        self._ref_epoch = torch.zeros(2, dtype=torch.int64, device=dev)  # [epoch, finished blocks] of the marking kernel
        self._route = {}        # modality name -> buffers of the current call (empty: whole-table route everywhere)
        self._eval_ids = torch.tensor([self.mod_names.index(m) for m in sorted(self.eval_modalities)],
                                      dtype=torch.uint8, device=dev)
        self._dev_ready = dev

    def _src_blob(self, grads):
        route_sig = tuple(sorted((n, b["capacity"]) for n, b in self._route.items()))
        key = (id(grads) if grads is not None else 0, route_sig)
        hit = self._srcs_cache.get(key)
        if hit is not None and hit[0] is grads:
            self.n_keys = hit[3]
            return hit[1]
        entries, key_base = [], 0
        for name in self.mod_names:
            df, fe = self.dfeat[name], self.modality_modules[name]
            if name in self._route:
                # projected rows of THIS step, addressed indirectly: feature row -> compact position (like a categorical
                # source whose codes are the positions handed out by sbr_mark_referenced)
                b = self._route[name]
                entries.append(dict(kind=SRC_CATEGORICAL, remap=df.remap, table=b["T"], codes=b["pos"],
                                    key_base=key_base, grad=b["G"] if grads is not None else None))
                key_base += int(b["capacity"])
            elif name in self.tables:
                entries.append(dict(kind=SRC_TABLE, remap=df.remap, table=self.tables[name], key_base=key_base,
                                    grad=self.table_grads[name] if grads is not None else None))
                key_base += int(df.n_rows)
            else:
                w = fe.embedding_layer.weight
                g = grads[id(w)] if grads is not None else None
                kind = SRC_CATEGORICAL if df.kind == "categorical" else SRC_TAG
                entries.append(dict(kind=kind, remap=df.remap, table=w.detach(), grad=g, codes=df.codes,
                                    max_tags=df.max_tags, pad_id=df.pad_id, key_base=key_base))
                # segment key: the category (Embedding) or the entity's feature row (EmbeddingBag)
                key_base += int(w.shape[0]) if kind == SRC_CATEGORICAL else int(df.n_rows)
        self.n_keys = key_base
        blob = ops.make_modality_srcs(entries, self._device())
        self._srcs_cache[key] = (grads, blob, entries, key_base)
        return blob

    # ---- referenced-rows route
    def _select_route(self, n_idx: int):
        """which modality projections run on the rows this call references only (static in the call's shape): vector /
        sparse features whose table has at least twice as many rows as the call has entities"""
        self._route = {}
        if os.environ.get("SBR_REF_ROWS", "1") == "0":
            return
        dev = self._device()
        C_ = self.entity_config.common_modality_dim
        for name, chain in self.proj.items():
            df = self.dfeat[name]
            if df.kind not in ("dense", "csr") or 2 * n_idx > df.n_rows:
                continue
            st = self._ref_state.get(name)
            if st is None:
                st = self._ref_state[name] = dict(
                    stamp=torch.zeros(df.n_rows, dtype=torch.int32, device=dev),
                    pos=torch.zeros(df.n_rows, dtype=torch.int32, device=dev),
                    gT=torch.zeros((chain.stages[0].in_f, chain.stages[0].out_f), dtype=F32, device=dev)
                    if df.kind == "csr" else None)
            cap = int(n_idx)
            b = self._ref_bufs.get((name, cap))
            if b is None:
                seg = df.csr_seg if df.kind == "csr" else None
                extra = int(seg[1].numel() - df.n_rows) if seg is not None else 0
                b = self._ref_bufs[(name, cap)] = dict(
                    capacity=cap, pos=st["pos"], stamp=st["stamp"], gT=st["gT"],
                    list=torch.zeros(cap, dtype=torch.int32, device=dev),
                    T=torch.zeros((cap, C_), dtype=F32, device=dev), G=torch.zeros((cap, C_), dtype=F32, device=dev),
                    X=torch.zeros((cap, df.x16.shape[1]), dtype=BF16, device=dev) if df.kind == "dense" else None,
                    seg_first=seg[3] if seg is not None else None, max_units=cap + extra,
                    seg_list=torch.zeros(cap + extra, dtype=torch.int32, device=dev) if seg is not None else None)
            self._route[name] = b

    def begin_call(self, flat, mods, k):
        """route selection + row marking of one embed call.  The trainer issues it before it forks the side stream (the
        gather plan of this entity is built there and needs the compact positions); ``embed`` does it otherwise."""
        self._select_route(flat.numel())
        self._mark_referenced(flat, mods, k)

    def _mark_referenced(self, flat, mods, k):
        """one pass over the call's (entity, modality) slots: row lists + compact positions of the routed modalities"""
        if not self._route:
            return
        rt = self._rt()
        entries = []
        for name in self.mod_names:
            b = self._route.get(name)
            if b is None:
                entries.append(None)
                continue
            cnt = rt.arena.take(2, torch.int32)  # [rows, segments] of this call (the arena is cleared per step)
            b["count"], b["seg_count"] = cnt[0:1], cnt[1:2]
            entries.append(dict(stamp=b["stamp"], pos=b["pos"], list=b["list"], count=b["count"],
                                seg_first=b["seg_first"], seg_list=b["seg_list"], seg_count=b["seg_count"]))
        sig = tuple(None if e is None else (e["list"].data_ptr(), e["count"].data_ptr()) for e in entries)
        hit = self.__dict__.get("_ref_tabs")
        if hit is None or hit[0] != sig:
            self._ref_tabs = hit = (sig, ops.make_ref_tables(entries, self._device()), entries)
        ops.mark_referenced(self._src_blob(None), len(self.mod_names), flat, mods, k, self._ref_epoch, hit[1])

    def _aux_streams(self, n):
        """side streams of this entity (None unless the trainer enabled branch parallelism on the runtime)"""
        if not getattr(self._rt(), "branches", False):
            return None
        pool = self.__dict__.setdefault("_aux_pool", [])
        while len(pool) < n:
            pool.append(torch.cuda.Stream(device=self._device()))
        return pool[:n]

    def _project_tables(self, names, training, arena):
        """entity-table projection: T_m = PolyLinear_m(X_m) for ALL rows of every listed modality (independent of
        each other: parallel branches)"""
        def one(name):
            df = self.dfeat[name]
            if name in self.bags:
                w = self.modality_modules[name].embedding_layer.weight.detach()
                ops.tag_bag_fwd(df.codes, df.max_tags, df.pad_id, w, self.tables[name])
                return
            chain = self.proj[name]
            b = self._route.get(name)
            if b is None:
                chain.forward(df.x16, df.n_rows, training, arena, keep_for_backward=training, out32=self.tables[name])
                return
            seg = b["seg_list"] is not None
            ref = dict(list=b["list"], count=b["count"], pos=b["pos"], gT=b["gT"], max_units=b["max_units"],
                       units=b["seg_list"] if seg else b["list"], n_units=b["seg_count"] if seg else b["count"])
            if df.kind == "dense":
                ops.gather_rows_bf16(df.x16, b["list"], b["count"], b["capacity"], b["X"])
            chain.forward(b["X"], b["capacity"], training, arena, keep_for_backward=training, out32=b["T"], ref=ref)
        todo = [n for n in names if n in self.tables]
        run_branches([lambda n=n: one(n) for n in todo], self._aux_streams(len(todo) - 1) if training else None)

    def sample_modalities(self, n_idx: int):
        """device replacement of ``_sample_modalities`` (sgd_alg.py:1904-1927) for training"""
        rt = self._rt()
        k = self.k_train
        mods = torch.empty(n_idx * k, dtype=torch.uint8, device=self._device())
        central = -1
        if self.reg_type == EmbeddingRegularizationType.CentralModality:
            central = self.mod_names.index(self.entity_config.central_modality)
        seed = (int(self.entity_config.sampling_seed) << 8) ^ (1 if self.entity_name == "user" else 2) ^ \
            (int(rt.seed_salt) << 24)
        ops.sample_modalities(mods, n_idx, k, len(self.mod_names), central, seed, rt.step_dev)
        return mods

    def embed(self, idx, training, mods=None, keep_mask=None, defer_final_bn=False):
        """indices [...] -> per-slot embeddings fp32 [N = numel * k, D] (``_embed``, sgd_alg.py:1865-1877).
        With ``defer_final_bn`` (fused trainer) the trailing BatchNorm is left to the score/loss kernel: returns None
        and ``self.sb_chain.deferred`` describes it."""
        self._materialize()
        rt = self._rt()
        cfg = self.entity_config
        dev = self._device()
        flat = idx.reshape(-1).contiguous()
        n_idx = flat.numel()
        if training:
            k = self.k_train
            if mods is None:
                mods = self.sample_modalities(n_idx)
            used = self.mod_names
        else:
            k = self.k_eval
            mods = self._eval_ids.repeat(n_idx)
            used = sorted(self.eval_modalities)
        if not self.__dict__.pop("_premarked", False):
            self.begin_call(flat, mods, k)
        self._project_tables(used, training, rt.arena)
        srcs = self._src_blob(None)
        C_ = cfg.common_modality_dim
        N = n_idx * k
        p_drop = cfg.single_branch_input_dropout if (training and cfg.single_branch_input_dropout) else 0.0
        seed = (int(cfg.sampling_seed) << 8) ^ (0x11 if self.entity_name == "user" else 0x22) ^ \
            (int(rt.seed_salt) << 24)
        self._fused = None
        if training and self._fused_mlp_ok():
            self._ctx = (flat, mods, keep_mask, k, p_drop, seed, None)
            return self._embed_fused(srcs, flat, mods, keep_mask, k, N, p_drop, seed, defer_final_bn and k == 1)
        x0 = torch.empty((N, ops.pad8(C_)), dtype=BF16, device=dev)
        keep_bits = torch.empty((N, (C_ + 7) // 8), dtype=torch.uint8, device=dev) if (training and p_drop) else None
        ops.row_gather_fwd(srcs, len(self.mod_names), flat, mods, k, C_, cfg.normalize_single_branch_input, p_drop,
                           seed, rt.step_dev, keep_mask, out_bf16=x0, err_flag=rt.err_flag, keep_bits_out=keep_bits)
        E = self.sb_chain.forward(x0, N, training, rt.arena, keep_for_backward=training,
                                  defer_final_bn=defer_final_bn and k == 1)
        self._ctx = (flat, mods, keep_mask, k, p_drop, seed, keep_bits)
        return E

    # ---- fused gather + single-branch MLP (csrc/mlp_fused.cu): chains of 1 or 2 Linear layers, all widths <= 64
    def _fused_mlp_ok(self):
        if os.environ.get("SBR_FUSED_MLP", "1") == "0":
            return False
        st = self.sb_chain.stages
        if len(st) not in (1, 2) or self.entity_config.common_modality_dim > 64 or len(self.mod_names) > 16:
            return False
        if any(s.in_f > 64 or s.out_f > 64 for s in st):
            return False
        if any(s.bn is not None or s.act2 is not None for s in st[:-1]) or st[-1].act2 is not None:
            return False  # (BatchNorm between the layers needs all rows between two GEMMs: layer-by-layer path)
        return True

    def _embed_fused(self, srcs, flat, mods, keep_mask, k, N, p_drop, seed, defer_bn):
        rt = self._rt()
        cfg = self.entity_config
        dev = self._device()
        st = self.sb_chain.stages
        last = st[-1]
        C_, D = cfg.common_modality_dim, last.out_f
        layers = []
        for s_ in st:
            s_.refresh(False)
            layers.append((s_.w16, s_.linear.bias.detach() if s_.linear.bias is not None else None, s_.in_f, s_.out_f,
                           s_.act1))
        desc = ops.mlp2_desc(srcs, len(self.mod_names), flat, mods, k, C_, cfg.normalize_single_branch_input, p_drop,
                             seed, rt.step_dev, keep_mask, rt.err_flag, layers)
        z = torch.empty((N, D), dtype=F32, device=dev)
        stats, n_part, mi = None, 0, None
        if last.bn is not None:
            n_part = ops.mlp2_colstats_rows(N)
            stats = torch.empty((n_part, 2 * D), dtype=F32, device=dev)
        sync = self.sb_chain.bn_sync if last.bn is not None else None
        tail = None
        if last.bn is not None and sync is None and os.environ.get("SBR_MLP2_BN_TAIL", "0") == "1":
            # BatchNorm statistics finalised by the last CTA of the forward kernel (no sbr_bn_finalize launch).  Opt-in:
            # measured on the ML-1M step the serial tail of the last CTA (4 us) costs what the PDL-overlapped finalize
            # launch did (0.295-0.299 ms with it, 0.293 ms without)
            if getattr(self, "_bn_ticket", None) is None or self._bn_ticket.device != dev:
                self._bn_ticket = torch.zeros(1, dtype=torch.int32, device=dev)
            mi = torch.empty(2 * D, dtype=F32, device=dev)
            tail = dict(counter=self._bn_ticket, eps=last.bn.eps, momentum=last.bn.momentum, mean_invstd=mi,
                        running_mean=last.bn.running_mean, running_var=last.bn.running_var,
                        num_batches_tracked=last.bn.num_batches_tracked)
        ops.mlp2_fwd(desc, N, C_, z, stats, n_part, bn_tail=tail)
        E = z
        self.sb_chain.deferred = None
        if last.bn is not None:
            bn = last.bn
            if tail is None:
                mi = torch.empty(2 * D, dtype=F32, device=dev)
                rows_g = N
                if sync is not None:
                    stats, n_part, rows_g = sync.gather_stats(stats), n_part * sync.world, N * sync.world
                ops.bn_finalize(stats, rows_g, D, mi, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                eps=bn.eps, momentum=bn.momentum, n_partials=n_part)
            if defer_bn:
                self.sb_chain.deferred = dict(z=z, mean_invstd=mi, gamma=bn.weight.detach(), beta=bn.bias.detach())
                E = None
            else:
                E = torch.empty((N, D), dtype=F32, device=dev)
                ops.bn_apply(z, mi, bn.weight.detach(), bn.bias.detach(), None, N, D, out_f32=E)
        self._fused = dict(desc=desc, z=z, mi=mi, N=N)
        return E

    def _backward_fused(self, dE, grads, final_bn_sums):
        rt = self._rt()
        f = self._fused
        st = self.sb_chain.stages
        last = st[-1]
        C_, D, N = self.entity_config.common_modality_dim, last.out_f, f["N"]
        bn = None
        if last.bn is not None:
            if final_bn_sums is not None:
                sums, reps = final_bn_sums, ops.BN_SUM_REPLICAS  # produced by the fused score/loss kernel
            else:
                sums, reps = rt.arena.take(2 * D), 1
                ops.bn_bwd_reduce(dE, None, None, f["z"], f["mi"], N, D, sums)
            if self.sb_chain.bn_sync is not None:
                self.sb_chain.bn_sync.reduce_sums(sums)
            bn = dict(mean_invstd=f["mi"], gamma=last.bn.weight.detach(), sums=sums, n_replicas=reps,
                      dgamma=grads[id(last.bn.weight)], dbeta=grads[id(last.bn.bias)])
        gw = [grads[id(s_.linear.weight)] for s_ in st]
        gb = [grads[id(s_.linear.bias)] if s_.linear.bias is not None else None for s_ in st]
        if last.bn is not None and last.act1 is None:
            gb[-1] = None  # the Linear bias in front of a BatchNorm has an exactly-zero gradient: not computed
        dx0 = torch.empty((N, C_), dtype=F32, device=dE.device)
        ops.mlp2_bwd(f["desc"], N, C_, dE, f["z"], bn, gw, gb, dx0)
        return dx0

    def dropout_keep_bits(self):
        """debugging / parity tests: the dropout keep decisions of the last training forward (uint8 [N, ceil(C / 8)],
        bit j of byte c / 8 = column c), regenerated from the counter-based hash (seed, step, row, column group)"""
        rt = self._rt()
        flat, mods, keep_mask, k, p_drop, seed, keep_bits = self._ctx
        if keep_bits is not None:
            return keep_bits
        C_ = self.entity_config.common_modality_dim
        N = flat.numel() * k
        bits = torch.empty((N, (C_ + 7) // 8), dtype=torch.uint8, device=flat.device)
        tmp = torch.empty((N, ops.pad8(C_)), dtype=BF16, device=flat.device)
        ops.row_gather_fwd(self._src_blob(None), len(self.mod_names), flat, mods, k, C_,
                           self.entity_config.normalize_single_branch_input, p_drop, seed, rt.step_dev, keep_mask,
                           out_bf16=tmp, keep_bits_out=bits)
        return bits

    def build_plan(self, idx, mods, k, grads):
        """sorted-run plan of the gather backward; depends only on (indices, sampled modalities), so the trainer
        builds it on its side stream while the forward GEMMs run"""
        flat = idx.reshape(-1)
        self._select_route(flat.numel())
        srcs = self._src_blob(grads)
        plan = _gather_plan(self, self.n_keys, flat.numel() * k, flat.device)
        plan.build(srcs, len(self.mod_names), flat, mods, k)
        self._plan_built = True
        return plan

    def backward(self, dE, grads, final_bn_sums=None):
        rt = self._rt()
        cfg = self.entity_config
        flat, mods, keep_mask, k, p_drop, seed, keep_bits = self._ctx
        C_ = cfg.common_modality_dim
        aux = self._aux_streams(max(1, len(self.tables) - 1))
        if getattr(self, "_fused", None) is not None:
            dx0 = self._backward_fused(dE, grads, final_bn_sums)
        else:
            dx0 = self.sb_chain.backward(dE, grads, need_dx=True, arena=rt.arena, final_bn_sums=final_bn_sums,
                                         wgrad_stream=aux[0] if aux else None)
        srcs = self._src_blob(grads)
        if getattr(self, "_plan_built", False):
            plan = _gather_plan(self, self.n_keys, flat.numel() * k, flat.device)
        else:
            plan = self.build_plan(flat, mods, k, grads)
        self._plan_built = False
        plan.backward(srcs, len(self.mod_names), C_, cfg.normalize_single_branch_input, p_drop, seed, rt.step_dev,
                      keep_mask, dx0, keep_bits=keep_bits)
        # table-level backward, one independent branch per modality; the accumulator table is cleared by the kernel
        # that consumes it
        thunks = [lambda n=n, c=c: c.backward(self._route[n]["G"] if n in self._route else self.table_grads[n], grads,
                                              need_dx=False, arena=rt.arena, zero_dy=True)
                  for n, c in self.proj.items()]
        for n in self.bags:
            df, w = self.dfeat[n], self.modality_modules[n].embedding_layer.weight

            def bag_bwd(n=n, df=df, w=w):
                if w.numel() <= ops.TAG_BAG_SMEM_FLOATS or C_ % 4 != 0:
                    # small vocabulary: block-private shared-memory copies of the whole gradient matrix
                    ops.tag_bag_bwd(df.codes, df.max_tags, df.pad_id, self.table_grads[n], grads[id(w)])
                else:
                    seg_ptr, seg_rows, seg_vals, seg_tag = df.tag_segments
                    ops.spmm_csr(seg_ptr, seg_rows, seg_tag.numel(), self.table_grads[n], C_, None, None, grads[id(w)],
                                 vals=seg_vals, row_map=seg_tag, atomic=True)
                    self.table_grads[n].zero_()  # (memset: the accumulator table is cleared by its consumer)
            thunks.append(bag_bwd)
        if len(thunks) > 2 and os.environ.get("SBR_TAIL_BRANCHES", "small_first") == "small_first":
            # the short chains back to back on the current stream, the longest one (first modality: 'interactions', a
            # 290-CTA GEMM that fills the machine) on a branch: measured 0.292-0.295 vs 0.297 ms per ML-1M step, 0.176 vs
            # 0.185 ms at B = 256, against one branch per modality (scripts/step_timeline.py shows the small kernels of
            # separate branches starting only behind the big GEMMs either way)
            rest = thunks[1:]
            thunks = [lambda rest=rest: [t() for t in rest], thunks[0]]
        run_branches(thunks, aux)

    def chains(self):
        """every Linear chain of the entity (bf16 weight shadows are maintained per stage)"""
        self._materialize()
        return list(self.proj.values()) + [self.sb_chain]

    def get_and_reset_other_loss(self) -> Dict:
        loss = self.regularization_loss
        if loss is None:
            loss = torch.zeros(1, device=self._device())
        self.regularization_loss = None
        return {"reg_loss": loss * self.entity_config.regularization_weight}


# ================================================================================================ the model
class _Runtime:
    def __init__(self, device):
        self.device = device
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=device)
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.arena = ZeroArena(device)
        self.branches = False  # FusedTrainer: independent kernel chains of a step on parallel streams
        self.seed_salt = 0     # data parallel: the rank, so that ranks draw different modalities / dropout masks


class SingleBranchNet(nn.Module):
    """drop-in for the reference ``SingleBranchNet``"""

    def __init__(self, config: SingleBranchNetConfig, dataset):
        super().__init__()
        self.name = self.__class__
        from .synthetic import SynFeature  # light Feature-like container (no arithmetic)
        user_features = dataset.user_features
        user_features["interactions"] = SynFeature("interactions", "vector", dataset.user_sampling_matrix_train)
        user_features["user_embedding"] = SynFeature("user_embedding", "categorical", np.arange(dataset.n_users))
        item_features = dataset.item_features
        item_features["interactions"] = SynFeature("interactions", "vector", dataset.item_sampling_matrix_train)
        item_features["item_embedding"] = SynFeature("item_embedding", "categorical", np.arange(dataset.n_items))
        self.config = config
        self.n_users, self.n_items = dataset.n_users, dataset.n_items
        self.is_user_sb_module = config.is_user_sb_module
        self.is_item_sb_module = config.is_item_sb_module
        D = config.shared_common_dim
        self.user_embedding_module = self._build_entity("user", config.user, user_features, D,
                                                        not dataset.is_cold_start_user, dataset.n_users)
        self.item_embedding_module = self._build_entity("item", config.item, item_features, D,
                                                        not dataset.is_cold_start_item, dataset.n_items)
        self._runtime = None
        import weakref
        ref = weakref.ref(self)
        for ent in (self.user_embedding_module, self.item_embedding_module):
            ent._owner = lambda r=ref: r()._rt()

    @staticmethod
    def _build_entity(name, conf, features, D, interactions_available, n_entities):
        if isinstance(conf, SingleBranchNetEntityConfig):
            return SingleBranchNetEntity(name, features, conf, D, interactions_available, n_entities)
        if conf.embedding_dim == -1:
            conf.embedding_dim = D
        ent = PlainEntity(features[conf.feature_name], conf, n_entities)
        if ent.output_dim != D:
            # (the reference fails later, inside the einsum of combine_user_item_representations, sgd_alg.py:2109-2114)
            raise ValueError(f'plain "{name}" entity produces {ent.output_dim}-d representations, the model\'s '
                             f'shared_common_dim is {D}')
        return ent

    @staticmethod
    def build_from_conf(conf: dict, dataset):
        return SingleBranchNet(SingleBranchNetConfig.from_dict(conf), dataset)

    # ---- runtime
    @property
    def device(self):
        return next(iter(self.parameters())).device

    def _rt(self) -> _Runtime:
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("sibrar_b200.SingleBranchNet runs on a CUDA device (sm_100a) only -- there is no CPU "
                               "fallback; move the model with .to('cuda')")
        if self._runtime is None or self._runtime.device != dev:
            self._runtime = _Runtime(dev)
        return self._runtime

    def refresh_shadows(self):
        """bring the bf16 weight shadows up to date with the fp32 masters (host-side version check, cast kernels only
        for weights that changed); CUDA-graph replays of the evaluation call this first"""
        for ent in (self.user_embedding_module, self.item_embedding_module):
            for chain in ent.chains():
                for st in chain.stages:
                    st.refresh(False)

    def check_errors(self):
        """raises KeyError like ``Feature.__getitem__`` (data/Feature.py:146) if a kernel met an entity index
        without a feature row (host sync: call outside the hot loop)"""
        if self._runtime is not None and int(self._runtime.err_flag.item()) != 0:
            self._runtime.err_flag.zero_()
            raise KeyError("an entity index without a feature row was requested")

    # ---- reference API (model(u, i) is differentiable in training mode; FusedTrainer is the fast path)
    def _represent(self, ent, idx):
        """representations WITHOUT an autograd graph (evaluation, analysis); gradients flow through ``forward`` (the
        reference loop) or ``FusedTrainer.step``"""
        training = self.training
        rt = self._rt()
        rt.arena.reset()
        E = ent.embed(idx, training)
        k = ent.k_train if training else ent.k_eval
        D = self.config.shared_common_dim
        if k == 1:
            return E.view(*idx.shape, D)
        out = torch.empty((idx.numel(), D), dtype=F32, device=E.device)
        ops.aggregate(E, idx.numel(), k, D, ent.agg_max, out_f32=out)
        return out.view(*idx.shape, D)

    def get_user_representations(self, u_idxs: torch.Tensor):
        return self._represent(self.user_embedding_module, u_idxs)

    def get_item_representations(self, i_idxs: torch.Tensor):
        return self._represent(self.item_embedding_module, i_idxs)

    def combine_user_item_representations(self, u_repr, i_repr):
        from .autograd import combine
        return combine(u_repr, i_repr)

    def forward(self, u_idxs, i_idxs):
        if self.training and torch.is_grad_enabled():
            from .autograd import train_forward
            return train_forward(self, u_idxs, i_idxs)  # differentiable: loss.backward() fills param.grad
        u_repr = self.get_user_representations(u_idxs)
        i_repr = self.get_item_representations(i_idxs)
        return self.combine_user_item_representations(u_repr, i_repr)

    @torch.no_grad()
    def predict(self, u_idxs, i_idxs):
        self.eval()
        return self(u_idxs, i_idxs)

    def get_and_reset_other_loss(self) -> Dict:
        losses = {"reg_loss": torch.zeros(1, device=self.device)}
        for name, ent, sb in (("user", self.user_embedding_module, self.is_user_sb_module),
                              ("item", self.item_embedding_module, self.is_item_sb_module)):
            if sb:
                r = ent.get_and_reset_other_loss()
                losses["reg_loss"] = losses["reg_loss"] + r["reg_loss"]
                losses.update({f"{name}_{k}": v for k, v in r.items()})
        return losses

    def save_model_to_path(self, path: str):
        torch.save(self.state_dict(), os.path.join(path, "model.pth"))
        print("Model Saved")

    def load_model_from_path(self, path: str):
        self.load_state_dict(torch.load(os.path.join(path, "model.pth"), map_location=self.device))
        print("Model Loaded")
