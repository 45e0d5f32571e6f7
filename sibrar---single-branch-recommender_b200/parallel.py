"""One-process-per-GPU scaling of the hot path over NCCL / NVLink (``torch.distributed`` is plumbing only).

The reference is single-device (no collective anywhere); this is the B200 addition:
  * training: data parallel.  Every rank runs the fused step on its own B interactions; the flat fp32 gradient buffer
    lives in NVSwitch multicast (symmetric) memory and the all-reduce is the prologue of the optimizer kernel
    (``sbr_adam_step_mc``: ``multimem.ld_reduce`` / ``multimem.st``, flag barriers over peer memory, 1/world folded into
    the update) -- no collective node in the step.  Fallback (``SBR_DP_COLLECTIVE=nccl`` or no multicast): two NCCL
    all-reduce buckets, the item-entity bucket overlapping the user-entity backward.  BatchNorm statistics are rank-local
    unless ``sync_bn=True``; the user-side in-batch InfoNCE stays rank-local.
  * evaluation: the item catalogue is sharded; each rank computes its items' representations and an exact local
    top-k (packed keys with GLOBAL positions), the [U, k] key lists are all-gathered and merged by
    ``sbr_topk_merge`` -- the only exchange step of the path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .evaluator import FullEvaluator
from .feature_store import csr_to_device
from .trainer import FusedTrainer


def shard_range(n: int, rank: int, world: int):
    """contiguous shard [lo, hi) of n items; shards differ by at most one element"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_keys(scores: np.ndarray, positions: np.ndarray) -> np.ndarray:
    """host statement of the kernels' candidate key: order-preserving fp32 bits << 32 | (0xFFFFFFFF - position)"""
    u = np.asarray(scores, np.float32).view(np.uint32).astype(np.uint64)
    u = np.where(u & np.uint64(0x80000000), ~u & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))
    return (u << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.asarray(positions).astype(np.uint64))


def unpack_keys(keys: np.ndarray):
    keys = np.asarray(keys).astype(np.uint64)
    u = (keys >> np.uint64(32)).astype(np.uint32)
    u = np.where(u & np.uint32(0x80000000), u & np.uint32(0x7FFFFFFF), ~u)
    pos = (np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF))).astype(np.int64)
    return u.astype(np.uint32).view(np.float32), pos


def merge_keys_host(all_keys: np.ndarray, k: int) -> np.ndarray:
    """[L, U, k] packed keys -> [U, k] largest keys per user, descending (reference statement of sbr_topk_merge)"""
    L, U, _ = all_keys.shape
    cat = np.transpose(all_keys, (1, 0, 2)).reshape(U, -1)
    return np.sort(cat.astype(np.uint64), axis=1)[:, ::-1][:, :k]


def allreduce_mean_(flat: torch.Tensor, world: int):
    """in-place mean over ranks (used by the CPU/gloo test of the protocol; the GPU path folds 1/world into Adam)"""
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    return flat


class DataParallelTrainer(FusedTrainer):
    """``sync_bn``: BatchNorm statistics (forward and backward) over the global batch -- an N-rank step then equals the
    single-process step of the reference on the concatenated batch (2 small collectives per BatchNorm layer and step);
    off by default, like ``DistributedDataParallel`` without ``SyncBatchNorm``."""

    def __init__(self, model, learn, n_negative_samples, sync_bn: bool = False, **kw):
        self.world = dist.get_world_size()
        # identical initial weights on every rank
        for p in model.parameters():
            dist.broadcast(p.data, src=0)
        for b in model.buffers():
            dist.broadcast(b, src=0)
        super().__init__(model, learn, n_negative_samples, grad_scale=1.0 / self.world, **kw)
        self._work = []
        # every rank draws its own modalities / dropout masks (the rank enters the Philox seeds)
        self.rt.seed_salt = dist.get_rank()
        if sync_bn:
            from .sbnet import BnSync, SingleBranchNetEntity
            for ent in (self.user, self.item):
                if isinstance(ent, SingleBranchNetEntity):
                    ent.sb_chain.bn_sync = BnSync(self.world)

    def _alloc_flat_grads(self, total: int, dev):
        """the flat gradient buffer in SYMMETRIC memory (CUDA VMM allocations mapped by every rank and bound to an
        NVSwitch multicast object; ``torch.distributed._symmetric_memory`` does the allocation and the handle exchange --
        plumbing).  The gradient collective is then part of the optimizer kernel (``sbr_adam_step_mc``: in-switch
        reduction with ``multimem.ld_reduce``, broadcast with ``multimem.st``, Adam on the sums): no NCCL node in the
        step.  SBR_DP_COLLECTIVE=nccl keeps the two NCCL all-reduce buckets (also the fallback when multicast memory
        is not available: no NVSwitch, older driver)."""
        import os
        self._mc = None
        self._mc_error = None
        if os.environ.get("SBR_DP_COLLECTIVE", "multimem") != "nccl" and self.world > 1:
            try:
                self._mc = self._setup_multicast(total, dev)
                return self._mc["grads"][:total]
            except Exception as e:  # noqa: BLE001
                self._mc, self._mc_error = None, repr(e)
        return torch.zeros(total, dtype=torch.float32, device=dev)

    def _setup_multicast(self, total: int, dev):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import McComm
        world, rank = self.world, dist.get_rank()
        group = dist.group.WORLD.group_name
        padded = -(-total // (4 * world)) * 4 * world
        bufs = {}
        for name, shape, dtype in (("grads", padded, torch.float32), ("sums", padded, torch.float32),
                                   ("flags", max(64, 2 * world), torch.int32)):
            t = symm_mem.empty(shape, dtype=dtype, device=dev)
            t.zero_()
            bufs[name] = t
            bufs[name + "_hdl"] = symm_mem.rendezvous(t, group)
        torch.cuda.synchronize()
        dist.barrier()  # every rank's buffers are zero before anybody signals
        hg, hs, hf = bufs["grads_hdl"], bufs["sums_hdl"], bufs["flags_hdl"]
        if not int(hg.multicast_ptr) or not int(hs.multicast_ptr):
            raise RuntimeError("symmetric memory without a multicast mapping (no NVSwitch multicast on this system)")
        state = torch.zeros(4, dtype=torch.int64, device=dev)
        peer_flags = torch.tensor([int(p) for p in hf.buffer_ptrs], dtype=torch.int64, device=dev)
        comm = McComm()
        comm.flat_grads, comm.mc_grads = bufs["grads"].data_ptr(), int(hg.multicast_ptr)
        comm.sum_local, comm.sum_mc = bufs["sums"].data_ptr(), int(hs.multicast_ptr)
        comm.total = padded
        comm.peer_flags_dev, comm.flags_local, comm.state = peer_flags.data_ptr(), bufs["flags"].data_ptr(), state.data_ptr()
        comm.world, comm.rank = world, rank
        n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
        bufs.update(comm=comm, state=state, peer_flags=peer_flags, total=total,
                    grid=2 * n_sms)  # persistent, all blocks resident (they wait for each other)
        return bufs

    @property
    def collective(self) -> str:
        if self._mc is not None:
            return ("in-switch all-reduce (multimem.ld_reduce / multimem.st over NVSwitch multicast memory) fused into "
                    "the optimizer kernel sbr_adam_step_mc (1 launch / step, no NCCL)")
        return "nccl all-reduce (2 buckets / step)" + (f" [multicast unavailable: {self._mc_error}]" if self._mc_error else "")

    def _after_item_backward(self):
        lo, mid, _ = self.bucket_bounds
        if mid <= lo or not self._reduce_now or self._mc is not None:
            return
        if torch.cuda.is_current_stream_capturing():
            # inside the step's CUDA graph the collective is captured in stream order (the buckets are a few MB:
            # tens of microseconds over NVLink; replay removes ~40 launch latencies instead)
            dist.all_reduce(self.flat_grads[lo:mid], op=dist.ReduceOp.SUM)
        else:
            self._work.append(dist.all_reduce(self.flat_grads[lo:mid], op=dist.ReduceOp.SUM, async_op=True))

    def _after_user_backward(self):
        _, mid, hi = self.bucket_bounds
        if self._mc is not None:
            return  # (the collective is the prologue of the optimizer kernel)
        if hi > mid and self._reduce_now:
            if torch.cuda.is_current_stream_capturing():
                dist.all_reduce(self.flat_grads[mid:hi], op=dist.ReduceOp.SUM)
            else:
                self._work.append(dist.all_reduce(self.flat_grads[mid:hi], op=dist.ReduceOp.SUM, async_op=True))
        for w in self._work:
            w.wait()  # stream-level wait: no host synchronisation
        self._work.clear()

    def optimizer_step(self, ticked: bool = False):
        if self._mc is None:
            if not ticked:  # called on its own after accumulation steps: their rank-local sums are reduced here, once
                dist.all_reduce(self.flat_grads, op=dist.ReduceOp.SUM)
            return super().optimizer_step(ticked)
        b1, b2 = self.betas
        if not ticked:
            ops.tick(self.opt_step_dev)
        mode = {"adam": 0, "adamw": 1, "adagrad": 2}[self.learn.optimizer]
        eps = 1e-10 if mode == 2 and self.eps == 1e-8 else self.eps
        mc = self._mc
        self.adam.step_mc(mc["comm"], mc["grid"], self.learn.lr, b1, b2, eps, self.learn.wd, mode, self.opt_step_dev,
                          self.grad_scale)
        if self.snapshot_grads:  # the summed gradients the update consumed
            if self.grads_snapshot is None:
                self.grads_snapshot = torch.empty_like(self.flat_grads)
            self.grads_snapshot.copy_(mc["sums"][:mc["total"]])


class ShardedEvaluator(FullEvaluator):
    """item-sharded full-catalog evaluation with an all-gather top-k merge"""

    def _shard(self, dataset, dev, rank, world):
        """device-resident inputs of this rank's item shard (built once per dataset)"""
        key = (id(dataset), str(dev), rank, world)
        hit = getattr(self, "_shard_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        users = np.asarray(dataset.users_in_split)
        items = np.asarray(dataset.items_in_split)
        lo, hi = shard_range(len(items), rank, world)
        seen = dataset.exclude_data[users][:, lo:hi].tocsr()
        tgt = dataset.user_sampling_matrix[users][:, items]
        pack = dict(lo=lo, hi=hi, n_users=len(users), n_items=len(items),
                    users=torch.from_numpy(users.astype(np.int64)).to(dev),
                    items=torch.from_numpy(items[lo:hi].astype(np.int64)).to(dev),
                    seen=csr_to_device(seen, dev) if seen.nnz > 0 else (None, None),
                    tgt=csr_to_device(tgt, dev))
        self._shard_cache = (key, pack)
        return pack

    @torch.no_grad()
    def evaluate(self, model, dataset=None, return_topk: bool = False):
        dataset = dataset or self.dataset
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = model.device
        d = self._shard(dataset, dev, rank, world)
        lo, hi, n_users, n_items = d["lo"], d["hi"], d["n_users"], d["n_items"]
        was_training = model.training
        model.eval()
        i_repr = model.get_item_representations(d["items"])
        u_repr = model.get_user_representations(d["users"])
        if was_training:
            model.train()
        ks = sorted(set(int(k) for k in self.config.top_k))
        kmax = min(max(ks), n_items)
        u16, i16 = ops.cast_bf16(u_repr.contiguous()), ops.cast_bf16(i_repr.contiguous())
        k_local = min(kmax, hi - lo)
        local = ops.topk_scores_masked(u16, i16, n_users, hi - lo, u16.shape[1], d["seen"][0], d["seen"][1], k_local,
                                       item_offset=lo, return_keys=True)
        if local.shape[1] < kmax:  # tiny shard: pad with empty slots
            pad = torch.zeros((local.shape[0], kmax - local.shape[1]), dtype=local.dtype, device=dev)
            local = torch.cat([local, pad], dim=1)
        gathered = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=dev)
        dist.all_gather_into_tensor(gathered, local.contiguous())
        vals, idx = ops.topk_merge(gathered, world, n_users, kmax)
        out = self.metrics_from_topk(idx, d["tgt"], ks, n_items)
        return (out, (vals, idx)) if return_topk else out
