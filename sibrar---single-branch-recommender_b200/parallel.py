"""One-process-per-GPU scaling of the hot path over NCCL / NVLink (``torch.distributed`` is plumbing only).

The reference is single-device (no collective anywhere); this is the B200 addition:
  * training: data parallel.  Every rank runs the fused step on its own B interactions; the flat fp32 gradient buffer
    is all-reduced in two buckets -- the item-entity bucket is launched as soon as the item backward has been
    enqueued, so the collective overlaps the user-entity backward -- and the multi-tensor Adam kernel divides by the
    world size.  BatchNorm statistics and the user-side in-batch InfoNCE stay rank-local.
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
        """the flat gradient buffer in SYMMETRIC memory (every rank maps every peer's buffer + the NVSwitch multicast
        address): the gradient collective is then ONE in-switch all-reduce (`multimem.ld_reduce` of each rank's slice +
        `multimem.st` of the sums, torch's symm_mem kernels) instead of two NCCL calls.  SBR_DP_COLLECTIVE=nccl keeps
        the NCCL buckets (also the fallback when symmetric memory / multicast is not available)."""
        import os
        self._symm = None
        if os.environ.get("SBR_DP_COLLECTIVE", "symm") != "nccl" and self.world > 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                t = symm_mem.empty(total, dtype=torch.float32, device=dev)
                t.zero_()
                hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
                self._symm = (hdl, dist.group.WORLD.group_name,
                              "multimem" if hdl.has_multicast_support() else "two_shot")
                return t
            except Exception as e:  # noqa: BLE001  (older drivers / no P2P): NCCL
                self._symm_error = repr(e)
        return torch.zeros(total, dtype=torch.float32, device=dev)

    @property
    def collective(self) -> str:
        return f"symm_mem {self._symm[2]} all-reduce (1 call / step)" if self._symm else "nccl all-reduce (2 buckets / step)"

    def _after_item_backward(self):
        lo, mid, _ = self.bucket_bounds
        if mid <= lo or not self._reduce_now or self._symm:
            return
        if torch.cuda.is_current_stream_capturing():
            # inside the step's CUDA graph the collective is captured in stream order (the buckets are a few MB:
            # tens of microseconds over NVLink; replay removes ~40 launch latencies instead)
            dist.all_reduce(self.flat_grads[lo:mid], op=dist.ReduceOp.SUM)
        else:
            self._work.append(dist.all_reduce(self.flat_grads[lo:mid], op=dist.ReduceOp.SUM, async_op=True))

    def _after_user_backward(self):
        _, mid, hi = self.bucket_bounds
        if self._symm and self._reduce_now:
            _, group, kind = self._symm
            if kind == "multimem":
                torch.ops.symm_mem.multimem_all_reduce_(self.flat_grads, "sum", group)
            else:
                torch.ops.symm_mem.two_shot_all_reduce_(self.flat_grads, "sum", group)
            return
        if hi > mid and self._reduce_now:
            if torch.cuda.is_current_stream_capturing():
                dist.all_reduce(self.flat_grads[mid:hi], op=dist.ReduceOp.SUM)
            else:
                self._work.append(dist.all_reduce(self.flat_grads[mid:hi], op=dist.ReduceOp.SUM, async_op=True))
        for w in self._work:
            w.wait()  # stream-level wait: no host synchronisation
        self._work.clear()


    def optimizer_step(self, ticked: bool = False):
        if not ticked:  # called on its own after accumulation steps: their rank-local sums are reduced here, once
            dist.all_reduce(self.flat_grads, op=dist.ReduceOp.SUM)
        super().optimizer_step(ticked)


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
