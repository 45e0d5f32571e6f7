"""world_size-2 gloo tests (CPU) of the multi-GPU protocols: contiguous item sharding, gradient all-reduce of the flat
buffer, and the item-sharded top-k exchange (local exact top-k as packed keys with GLOBAL positions -> all-gather ->
k-way merge) against a single-process global top-k with the same tie rule (lowest position first)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sibrar_b200.parallel import allreduce_mean_, merge_keys_host, pack_keys, shard_range, unpack_keys


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_topk(scores, k):
    order = np.lexsort((np.broadcast_to(np.arange(scores.shape[1]), scores.shape), -scores), axis=-1)[:, :k]
    return np.take_along_axis(scores, order, 1), order


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. gradient all-reduce (mean) of the flat buffer
        flat = torch.full((1000,), float(rank + 1))
        allreduce_mean_(flat, world)
        assert torch.allclose(flat, torch.full((1000,), (1 + world) / 2))
        # 2. item-sharded top-k
        rng = np.random.default_rng(0)  # same data on every rank
        U, I, D, k = 37, 501, 16, 10
        u = rng.integers(-3, 4, size=(U, D)).astype(np.float32)
        it = rng.integers(-3, 4, size=(I, D)).astype(np.float32)  # integer scores: plenty of ties
        seen = rng.random((U, I)) < 0.05
        scores = u @ it.T
        scores[seen] = -np.inf
        lo, hi = shard_range(I, rank, world)
        lv, li = _global_topk(scores[:, lo:hi], k)
        keys = pack_keys(lv, li + lo)
        keys[~np.isfinite(lv)] = 0  # masked items never become candidates
        local = torch.from_numpy(keys.view(np.int64).copy())
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        merged = merge_keys_host(np.stack([g.numpy().view(np.uint64) for g in gathered]), k)
        vals, pos = unpack_keys(merged)
        gv, gi = _global_topk(scores, k)
        ok = np.isfinite(gv)
        assert (pos[ok] == gi[ok]).all() and (vals[ok] == gv[ok]).all()
        assert (merged[~ok] == 0).all()
        open(os.path.join(tmp, f"ok{rank}"), "w").write("1")
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 370, 3706, 10 ** 6 + 3):
        for world in (1, 2, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_key_packing_orders_like_score_then_lowest_position():
    s = np.array([3.5, 3.5, -1.0, 0.0, -0.0, np.float32(-np.inf), 1e-30], dtype=np.float32)
    p = np.array([7, 2, 0, 5, 4, 9, 1])
    keys = pack_keys(s, p)
    order = np.argsort(keys)[::-1]  # uint64 keys are all distinct
    assert list(order[:2]) == [1, 0]            # tie on 3.5 -> position 2 before 7
    v, q = unpack_keys(keys)
    assert (q == p).all() and (v.view(np.uint32) == s.view(np.uint32)).all()


def test_two_rank_protocols_over_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
