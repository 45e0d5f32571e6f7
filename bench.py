#!/usr/bin/env python
"""Benchmark of the SBNet hot path on B200 (one process per GPU).

    python bench.py --gpus N --steps K --warmup W              # B200 path
    python bench.py --impl reference --steps K --warmup W       # the reference's CPU path (rank 0 only)

Headline metric (BASELINE.json): SBNet train interactions/sec on the synthetic ML-1M shape, cold-start item split,
modality dropout, bf16 (configs[1]); full-catalog eval users/sec is reported in the same JSON line under "eval".
A "step" is one fused train step (forward, BPR + reg losses, backward, AdamW) on one batch of B interactions
(B per GPU is fixed -> weak scaling).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sbnet_train_interactions_per_sec"
UNIT = "interactions/s"


def ml1m_model_conf(D=64):
    """model block of the reference's conf/single/algorithms/sbnet_ml1m_conf.yml:21-52"""
    ent = lambda feats, hidden, drop: dict(  # noqa: E731
        features=[dict(feature_name=f, feature_hidden_layers=[]) for f in feats], single_branch_hidden_layers=hidden,
        preference_hidden_layers=[], common_modality_dim=D, activation_fn="relu", single_branch_input_dropout=drop)
    return dict(shared_common_dim=D, user=ent(["interactions", "gender", "occupation"], [], None),
                item=ent(["interactions", "genres", "plot_mpnet"], [D], 0.2))


LEARN = dict(lr=1e-3, wd=1e-6, optimizer="adamw", rec_loss="bpr", loss_aggregator="mean")
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of this workload's kernels
# (profiles/r01_segreduce_ncu_raw.txt, profiles/r01_gemm_bits_ncu_raw.txt); keyed like CallProfiler._key
NCU_DRAM_TRAFFIC = {("sbr_row_gather_bwd_segmented", 180224, 64): 49.25e6,
                    # bit matrix + weights from DRAM; the fp32 K-partition slices stay in L2 for the reduce pass
                    ("sbr_gemm_bits_bf16", 3706, 64, 6040, 10): 3.668e6}
N_NEG = 10


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_train_arm(corpus, steps, warmup, batch=256, budget_s=25.0):
    """the reference's CPU implementation of the train step: the real reference (PyTorch CPU, all host threads) when
    its tree is present, otherwise the numpy oracle port pinned against it (oracle/sbnet_oracle.py)."""
    from oracle import ref_shims
    from sibrar_b200.synthetic import sample_batch
    conf = ml1m_model_conf()
    rng = np.random.default_rng(5)
    train = corpus.dataset("train")
    cores = os.cpu_count() or 1
    # the real reference is only timed on request: nothing reads /root/reference at run time by default
    if os.environ.get("SBR_USE_REFERENCE") == "1" and ref_shims.find_reference() is not None:
        import copy
        import torch
        from oracle.make_golden import build_reference_datasets
        ref_shims.install()
        from algorithms.sgd_alg import SingleBranchNet as RefNet
        from train.rec_losses import RecBayesianPersonalizedRankingLoss
        torch.set_num_threads(cores)
        ds = build_reference_datasets(corpus)["train"]
        model = RefNet.build_from_conf(copy.deepcopy(conf), ds).train()
        loss_fn = RecBayesianPersonalizedRankingLoss(n_items=ds.n_items, aggregator="mean",
                                                     train_neg_strategy="uniform_recbole", neg_train=N_NEG)
        opt = torch.optim.AdamW(model.parameters(), lr=LEARN["lr"], weight_decay=LEARN["wd"])

        def one():
            u, i = sample_batch(train, batch, rng, N_NEG)
            ut, it = torch.from_numpy(u), torch.from_numpy(i)
            labels = torch.zeros(i.shape, dtype=torch.float64)
            labels[:, 0] = 1.
            out = model(ut, it)
            loss = loss_fn.compute_loss(out, labels) + model.get_and_reset_other_loss()["reg_loss"]
            loss.item()
            loss.backward()
            opt.step()
            opt.zero_grad()
        kind = "reference"
    else:
        from oracle import sbnet_oracle as O
        import torch
        from sibrar_b200.sbnet import SingleBranchNet
        torch.manual_seed(0)
        init = SingleBranchNet.build_from_conf(conf, corpus.dataset("train"))
        p = {k: v.detach().numpy().astype(np.float64) if v.dtype.is_floating_point else v.numpy()
             for k, v in init.state_dict().items()}
        names = {"user": init.user_embedding_module.mod_names, "item": init.item_embedding_module.mod_names}
        net, state, cnt = O.OracleSBNet(conf, train), {}, [0]

        def one():
            u, i = sample_batch(train, batch, rng, N_NEG)
            mods = {"user": rng.integers(0, 3, size=(batch, 1)), "item": rng.integers(0, 3, size=(batch, 1 + N_NEG, 1))}
            drop = {"item": (rng.random((batch * (1 + N_NEG), 64)) >= 0.2).astype(np.float32)}
            r = net.train_step_fwd_bwd(p, u, i, mods, names, drop, loss_kind="bpr")
            cnt[0] += 1
            O.adam_step(p, r["grads"], state, LEARN["lr"], LEARN["wd"], cnt[0], decoupled=True)
            p.update(r["new_stats"])
        kind = "port"
    for _ in range(max(1, min(warmup, 3))):
        one()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        one()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=batch * done / dt, unit=UNIT, cores=cores, kind=kind,
                sample=f"{done} train steps of B={batch} (1+{N_NEG} items each) on the same synthetic ML-1M corpus, "
                       f"{'reference PyTorch CPU path' if kind == 'reference' else 'numpy oracle port'}"), dt / done


# ------------------------------------------------------------------------------------------------ profiling pass
class CallProfiler:
    """brackets every C-ABI call with CUDA events (separate, untimed pass) to find the dominant kernel"""

    def __init__(self, ops_mod, torch):
        self.ops, self.torch, self.rec = ops_mod, torch, []
        self._orig = ops_mod.call

    def __enter__(self):
        def wrapped(name, *args):
            a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            a.record()
            self._orig(name, *args)
            b.record()
            self.rec.append((name, self._key(name, args), a, b))
        self.ops.call = wrapped
        return self

    def __exit__(self, *exc):
        self.ops.call = self._orig

    @staticmethod
    def _key(name, args):
        if name == "sbr_gemm_bf16":
            ep = args[9]._obj
            return ("gemm", int(args[6]), int(args[7]), int(args[8]), int(args[2]), int(args[5]),
                    bool(ep.out_bf16), bool(ep.out_f32), int(ep.transpose_out))
        if name in ("sbr_row_gather_fwd", "sbr_row_gather_bwd"):
            return (name, int(args[4]) * int(args[5]), int(args[6]))
        if name == "sbr_row_gather_bwd_segmented":
            return (name, int(args[6]), int(args[7]))
        if name == "sbr_bn_apply":
            return (name, int(args[6]), int(args[7]), bool(args[8]), bool(args[10]))
        if name == "sbr_bn_bwd_reduce":
            return (name, int(args[9]), int(args[10]))
        if name == "sbr_bn_bwd_apply":
            return (name, int(args[12]), int(args[13]))
        if name == "sbr_actgrad_colsum":
            return (name, int(args[6]), int(args[7]))
        if name == "sbr_score_loss":
            return (name, int(args[2]), int(args[3]), int(args[4]), int(args[5]), int(args[6]))
        if name == "sbr_score_loss_bn":
            return (name, int(args[4]), int(args[5]), int(args[6]))
        if name == "sbr_gemm_bits_bf16":
            ep = args[8]._obj
            return (name, int(args[5]), int(args[6]), int(args[7]), max(1, int(ep.split_k)))
        if name == "sbr_splitk_reduce":
            return (name, int(args[1]), int(args[4]), int(args[5]))
        return (name,)

    def summary(self):
        self.torch.cuda.synchronize()
        agg = {}
        for name, key, a, b in self.rec:
            ms = a.elapsed_time(b)
            e = agg.setdefault(key, [0.0, 0])
            e[0] += ms
            e[1] += 1
        return agg


def algorithmic_work(key):
    """(flops, bytes) one launch must do/move at minimum (DESIGN.md 'rooflines')"""
    if key[0] == "gemm":
        _, M, N, K, a_mn, b_mn, o16, o32, tr = key
        out = M * N * ((2 if o16 else 0) + (4 if o32 else 0))
        return 2.0 * M * N * K, 2.0 * (M * K + N * K) + out
    if key[0] == "sbr_row_gather_fwd":
        _, rows, C = key
        return 0.0, rows * C * (4 + 2) + rows * 9
    if key[0] == "sbr_row_gather_bwd":
        _, rows, C = key
        return 0.0, rows * C * (4 + 4 + 4) + rows * 9
    if key[0] == "sbr_row_gather_bwd_segmented":  # dx rows (fp32) + (sorted key, permutation) per row
        _, rows, C = key
        return 0.0, rows * C * 4 + rows * 8
    if key[0] == "sbr_bn_apply":  # z fp32 in, bf16 and/or fp32 out
        _, rows, C, o16, o32 = key
        return 0.0, rows * C * (4 + (2 if o16 else 0) + (4 if o32 else 0))
    if key[0] == "sbr_bn_bwd_reduce":  # dy, z
        _, rows, C = key
        return 0.0, rows * C * 8
    if key[0] == "sbr_bn_bwd_apply":  # dy, z in; dz bf16 out
        _, rows, C = key
        return 0.0, rows * C * 10
    if key[0] == "sbr_actgrad_colsum":  # dT fp32 in (+ cleared), T in, bf16 out
        _, rows, C = key
        return 0.0, rows * C * (4 + 4 + 4 + 2)
    if key[0] == "sbr_score_loss":  # embeddings in, gradients out (fp32)
        _, B, n, ku, ki, D = key
        return 0.0, 2.0 * B * D * 4 * (ku + n * ki)
    if key[0] == "sbr_score_loss_bn":  # pre-BatchNorm item/user rows in (fp32), their gradients out (fp32), logits
        _, B, n, D = key
        return 0.0, 2.0 * B * D * 4 * (1 + n) + B * n * 4
    if key[0] == "sbr_gemm_bits_bf16":  # bit-packed A, bf16 B, fp32 output (one slice per K partition)
        _, M, N, K, split = key
        return 2.0 * M * N * K, M * K / 8 + 2.0 * N * K + 4.0 * M * N * split
    if key[0] == "sbr_splitk_reduce":  # fp32 slices in, fp32 + bf16 out
        _, split, rows, cols = key
        return 0.0, rows * cols * (4.0 * split + 6)
    return 0.0, 0.0


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("SBR_BENCH_BATCH", 16384)),
                    help="interactions per GPU per step")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel (no CUDA graph)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    import sibrar_b200  # noqa: F401
    from sibrar_b200.synthetic import SynCorpus
    workload = dict(workload="SBNet train step, synthetic ML-1M shape (6040 users x 3706 items, 1,000,209 interactions; "
                             "user: interactions/gender/occupation, item: interactions/genres(18 tags)/plot_mpnet(768)), "
                             "cold_start_item split, sbnet_ml1m_conf model (C=D=64, item MLP [64], item input dropout 0.2, "
                             "trailing BatchNorm), BPR, AdamW, n_neg=10",
                    batch_per_gpu=args.batch, n_neg=N_NEG)

    if args.impl == "reference":
        if rank != 0:
            return
        corpus = SynCorpus("ml1m", "cold_start_item", seed=42)
        cb, s_per_step = cpu_train_arm(corpus, args.steps, args.warmup, budget_s=150.0)
        line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=s_per_step * 1e3, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                    config=dict(workload, batch_per_gpu=256, note="CPU arm: bounded sample, one B=256 batch per step"),
                    cpu_baseline=cb, e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from sibrar_b200 import _lib, ops
    from sibrar_b200.evaluator import FullEvaluator
    from sibrar_b200.sbnet import SingleBranchNet
    from sibrar_b200.trainer import FusedTrainer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    _lib.lib()  # fail loudly here if the CUDA extension is missing

    corpus = SynCorpus("ml1m", "cold_start_item", seed=42)
    train = corpus.dataset("train")
    torch.manual_seed(1234)
    model = SingleBranchNet.build_from_conf(ml1m_model_conf(), train).to(dev).train()
    from sibrar_b200.parallel import DataParallelTrainer
    tr = DataParallelTrainer(model, LEARN, n_negative_samples=N_NEG, cuda_graph=not args.no_graph) if world > 1 else \
        FusedTrainer(model, LEARN, n_negative_samples=N_NEG, cuda_graph=not args.no_graph)

    # ---- synthetic batches, sampled on the device by the GPU sampler (resident in HBM before the timed region)
    B, n = args.batch, 1 + N_NEG
    coo = train.interaction_matrix
    d = lambda a, t: torch.from_numpy(np.ascontiguousarray(a).astype(t)).to(dev)  # noqa: E731
    csr = train.user_sampling_matrix_train
    coo_u, coo_i = d(coo.row, np.int32), d(coo.col, np.int32)
    ip, ix, items = d(csr.indptr, np.int64), d(csr.indices, np.int32), d(train.items_in_split, np.int32)
    n_batches = args.steps + args.warmup
    sample_step = torch.zeros(1, dtype=torch.int64, device=dev)
    batches = []
    for b in range(n_batches):
        ops.tick(sample_step)
        u = torch.empty(B, dtype=torch.int64, device=dev)
        i = torch.empty((B, n), dtype=torch.int64, device=dev)
        ops.sample_batch(coo_u, coo_i, ip, ix, items, B, N_NEG, 1000 + rank, sample_step, u, i)
        batches.append((u, i))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        tr.step(*batches[w])
    sync_all()
    model.check_errors()
    tr.read_losses()

    # ---- timed region: K steps, CUDA events on the launching stream, L2 flushed between steps (outside the events)
    _lib.reset_launch_counter()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    t_wall = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        tr.step(*batches[args.warmup + k])
        ev[k][1].record()
    sync_all()
    t_wall = time.perf_counter() - t_wall
    clk = clocks.stop()
    launches = _lib.launch_counter()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    losses = tr.read_losses()
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- e2e: host buffers in, loss out, through the public trainer API
    # (double-buffered: the pinned-memory fill and the H2D copy of batch k+1 run on a copy stream while step k
    # computes; every step still ends with a device -> host read of its losses)
    hu = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(2)]
    hi = [torch.empty((B, n), dtype=torch.int64).pin_memory() for _ in range(2)]
    host_batches = [(b[0].cpu(), b[1].cpu()) for b in batches[:4]]
    dbuf = [(torch.empty_like(batches[0][0]), torch.empty_like(batches[0][1])) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    loss_host = torch.empty(tr.loss_acc.shape, dtype=tr.loss_acc.dtype).pin_memory()
    loss_read = torch.cuda.Event()

    def stage(k):
        s_ = k % 2
        src = host_batches[k % len(host_batches)]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s_])  # the step that last read this device buffer has finished
            hu[s_].copy_(src[0])
            hi[s_].copy_(src[1])
            dbuf[s_][0].copy_(hu[s_], non_blocking=True)
            dbuf[s_][1].copy_(hi[s_], non_blocking=True)
            ready[s_].record(copy_stream)

    sync_all()
    for e in consumed:
        e.record()
    t0 = time.perf_counter()
    stage(0)
    for k in range(args.steps):
        torch.cuda.current_stream().wait_event(ready[k % 2])
        tr.step(*dbuf[k % 2])
        consumed[k % 2].record()
        if k + 1 < args.steps:
            stage(k + 1)
        loss_host.copy_(tr.loss_acc, non_blocking=True)  # device -> host read of the step's losses ...
        loss_read.record()
        loss_read.synchronize()                           # ... waited for before the next step is issued
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = dict(value=world * B * args.steps / float(t.item()), unit=UNIT, h2d_bytes_per_step=B * 8 + B * n * 8,
               d2h_bytes_per_step=int(tr.loss_acc.numel() * 8))
    tr.read_losses()

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="bf16", data="synthetic",
                config=dict(workload, global_batch=world * B, parallelism=f"dp{world}", l2="flushed between timed steps "
                            "(256 MiB write)", wall_ms_per_step_incl_flush=t_wall / args.steps * 1e3),
                clocks=clk, e2e=e2e, gpu_launches=launches, train_loss=losses.get("train/loss"))

    # ---- item-sharded evaluation (all ranks): local exact top-k per shard -> all-gather of [U, k] keys -> merge
    sharded = None
    if world > 1 and not args.no_eval:
        from sibrar_b200.parallel import ShardedEvaluator
        sev = ShardedEvaluator(dict(top_k=[1, 10, 20], metrics=["ndcg", "recall", "precision", "coverage"],
                                    calculate_std=False))
        val = corpus.dataset("val")
        res = sev.evaluate(model, val)
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            res = sev.evaluate(model, val)
        b.record()
        sync_all()
        t = torch.tensor([a.elapsed_time(b) / 3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sharded = dict(metric="eval_users_per_sec", value=val.n_users_in_split / (float(t.item()) * 1e-3),
                       unit="users/s", ms=float(t.item()), shards=world, ndcg10=res.get("ndcg@10"))

    # ---- roofline of the dominant kernel (separate profiled pass, CUDA events around every C-ABI call); every rank
    # runs these steps (they contain the gradient all-reduce), rank 0 records them
    tr.cuda_graph = False  # the per-call pass launches kernel by kernel ...
    tr.branches = 0        # ... on one stream: concurrent branches would stretch each other's bracketed durations
    if rank != 0:
        for k in range(3):
            flush.zero_()
            torch.cuda._sleep(6_000_000)
            tr.step(*batches[(args.warmup + k) % len(batches)])
    if rank == 0:
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            which = "measured"
        except Exception:
            peaks, which = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), "fallback"
        with CallProfiler(ops, torch) as prof:
            for k in range(3):
                flush.zero_()
                # the host needs longer to launch the ~45 kernels of an eager step than the GPU to run them: a spin
                # kernel in front lets the host run ahead, so that the events bracket kernel time, not launch gaps
                torch.cuda._sleep(6_000_000)
                tr.step(*batches[(args.warmup + k) % len(batches)])
        agg = prof.summary()
        step_ms = sum(v[0] for v in agg.values()) / 3
        top = sorted(agg.items(), key=lambda kv: -kv[1][0])
        line["kernel_shares"] = [dict(kernel=str(k), share=round(v[0] / 3 / step_ms, 4), launches_per_step=v[1] / 3,
                                      avg_us=round(v[0] / v[1] * 1e3, 2)) for k, v in top[:8]]
        # dominant kernel = largest share among the calls with a roofline model; of two calls within 10 % of each other
        # (the user- and the item-side projection GEMM trade places from run to run) the one with an ncu DRAM capture
        modelled = [(k, v) for k, v in top if algorithmic_work(k) != (0.0, 0.0)]
        if modelled:
            near = [kv for kv in modelled if kv[1][0] >= 0.9 * modelled[0][1][0] and kv[0] in NCU_DRAM_TRAFFIC]
            modelled = (near or modelled)[:1]
        for key, (ms, cnt) in modelled:
            flops, nbytes = algorithmic_work(key)
            if flops == 0 and nbytes == 0:
                continue
            dur = ms / cnt * 1e-3
            ridge = peaks["bf16_tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
            if flops > 0 and flops / max(1.0, nbytes) > ridge:
                ach, peak, unit, bound = flops / dur / 1e12, peaks["bf16_tflops"], "TFLOP/s", "tensor"
            else:
                ach, peak, unit, bound = nbytes / dur / 1e9, peaks["hbm_gbs"], "GB/s", "hbm"
            line["roofline"] = dict(bound=bound, achieved=ach, peak=peak, unit=unit, frac=ach / peak,
                                    traffic=NCU_DRAM_TRAFFIC.get(key),
                                    kernel=str(key), peak_source=which + (" burst" if bound == "tensor" else ""),
                                    share_of_step=ms / 3 / step_ms)
            break
        tr.read_losses()

        # ---- full-catalog evaluation (users/s): the workload's val split end to end + two sweep points (configs[4])
        if not args.no_eval:
            try:
                line["eval"] = bench_eval(torch, ops, model, corpus, FullEvaluator, dev)
                if sharded is not None:
                    line["eval"]["item_sharded_val_split"] = sharded
            except Exception as e:  # keep the headline line
                line["eval"] = {"error": repr(e)}
        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], _ = cpu_train_arm(corpus, 10 ** 6, 2, budget_s=20.0)
            except Exception as e:
                line["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(line))
    if world > 1:
        # graphs that captured NCCL kernels must go before the communicator does; then leave without running the
        # interpreter's teardown of NCCL (a captured collective can make destroy_process_group wait forever)
        tr._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def bench_eval(torch, ops, model, corpus, FullEvaluator, dev):
    out = {}
    val = corpus.dataset("val")
    ev = FullEvaluator(dict(top_k=[1, 10, 20], metrics=["ndcg", "recall", "precision", "coverage"], calculate_std=False))
    ev.evaluate(model, val)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a.record()
    for _ in range(reps):
        res = ev.evaluate(model, val)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    out["workload_val_split"] = dict(metric="eval_users_per_sec", value=val.n_users_in_split / (ms * 1e-3),
                                     unit="users/s", users=int(val.n_users_in_split), items=int(val.n_items_in_split),
                                     ms=ms, includes="item+user representations, scores, mask, top-20, metrics",
                                     ndcg10=res.get("ndcg@10"))
    evg = FullEvaluator(dict(top_k=[1, 10, 20], metrics=["ndcg", "recall", "precision", "coverage"],
                             calculate_std=False), cuda_graph=True)
    for _ in range(3):  # eager, capture, first replay
        resg = evg.evaluate(model, val)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        resg = evg.evaluate(model, val)
    b.record()
    torch.cuda.synchronize()
    msg = a.elapsed_time(b) / reps
    out["workload_val_split_graph"] = dict(metric="eval_users_per_sec", value=val.n_users_in_split / (msg * 1e-3),
                                           unit="users/s", ms=msg, same_result=bool(resg == res),
                                           includes="the same evaluation replayed as one CUDA graph + one D2H copy")
    g = torch.Generator(device="cpu").manual_seed(7)
    sweep = []
    for U, I, D, k in ((100_000, 100_000, 64, 10), (100_000, 1_000_000, 64, 10), (100_000, 1_000_000, 128, 50)):
        u16 = (torch.randn(U, D, generator=g) / math.sqrt(D)).to(torch.bfloat16).to(dev)
        i16 = torch.randn(I, D, generator=g).to(torch.bfloat16).to(dev)
        seen_ip = torch.arange(0, (U + 1) * 100, 100, dtype=torch.int64, device=dev)
        seen_ix = torch.sort(torch.randint(0, I, (U, 100), device=dev, dtype=torch.int32), dim=1).values.reshape(-1)
        ops.topk_scores_masked(u16, i16, U, I, D, seen_ip, seen_ix.contiguous(), k)
        torch.cuda.synchronize()
        a.record()
        ops.topk_scores_masked(u16, i16, U, I, D, seen_ip, seen_ix, k)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        sweep.append(dict(U=U, I=I, D=D, k=k, ms=ms, users_per_s=U / (ms * 1e-3),
                          tflops=2.0 * U * I * D / (ms * 1e-3) / 1e12))
        del u16, i16, seen_ix
    out["sweep"] = sweep
    return out


if __name__ == "__main__":
    main()
