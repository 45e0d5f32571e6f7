#!/usr/bin/env python
"""Benchmark of the SBNet hot path on B200 (one process per GPU).

    python bench.py --gpus N --steps K --warmup W                    # B200 path, headline + the other configs
    python bench.py --config onion18_huge --gpus N ...                # one BASELINE config as the headline
    python bench.py --impl reference --steps K --warmup W              # the reference's CPU path (rank 0 only)

Headline metric (BASELINE.json): SBNet train interactions/sec; the default headline workload is configs[1] (synthetic
ML-1M shape, cold-start item split, modality dropout, bf16).  The same JSON line carries
  * "configs": the train step of configs[2] (Onion18 shape, huge model) and configs[3] (AmazonVideo2024 shape, plain user
    embedding, missing-modality evaluation) measured the same way (fewer steps),
  * "paper_batch": the headline workload at the reference's batch size (B = 256, conf/single/dataloader_conf.yml:13),
  * "eval": full-catalog evaluation users/sec (the workload's own split, the configs[4] sweep with its own roofline, a CPU
    evaluation baseline, and under torchrun the item-sharded sweep),
  * "roofline": the dominant kernel FAMILY of the headline step (largest share of the step's kernel time).
A "step" is one fused train step (forward, BPR + InfoNCE losses, backward, AdamW) on one batch of B interactions per GPU
(B per GPU fixed -> weak scaling).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import gc
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
if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every host core whatever launched it
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np  # noqa: E402

METRIC = "sbnet_train_interactions_per_sec"
UNIT = "interactions/s"
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
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
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
def _all_cores():
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    return cores


def cpu_train_arm(corpus, conf, learn, batch, steps, warmup, budget_s=25.0):
    """the reference's CPU implementation of the train step on the host cores: the numpy oracle port pinned against the
    reference (oracle/sbnet_oracle.py; `kind: "port"`) -- the one place besides smoke()/tests that executes oracle/."""
    import torch
    from oracle import sbnet_oracle as O
    from sibrar_b200.sbnet import SingleBranchNet, SingleBranchNetEntity
    from sibrar_b200.synthetic import sample_batch
    cores = _all_cores()
    rng = np.random.default_rng(5)
    train = corpus.dataset("train")
    torch.manual_seed(0)
    init = SingleBranchNet.build_from_conf(conf, corpus.dataset("train"))
    p = {k: v.detach().numpy().astype(np.float64) if v.dtype.is_floating_point else v.numpy()
         for k, v in init.state_dict().items()}
    ents = {"user": init.user_embedding_module, "item": init.item_embedding_module}
    names = {n: e.mod_names for n, e in ents.items() if isinstance(e, SingleBranchNetEntity)}
    net, state, cnt = O.OracleSBNet(conf, train), {}, [0]
    n = 1 + N_NEG

    def one():
        u, i = sample_batch(train, batch, rng, N_NEG)
        mods, drop = {}, {}
        for en, e in ents.items():
            if not isinstance(e, SingleBranchNetEntity):
                continue
            shape = (batch,) if en == "user" else (batch, n)
            k, nm = e.k_train, len(e.mod_names)
            m = np.stack([rng.permutation(nm)[:k] for _ in range(int(np.prod(shape)))]).reshape(shape + (k,)) \
                if k > 1 else rng.integers(0, nm, size=shape + (1,))
            mods[en] = m
            pd = e.entity_config.single_branch_input_dropout
            if pd:
                drop[en] = (rng.random((int(np.prod(shape)) * k, e.entity_config.common_modality_dim)) >= pd
                            ).astype(np.float32)
        r = net.train_step_fwd_bwd(p, u, i, mods, names, drop, loss_kind=learn["rec_loss"])
        cnt[0] += 1
        O.adam_step(p, r["grads"], state, learn["lr"], learn["wd"], cnt[0], decoupled=learn["optimizer"] == "adamw")
        p.update(r["new_stats"])
    t_w = time.perf_counter()
    for _ in range(max(1, min(warmup, 2))):
        one()
        if time.perf_counter() - t_w > budget_s / 3:
            break
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        one()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=batch * done / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"{done} train steps of B={batch} (1+{N_NEG} items each) on the same synthetic corpus, numpy "
                       f"fp64 oracle port of the reference step (forward, BPR + InfoNCE, backward, AdamW), BLAS threads = "
                       f"{cores}"), dt / done


def cpu_eval_arm(U=2048, I=100_000, D=64, k=10, seen_per_user=100, seed=7):
    """CPU statement of eval/eval.py:205-222 on a bounded user sample: scores = U_repr @ I_repr^T, seen -> -inf, top-k,
    ndcg / recall / precision (oracle/sbnet_oracle.py: masked_topk + metrics_at_k)."""
    import scipy.sparse as sp
    from oracle import sbnet_oracle as O
    cores = _all_cores()
    rng = np.random.default_rng(seed)
    u = (rng.standard_normal((U, D)) / math.sqrt(D)).astype(np.float32)
    it = rng.standard_normal((I, D)).astype(np.float32)
    cols = rng.integers(0, I, size=(U, seen_per_user))
    seen = sp.csr_matrix((np.ones(cols.size, dtype=bool), (np.repeat(np.arange(U), seen_per_user), cols.reshape(-1))),
                         shape=(U, I))
    tcols = rng.integers(0, I, size=(U, 10))
    tgt = sp.csr_matrix((np.ones(tcols.size, dtype=np.int8), (np.repeat(np.arange(U), 10), tcols.reshape(-1))),
                        shape=(U, I))
    t0 = time.perf_counter()
    done = 0
    for lo in range(0, U, 256):  # the reference's eval batch (eval/eval.py:212)
        sl = slice(lo, min(U, lo + 256))
        _, idx = O.masked_topk(u[sl], it, seen[sl], k)
        O.metrics_at_k(idx, tgt[sl], [k], n_items=I)
        done += sl.stop - sl.start
        if time.perf_counter() - t0 > 20.0:
            break
    dt = time.perf_counter() - t0
    return dict(value=done / dt, unit="users/s", cores=cores, kind="port",
                sample=f"{done} users x {I} items, D={D}, top-{k}, {seen_per_user} seen items per user, batches of 256 "
                       f"users (numpy oracle port of eval/eval.py:205-222)")


# ------------------------------------------------------------------------------------------------ profiling pass
class CallProfiler:
    """brackets every C-ABI call with CUDA events (separate, untimed pass) to find the dominant kernel family"""

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
        """(name, shape...) of one call: enough to compute its algorithmic work"""
        i = lambda j: int(args[j])  # noqa: E731
        if name == "sbr_gemm_bf16":
            ep = args[9]._obj
            return ("sbr_gemm_bf16", i(6), i(7), i(8), bool(ep.out_bf16), bool(ep.out_f32), bool(ep.actgrad_y))
        if name == "sbr_gemm_bits_bf16":
            ep = args[8]._obj
            return (name, i(5), i(6), i(7), bool(ep.out_bf16), bool(ep.out_f32), False)
        if name in ("sbr_mlp2_fwd", "sbr_mlp2_fwd_bn"):  # (_bn: the same kernel + the in-kernel BatchNorm finalize)
            return ("sbr_mlp2_fwd", i(1), i(2))
        if name == "sbr_mlp2_bwd":
            return (name, i(1), i(2))
        if name == "sbr_row_gather_fwd":
            return (name, i(4) * i(5), i(6))
        if name == "sbr_row_gather_bwd_segmented":
            return (name, i(6), i(7))
        if name == "sbr_gather_plan":
            return (name, i(4) * i(5), i(6))
        if name == "sbr_bn_apply":
            return (name, i(6), i(7), bool(args[8]), bool(args[10]))
        if name == "sbr_bn_finalize":
            return (name, i(1), i(3))
        if name == "sbr_bn_bwd_reduce":
            return (name, i(9), i(10))
        if name == "sbr_bn_bwd_apply":
            return (name, i(12), i(13), bool(args[14]), bool(args[16]))
        if name == "sbr_actgrad_colsum":
            return (name, i(6), i(7))
        if name == "sbr_score_loss":
            return (name, i(2), i(3), i(4), i(5), i(6))
        if name == "sbr_score_loss_bn":
            return (name, i(4), i(5), i(6))
        if name == "sbr_infonce":
            return (name, i(1), i(2), i(3))
        if name == "sbr_splitk_reduce":
            return (name, i(1), i(4), i(5))
        if name in ("sbr_spmm_csr", "sbr_spmm_csr_bf16"):
            return (name, i(3), i(6))
        if name == "sbr_adam_step":
            return (name, i(2))
        if name in ("sbr_tag_bag_fwd", "sbr_tag_bag_bwd"):
            return (name, i(4), i(5), i(1))
        if name == "sbr_sample_modalities":
            return (name, i(1) * i(2))
        if name == "sbr_step_begin":
            return (name, i(3))
        if name == "sbr_aggregate":
            return (name, i(1), i(2), i(3))
        if name in ("sbr_cast_f32_to_bf16", "sbr_transpose_f32"):
            return (name, i(4), i(5))
        return (name,)

    def summary(self):
        self.torch.cuda.synchronize()
        agg = {}
        for name, key, a, b in self.rec:
            e = agg.setdefault(key, [0.0, 0])
            e[0] += a.elapsed_time(b)
            e[1] += 1
        return agg


def family_of(key):
    """kernel family = the kernel (template instantiation) a call runs: GEMMs by their N-tile class"""
    if key[0] in ("sbr_gemm_bf16", "sbr_gemm_bits_bf16"):
        N = key[2]
        return f"{key[0]}<BN={64 if N <= 64 else 128 if N <= 128 else 256}>"
    return key[0]


def algorithmic_work(key, nnz_hint=None):
    """(flops, bytes) ONE launch must do / move at minimum (split-K slices, scratch and re-reads are
    NOT algorithmic).  None = no model."""
    n = key[0]
    if n in ("sbr_gemm_bf16", "sbr_gemm_bits_bf16"):
        _, M, N, K, o16, o32, ag = key
        a_bytes = M * K / 8 if n == "sbr_gemm_bits_bf16" else 2.0 * M * K
        out = M * N * ((2 if o16 else 0) + (4 if o32 else 0)) + (2 * M * N if ag else 0)
        return 2.0 * M * N * K, a_bytes + 2.0 * N * K + out
    if n == "sbr_mlp2_fwd":  # gathered fp32 source row in, pre-BatchNorm z fp32 out (+ index / modality / keep bits)
        _, rows, C = key
        return 4.0 * rows * C * C, rows * C * (4 + 4) + rows * (9 + C / 8)
    if n == "sbr_mlp2_bwd":  # dE + z in (fp32), re-gathered source row in, dx fp32 out
        _, rows, C = key
        return 12.0 * rows * C * C, rows * C * (4 + 4 + 4 + 4) + rows * (9 + C / 8)
    if n == "sbr_row_gather_fwd":
        _, rows, C = key
        return 0.0, rows * C * (4 + 2) + rows * 9
    if n == "sbr_row_gather_bwd_segmented":  # dx rows (fp32) + (sorted key, permutation) per row
        _, rows, C = key
        return 0.0, rows * C * 4 + rows * 8
    if n == "sbr_gather_plan":  # index + modality in, row key / permutation / sorted key out, counts + offsets per key
        _, rows, n_keys = key
        return 0.0, rows * (8 + 1 + 12) + n_keys * 8
    if n == "sbr_bn_apply":
        _, rows, C, o16, o32 = key
        return 0.0, rows * C * (4 + (2 if o16 else 0) + (4 if o32 else 0))
    if n == "sbr_bn_finalize":
        _, parts, C = key
        return 0.0, max(1, parts) * 2 * C * 4 + 6 * C * 4
    if n == "sbr_bn_bwd_reduce":
        _, rows, C = key
        return 0.0, rows * C * 8
    if n == "sbr_bn_bwd_apply":
        _, rows, C, o16, o32 = key
        return 0.0, rows * C * (8 + (2 if o16 else 0) + (4 if o32 else 0))
    if n == "sbr_actgrad_colsum":
        _, rows, C = key
        return 0.0, rows * C * (4 + 4 + 4 + 2)
    if n == "sbr_score_loss":
        _, B, nn, ku, ki, D = key
        return 0.0, 2.0 * B * D * 4 * (ku + nn * ki)
    if n == "sbr_score_loss_bn":
        _, B, nn, D = key
        return 0.0, 2.0 * B * D * 4 * (1 + nn) + B * nn * 4
    if n == "sbr_infonce":  # both modality slots in, their gradients accumulated (read + write)
        _, G, nn, D = key
        return 6.0 * G * nn * nn * D, G * nn * 2 * D * 4 * 3
    if n == "sbr_splitk_reduce":
        _, split, rows, cols = key
        return 0.0, rows * cols * (4.0 * split + 6)  # (the slices exist only because the GEMM was split)
    if n in ("sbr_spmm_csr", "sbr_spmm_csr_bf16"):  # one dense row per stored entry (+ its index), output rows
        _, rows, C = key
        nnz = nnz_hint if nnz_hint else rows
        eb = 2.0 if n.endswith("bf16") else 4.0
        return 2.0 * nnz * C, nnz * (4 + eb * C) + rows * C * 4
    if n == "sbr_adam_step":  # p, g, m, v read; p, m, v, zeroed g written; bf16 shadow
        return 0.0, key[1] * 1024 * 34.0
    if n in ("sbr_tag_bag_fwd", "sbr_tag_bag_bwd"):
        _, rows, C, tags = key
        return 0.0, rows * (C * 4 + tags * 4)
    if n == "sbr_sample_modalities":
        return 0.0, float(key[1])
    if n == "sbr_step_begin":
        return 0.0, float(key[1])
    if n == "sbr_aggregate":
        _, rows, k, D = key
        return 0.0, rows * D * 4 * (k + 1)
    if n in ("sbr_cast_f32_to_bf16", "sbr_transpose_f32"):
        _, r, c = key
        return 0.0, r * c * (6 if n.startswith("sbr_cast") else 8)
    return None


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), "fallback (B200_PROFILING.md)"


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, by kernel family, from the `ncu --set full` captures of
    this round (profiles/r02_dram_traffic.json, written by scripts/ncu_summary.py from the .ncu-rep files)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json")))
    except Exception:
        return {}


def family_rooflines(agg, n_passes, peaks, nnz_hint=None):
    """per kernel family: share of the step's kernel time, algorithmic bytes / flops per launch, achieved rate"""
    fam = {}
    total_ms = sum(v[0] for v in agg.values())
    for key, (ms, cnt) in agg.items():
        f = fam.setdefault(family_of(key), dict(ms=0.0, launches=0, flops=0.0, bytes=0.0, modelled=True, shapes=set()))
        f["ms"] += ms
        f["launches"] += cnt
        w = algorithmic_work(key, nnz_hint)
        if w is None:
            f["modelled"] = False
        else:
            f["flops"] += w[0] * cnt
            f["bytes"] += w[1] * cnt
        f["shapes"].add(str(key[1:]))
    out = []
    ridge = peaks["bf16_tflops_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        dur = f["ms"] * 1e-3
        e = dict(family=name, share=round(f["ms"] / total_ms, 4), launches_per_step=f["launches"] / n_passes,
                 avg_us=round(f["ms"] / f["launches"] * 1e3, 2))
        if f["modelled"] and dur > 0 and (f["flops"] > 0 or f["bytes"] > 0):
            tensor = f["flops"] > 0 and f["flops"] / max(1.0, f["bytes"]) > ridge
            if tensor:  # timed inside a long step: the sustained figure
                ach, peak, unit = f["flops"] / dur / 1e12, peaks["bf16_tflops_sustained"], "TFLOP/s"
            else:
                ach, peak, unit = f["bytes"] / dur / 1e9, peaks["hbm_gbs"], "GB/s"
            e.update(bound="tensor" if tensor else "hbm", achieved=ach, peak=peak, unit=unit, frac=ach / peak,
                     algorithmic_bytes_per_launch=f["bytes"] / f["launches"],
                     algorithmic_flops_per_launch=f["flops"] / f["launches"])
        e["shapes"] = sorted(f["shapes"])[:4]
        out.append(e)
    return out, total_ms / n_passes


# ------------------------------------------------------------------------------------------------ one train workload
def run_train_workload(name, args, env, batch, steps, warmup, full, scale=1.0):
    """times `steps` fused train steps of BASELINE workload `name` on this rank's GPU.  Returns (result dict, objects)."""
    import torch
    import torch.distributed as dist
    from sibrar_b200 import _lib, ops, workloads
    from sibrar_b200.sbnet import SingleBranchNet
    from sibrar_b200.trainer import FusedTrainer
    rank, world, dev = env["rank"], env["world"], env["dev"]
    corpus, conf, learn, _, text = workloads.build(name, scale=scale)
    train = corpus.dataset("train")
    torch.manual_seed(1234)
    model = SingleBranchNet.build_from_conf(conf, train).to(dev).train()
    if world > 1:
        from sibrar_b200.parallel import DataParallelTrainer
        tr = DataParallelTrainer(model, learn, n_negative_samples=N_NEG, cuda_graph=not args.no_graph)
    else:
        tr = FusedTrainer(model, learn, n_negative_samples=N_NEG, cuda_graph=not args.no_graph)

    # ---- synthetic batches, sampled on the device by the GPU sampler (resident in HBM before the timed region)
    B, n = batch, 1 + N_NEG
    coo = train.interaction_matrix
    d = lambda a, t: torch.from_numpy(np.ascontiguousarray(a).astype(t)).to(dev)  # noqa: E731
    csr = train.user_sampling_matrix_train
    coo_u, coo_i = d(coo.row, np.int32), d(coo.col, np.int32)
    ip, ix, items = d(csr.indptr, np.int64), d(csr.indices, np.int32), d(train.items_in_split, np.int32)
    n_batches = min(steps + warmup, 32)
    sample_step = torch.zeros(1, dtype=torch.int64, device=dev)
    batches = []
    for b in range(n_batches):
        ops.tick(sample_step)
        u = torch.empty(B, dtype=torch.int64, device=dev)
        i = torch.empty((B, n), dtype=torch.int64, device=dev)
        ops.sample_batch(coo_u, coo_i, ip, ix, items, B, N_NEG, 1000 + rank, sample_step, u, i)
        batches.append((u, i))
    flush = env["flush"]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(warmup):
        tr.step(*batches[w % n_batches])
    sync_all()
    model.check_errors()
    tr.read_losses()

    # ---- timed region: K steps, CUDA events on the launching stream, L2 flushed between steps (outside the events)
    _lib.reset_launch_counter()
    clocks = ClockSampler(env["local_rank"])
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sync_all()
    t_wall = time.perf_counter()
    for k in range(steps):
        flush.zero_()
        ev[k][0].record()
        tr.step(*batches[(warmup + k) % n_batches])
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
    res = dict(value=world * B * steps / (total_ms * 1e-3), unit=UNIT, ms_per_step=total_ms / steps, steps=steps,
               warmup=warmup, batch_per_gpu=B, global_batch=world * B, gpu_launches=launches,
               train_loss=losses.get("train/loss"), reg_loss=losses.get("train/reg_loss"), clocks=clk,
               collective=getattr(tr, "collective", None),
               wall_ms_per_step_incl_flush=t_wall / steps * 1e3, workload=text,
               params=int(sum(p.numel() for p in model.parameters())))

    # ---- e2e: host buffers in, loss out, through the public trainer API
    # (double-buffered: the pinned-memory fill and the H2D copy of batch k+1 run on a copy stream while step k
    # computes; every step is followed by a device -> host read of its losses, waited for one step later)
    hu = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(2)]
    hi = [torch.empty((B, n), dtype=torch.int64).pin_memory() for _ in range(2)]
    host_batches = [(b[0].cpu(), b[1].cpu()) for b in batches[:4]]
    dbuf = [(torch.empty_like(batches[0][0]), torch.empty_like(batches[0][1])) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = [torch.empty(tr.loss_acc.shape, dtype=tr.loss_acc.dtype).pin_memory() for _ in range(2)]
    loss_read = [torch.cuda.Event() for _ in range(2)]

    def stage(k):
        s_ = k % 2
        src = host_batches[k % len(host_batches)]
        if k >= 2:
            ready[s_].synchronize()  # the H2D copy that last read this pinned buffer has finished (host runs one step ahead)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s_])  # the step that last read this device buffer has finished
            hu[s_].copy_(src[0])
            hi[s_].copy_(src[1])
            dbuf[s_][0].copy_(hu[s_], non_blocking=True)
            dbuf[s_][1].copy_(hi[s_], non_blocking=True)
            ready[s_].record(copy_stream)

    def e2e_loop(n_steps):
        stage(0)
        for k in range(n_steps):
            torch.cuda.current_stream().wait_event(ready[k % 2])
            tr.step(*dbuf[k % 2])
            consumed[k % 2].record()
            if k + 1 < n_steps:
                stage(k + 1)
            # device -> host read of THIS step's losses, every step; the host waits for it after the NEXT step has been
            # enqueued (a logging loop has no use for the value before that), so the launch of step k+1 overlaps step k
            loss_host[k % 2].copy_(tr.loss_acc, non_blocking=True)
            loss_read[k % 2].record()
            if k > 0:
                loss_read[(k - 1) % 2].synchronize()
        loss_read[(n_steps - 1) % 2].synchronize()

    sync_all()
    for e in consumed:
        e.record()
    e2e_loop(max(3, warmup))  # untimed: first use of the copy stream, the events and the pinned buffers
    sync_all()
    for e in consumed:
        e.record()
    t0 = time.perf_counter()
    e2e_loop(steps)
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["e2e"] = dict(value=world * B * steps / float(t.item()), unit=UNIT, h2d_bytes_per_step=B * 8 + B * n * 8,
                      d2h_bytes_per_step=int(tr.loss_acc.numel() * 8))
    tr.read_losses()

    # ---- kernel families + roofline (separate pass, CUDA events around every C-ABI call, one stream, no graph); every
    # rank runs these steps (they contain the gradient all-reduce), rank 0 records them
    graph, branches = tr.cuda_graph, tr.branches
    tr.cuda_graph, tr.branches = False, 0
    n_prof = 3
    nnz_hint = float(train.user_sampling_matrix_train.nnz)
    if rank != 0:
        for k in range(n_prof):
            flush.zero_()
            torch.cuda._sleep(20_000_000)
            tr.step(*batches[(warmup + k) % n_batches])
    else:
        peaks, which = load_peaks()
        with CallProfiler(ops, torch) as prof:
            for k in range(n_prof):
                flush.zero_()
                # the host needs longer to launch the kernels of an eager step than the GPU to run them: a spin kernel
                # in front lets the host run ahead, so that the events bracket kernel time, not launch gaps
                torch.cuda._sleep(20_000_000)
                tr.step(*batches[(warmup + k) % n_batches])
        fams, serial_ms = family_rooflines(prof.summary(), n_prof, peaks, nnz_hint)
        res["kernel_families"] = fams[:10] if full else fams[:5]
        res["serialised_kernel_ms_per_step"] = serial_ms
        traffic = load_ncu_traffic()
        top = fams[0]  # the dominant family is the one with the largest share, modelled or not
        roof = dict(kernel=top["family"], share_of_step=top["share"], launches_per_step=top["launches_per_step"],
                    avg_us=top["avg_us"], peak_source=which)
        if "bound" in top:
            roof.update(bound=top["bound"], achieved=top["achieved"], peak=top["peak"], unit=top["unit"],
                        frac=top["frac"], algorithmic_bytes_per_launch=top["algorithmic_bytes_per_launch"],
                        algorithmic_flops_per_launch=top["algorithmic_flops_per_launch"])
        t_ = traffic.get(name, {}).get(top["family"])
        roof["traffic"] = t_["dram_bytes_per_launch"] if t_ else None
        if t_:
            roof["traffic_source"] = t_.get("source")
        res["roofline"] = roof
    tr.cuda_graph, tr.branches = graph, branches
    tr.read_losses()
    return res, dict(corpus=corpus, conf=conf, learn=learn, model=model, trainer=tr, batches=batches)


def release(objs):
    import torch
    tr = objs.get("trainer")
    if tr is not None:
        tr._graphs.clear()
    objs.clear()
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="ml1m", choices=["ml1m", "onion18_huge", "amazon_nouser"],
                    help="headline workload (BASELINE.json configs[1] / [2] / [3])")
    ap.add_argument("--extra-configs", default=os.environ.get("SBR_BENCH_EXTRA", "onion18_huge,amazon_nouser"),
                    help="comma-separated workloads measured next to the headline ('' = none)")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("SBR_BENCH_BATCH", 16384)),
                    help="interactions per GPU per step")
    ap.add_argument("--scale", type=float, default=1.0, help="corpus scale (debugging)")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel (no CUDA graph)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    import sibrar_b200  # noqa: F401
    from sibrar_b200 import workloads

    if args.impl == "reference":
        if rank != 0:
            return
        corpus, conf, learn, _, text = workloads.build(args.config, scale=args.scale)
        cb, s_per_step = cpu_train_arm(corpus, conf, learn, args.batch, args.steps, args.warmup, budget_s=150.0)
        small, _ = cpu_train_arm(corpus, conf, learn, 256, 10 ** 6, 2, budget_s=15.0)
        line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=s_per_step * 1e3, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                    config=dict(workload=text, batch_per_gpu=args.batch, n_neg=N_NEG,
                                note="CPU arm: bounded sample of the same workload at the same batch size, rank 0 only"),
                    cpu_baseline=cb, paper_batch=dict(batch=256, value=small["value"], unit=UNIT, sample=small["sample"]),
                    e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from sibrar_b200 import _lib, ops
    from sibrar_b200.evaluator import FullEvaluator

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    _lib.lib()  # fail loudly here if the CUDA extension is missing
    env = dict(rank=rank, world=world, local_rank=local_rank, dev=dev,
               flush=torch.empty(256 << 20, dtype=torch.uint8, device=dev))  # > 126 MB L2

    head, objs = run_train_workload(args.config, args, env, args.batch, args.steps, args.warmup, full=True,
                                    scale=args.scale)
    line = dict(metric=METRIC, value=head["value"], unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=head["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="bf16", data="synthetic",
                config=dict(workload=head["workload"], name=args.config, batch_per_gpu=args.batch, n_neg=N_NEG,
                            global_batch=world * args.batch, parallelism=f"dp{world}", params=head["params"],
                            collective=head.get("collective"),
                            l2="flushed between timed steps (256 MiB write)",
                            wall_ms_per_step_incl_flush=head["wall_ms_per_step_incl_flush"]),
                clocks=head["clocks"], e2e=head["e2e"], gpu_launches=head["gpu_launches"],
                train_loss=head["train_loss"])
    for k in ("kernel_families", "serialised_kernel_ms_per_step", "roofline"):
        if k in head:
            line[k] = head[k]

    # ---- evaluation of the headline workload's own split (+ item-sharded under torchrun)
    if not args.no_eval:
        try:
            ev = bench_eval_split(torch, objs["model"], objs["corpus"], FullEvaluator, env)
            if rank == 0:
                line["eval"] = ev
        except Exception as e:  # keep the headline line
            line["eval"] = {"error": repr(e)}

    # ---- the headline workload at the reference's batch size (the CPU arm's paper_batch figure is measured at the same B)
    try:
        release(objs)
        small, o2 = run_train_workload(args.config, args, env, 256, max(10, args.steps), 3, full=False, scale=args.scale)
        release(o2)
        line["paper_batch"] = dict(batch=256, value=small["value"], unit=UNIT, ms_per_step=small["ms_per_step"],
                                   e2e=small["e2e"]["value"])
    except Exception as e:
        line["paper_batch"] = {"error": repr(e)}

    # ---- the other BASELINE train configs
    extras = [c for c in args.extra_configs.split(",") if c and c != args.config]
    if extras:
        line["configs"] = {}
    for name in extras:
        try:
            r, o = run_train_workload(name, args, env, args.batch, max(5, args.steps // 2), 3, full=False,
                                      scale=args.scale)
            if not args.no_eval:
                try:
                    r["eval"] = bench_eval_split(torch, o["model"], o["corpus"], FullEvaluator, env, graph=False)
                except Exception as e:
                    r["eval"] = {"error": repr(e)}
            release(o)
            r.pop("wall_ms_per_step_incl_flush", None)
            line["configs"][name] = r
        except Exception as e:
            line["configs"][name] = {"error": repr(e)}
            gc.collect()
            torch.cuda.empty_cache()

    # ---- full-catalog top-k sweep (configs[4]) + its roofline + the CPU evaluation baseline
    if not args.no_eval:
        try:
            sw = bench_eval_sweep(torch, ops, env)
            if rank == 0:
                line.setdefault("eval", {}).update(sw)
        except Exception as e:
            line.setdefault("eval", {})["sweep_error"] = repr(e)
    if rank == 0:
        if not args.no_cpu_baseline:
            try:
                corpus, conf, learn, _, _ = workloads.build(args.config, scale=args.scale)
                line["cpu_baseline"], _ = cpu_train_arm(corpus, conf, learn, args.batch, 10 ** 6, 1, budget_s=20.0)
                small, _ = cpu_train_arm(corpus, conf, learn, 256, 10 ** 6, 2, budget_s=8.0)
                line["cpu_baseline"]["paper_batch"] = dict(batch=256, value=small["value"], unit=UNIT)
                if not args.no_eval:
                    line.setdefault("eval", {})["cpu_baseline"] = cpu_eval_arm()
            except Exception as e:
                line["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # graphs that captured NCCL kernels go before the communicator does
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        # a watchdog in case the teardown of a communicator that had captured collectives does not return
        threading.Thread(target=lambda: (time.sleep(30), os._exit(0)), daemon=True).start()
        dist.destroy_process_group()


def bench_eval_split(torch, model, corpus, FullEvaluator, env, graph=True):
    """evaluation of the workload's validation split end to end (all ranks take part when world > 1)"""
    out = {}
    rank, world, dev = env["rank"], env["world"], env["dev"]
    val = corpus.dataset("val")
    conf = dict(top_k=[1, 10, 20], metrics=["ndcg", "recall", "precision", "coverage"], calculate_std=False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    if rank == 0:
        ev = FullEvaluator(conf)
        ev.evaluate(model, val)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            res = ev.evaluate(model, val)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        out["workload_val_split"] = dict(metric="eval_users_per_sec", value=val.n_users_in_split / (ms * 1e-3),
                                         unit="users/s", users=int(val.n_users_in_split),
                                         items=int(val.n_items_in_split), ms=ms,
                                         includes="item+user representations, scores, mask, top-20, metrics",
                                         ndcg10=res.get("ndcg@10"))
        if graph:
            evg = FullEvaluator(conf, cuda_graph=True)
            for _ in range(3):  # eager, capture, first replay
                resg = evg.evaluate(model, val)
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                resg = evg.evaluate(model, val)
            b.record()
            torch.cuda.synchronize()
            msg = a.elapsed_time(b) / reps
            out["workload_val_split_graph"] = dict(metric="eval_users_per_sec", ms=msg, unit="users/s",
                                                   value=val.n_users_in_split / (msg * 1e-3),
                                                   same_result=bool(resg == res),
                                                   includes="the same evaluation replayed as one CUDA graph + one D2H "
                                                            "copy")
    if world > 1:
        import torch.distributed as dist
        from sibrar_b200.parallel import ShardedEvaluator
        sev = ShardedEvaluator(conf)
        res = sev.evaluate(model, val)
        dist.barrier()
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            res = sev.evaluate(model, val)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["item_sharded_val_split"] = dict(metric="eval_users_per_sec", unit="users/s", shards=world,
                                             value=val.n_users_in_split / (float(t.item()) * 1e-3), ms=float(t.item()),
                                             ndcg10=res.get("ndcg@10"))
    return out


SWEEP = ((100_000, 100_000, 64, 10), (100_000, 1_000_000, 64, 10), (100_000, 1_000_000, 128, 50),
         (100_000, 1_000_000, 256, 10), (100_000, 1_000_000, 512, 10), (100_000, 10_000_000, 64, 10),
         (100_000, 10_000_000, 256, 50))  # (the largest point: 5 GB of item embeddings, ~0.5 s on one GPU)


def bench_eval_sweep(torch, ops, env):
    """BASELINE configs[4]: the fused score GEMM + seen mask + top-k kernel on synthetic embeddings (U = 1e5 users,
    I = 1e5 .. 1e7 items, D = 64 .. 512, k = 10 / 50, 100 seen items per user).  With world > 1 the item catalogue is
    sharded: every rank scores its I / world items, the [U, k] key lists are all-gathered and merged."""
    rank, world, dev = env["rank"], env["world"], env["dev"]
    peaks, which = load_peaks()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sweep, best = [], None
    for U, I, D, k in SWEEP:
        g = torch.Generator(device=dev).manual_seed(7)  # same embeddings on every rank
        u16 = (torch.randn(U, D, generator=g, device=dev) / math.sqrt(D)).to(torch.bfloat16)
        lo, hi = (0, I)
        if world > 1:
            from sibrar_b200.parallel import shard_range
            lo, hi = shard_range(I, rank, world)
        gi = torch.Generator(device=dev).manual_seed(1000 + lo)
        i16 = torch.randn(hi - lo, D, generator=gi, device=dev).to(torch.bfloat16)
        seen_ip = torch.arange(0, (U + 1) * 100, 100, dtype=torch.int64, device=dev)
        seen_all = torch.sort(torch.randint(0, I, (U, 100), device=dev, dtype=torch.int32, generator=g), dim=1).values
        if world > 1:  # this shard's slice of the seen CSR, positions relative to the shard
            inside = (seen_all >= lo) & (seen_all < hi)
            seen_ip = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), inside.sum(1).cumsum(0)])
            seen_ix = (seen_all[inside] - lo).to(torch.int32).contiguous()
        else:
            seen_ix = seen_all.reshape(-1).contiguous()

        def one():
            if world == 1:
                return ops.topk_scores_masked(u16, i16, U, I, D, seen_ip, seen_ix, k)
            import torch.distributed as dist
            local = ops.topk_scores_masked(u16, i16, U, hi - lo, D, seen_ip, seen_ix, k, item_offset=lo,
                                           return_keys=True)
            gathered = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=dev)
            dist.all_gather_into_tensor(gathered, local.contiguous())
            return ops.topk_merge(gathered, world, U, k)
        one()
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        a.record()
        one()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tf = 2.0 * U * I * D / (ms * 1e-3) / 1e12
        e = dict(U=U, I=I, D=D, k=k, ms=ms, users_per_s=U / (ms * 1e-3), tflops=tf, shards=world,
                 frac_of_bf16_burst=tf / (world * peaks["bf16_tflops"]),
                 frac_of_bf16_sustained=tf / (world * peaks["bf16_tflops_sustained"]))
        sweep.append(e)
        if best is None or tf > best["tflops"]:
            best = e
        del u16, i16, seen_ix, seen_all
        torch.cuda.empty_cache()
    out = {"sweep": sweep}
    if best is not None:  # the kernel is timed alone (one launch + finalize + merge): the burst figure
        out["roofline"] = dict(kernel="sbr_topk_scores_masked", bound="tensor", unit="TFLOP/s",
                               achieved=best["tflops"] / world, peak=peaks["bf16_tflops"],
                               frac=best["tflops"] / world / peaks["bf16_tflops"], peak_source=which + " burst",
                               point=dict(U=best["U"], I=best["I"], D=best["D"], k=best["k"]),
                               algorithmic_flops_per_launch=2.0 * best["U"] * best["I"] * best["D"] / world,
                               traffic=load_ncu_traffic().get("evalsweep", {}).get("sbr_topk_scores_masked", {}).get(
                                   "dram_bytes_per_launch"))
    return out


if __name__ == "__main__":
    main()
