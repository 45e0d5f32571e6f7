"""Thin tensor-level wrappers over the C ABI (``include/sibrar_b200.h``).  torch is used for device memory and the
current stream only; every function enqueues hand-written sm_100a kernels and returns immediately."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import (ACT, LOSS, AdamTensor, BnInline, GemmEpilogue, Mlp2Bn, Mlp2Desc, ModalitySrc, call, ptr,
                   stream_ptr)

BF16, F32 = torch.bfloat16, torch.float32


def pad8(n: int) -> int:
    return (int(n) + 7) // 8 * 8


def _act(a):
    return a if isinstance(a, int) else ACT[a]


def _epilogue(bias, act, out_bf16, out_f32, colstats, colstats_sum_only, actgrad_y, actgrad_act, transpose_out,
              atomic_out, split_k, alpha, split_stride=0, colstats_rows=0):
    ep = GemmEpilogue()
    ep.bias = ptr(bias)
    ep.act = _act(act)
    ep.out_bf16 = ptr(out_bf16)
    ep.ld_bf16 = out_bf16.stride(0) if out_bf16 is not None else 0
    ep.out_f32 = ptr(out_f32)
    ep.ld_f32 = out_f32.stride(0) if out_f32 is not None else 0
    ep.colstats = ptr(colstats)
    ep.colstats_sum_only = int(colstats_sum_only)
    ep.colstats_rows = int(colstats_rows)
    ep.actgrad_y = ptr(actgrad_y)
    ep.ld_actgrad = actgrad_y.stride(0) if actgrad_y is not None else 0
    ep.actgrad_act = _act(actgrad_act)
    ep.transpose_out = int(transpose_out)
    ep.atomic_out = int(atomic_out)
    ep.split_k = int(split_k)
    ep.split_stride = int(split_stride)
    ep.alpha = float(alpha)
    return ep


def gemm(A, B, M, N, K, *, a_mn=False, b_mn=False, lda=None, ldb=None, bias=None, act=None, out_bf16=None,
         out_f32=None, colstats=None, colstats_sum_only=False, actgrad_y=None, actgrad_act=None, transpose_out=False, atomic_out=False,
         split_k=1, alpha=1.0, split_stride=0, colstats_rows=0):
    """D[M,N] = alpha * A @ B^T (bf16 in, fp32 accumulate) + fused epilogue.  A/B: 2-D bf16 tensors whose last
    dimension is contiguous; K-major means [rows, K], MN-major means [K, rows]."""
    assert A.dtype == BF16 and B.dtype == BF16 and A.stride(-1) == 1 and B.stride(-1) == 1
    ep = _epilogue(bias, act, out_bf16, out_f32, colstats, colstats_sum_only, actgrad_y, actgrad_act, transpose_out,
                   atomic_out, split_k, alpha, split_stride, colstats_rows)
    call("sbr_gemm_bf16", ptr(A), lda if lda is not None else A.stride(0), int(a_mn), ptr(B),
         ldb if ldb is not None else B.stride(0), int(b_mn), int(M), int(N), int(K), C.byref(ep), stream_ptr())


def gemm_bits(A_bits, B, M, N, K, *, b_mn=False, bias=None, act=None, out_bf16=None, out_f32=None, colstats=None,
              colstats_sum_only=False, transpose_out=False, atomic_out=False, split_k=1, alpha=1.0, split_stride=0,
              colstats_rows=0):
    """the same with a bit-packed 0/1 A operand: int32 [M, ld_words], bit k of row m = word k // 32, bit k % 32"""
    assert A_bits.dtype == torch.int32 and A_bits.stride(-1) == 1 and B.dtype == BF16 and B.stride(-1) == 1
    ep = _epilogue(bias, act, out_bf16, out_f32, colstats, colstats_sum_only, None, None, transpose_out, atomic_out,
                   split_k, alpha, split_stride, colstats_rows)
    call("sbr_gemm_bits_bf16", ptr(A_bits), A_bits.stride(0), ptr(B), B.stride(0), int(b_mn), int(M), int(N), int(K),
         C.byref(ep), stream_ptr())


def gemm_colstats_rows(M: int, N: int) -> int:
    """rows of partial column statistics a [M, N] GEMM writes in deterministic mode (4 epilogue warps x CTAs)"""
    from ._lib import lib
    return int(lib().sbr_gemm_colstats_rows(int(M), int(N)))


def splitk_reduce(partials, n_splits, rows, cols, bias=None, act=None, out_f32=None, out_bf16=None, accumulate=False):
    """partials: fp32 [n_splits, rows, cols] written by a split-K GEMM with ``split_stride = rows * cols``"""
    call("sbr_splitk_reduce", ptr(partials), int(n_splits), int(rows) * int(cols), int(cols), int(rows), int(cols),
         ptr(bias), _act(act), ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0, int(bool(accumulate)),
         ptr(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0, stream_ptr())


SPLITK_MAX_PARTIAL_BYTES = 256 << 20


def effective_splits(K: int, split_k: int) -> int:
    """number of K partitions the GEMM really uses (no empty partition): mirrors sbr_gemm_bf16"""
    num_kb = -(-int(K) // 64)
    split_k = max(1, min(int(split_k), num_kb))
    per = -(-num_kb // split_k)
    return -(-num_kb // per)


def pack_bits(csr, device) -> torch.Tensor:
    """scipy CSR 0/1 matrix -> int32 [rows, ld_words] bit matrix (16-byte rows, whole 64-bit K blocks, zero padded)"""
    import numpy as np
    rows, cols = csr.shape
    ld_words = 4 * ((cols + 127) // 128)  # 16-byte row pitch: the kernel fetches the bit words by TMA
    out = np.zeros((rows, ld_words), dtype=np.uint32)
    coo = csr.tocoo()
    np.bitwise_or.at(out, (coo.row, coo.col // 32), (np.uint32(1) << (coo.col % 32).astype(np.uint32)))
    return torch.from_numpy(out.view(np.int32)).to(device)


def cast_bf16(src: torch.Tensor, dst: torch.Tensor = None) -> torch.Tensor:
    """fp32 [rows, cols] -> bf16 [rows, pad8(cols)] (pad columns zeroed)"""
    rows, cols = src.shape
    if dst is None:
        dst = torch.empty((rows, pad8(cols)), dtype=BF16, device=src.device)
    call("sbr_cast_f32_to_bf16", ptr(src), src.stride(0), ptr(dst), dst.stride(0), rows, cols, stream_ptr())
    return dst


def transpose_bf16(src: torch.Tensor, dst: torch.Tensor = None) -> torch.Tensor:
    """fp32 [rows, cols] -> bf16 [cols, pad8(rows)] (dst[c, r] = src[r, c]; pad columns must be pre-zeroed by the caller)"""
    rows, cols = src.shape
    if dst is None:
        dst = torch.zeros((cols, pad8(rows)), dtype=BF16, device=src.device)
    call("sbr_transpose_f32_to_bf16", ptr(src), src.stride(0), ptr(dst), dst.stride(0), rows, cols, stream_ptr())
    return dst


def transpose_f32(src: torch.Tensor, dst: torch.Tensor = None) -> torch.Tensor:
    rows, cols = src.shape
    if dst is None:
        dst = torch.empty((cols, rows), dtype=F32, device=src.device)
    call("sbr_transpose_f32", ptr(src), src.stride(0), ptr(dst), dst.stride(0), rows, cols, stream_ptr())
    return dst


def csr_to_dense_bf16(indptr, indices, rows, cols, vals=None) -> torch.Tensor:
    """``vals``: fp32 stored values (None: every stored entry is 1)"""
    dst = torch.empty((rows, pad8(cols)), dtype=BF16, device=indptr.device)
    call("sbr_csr_to_dense_bf16", ptr(indptr), ptr(indices), ptr(vals), rows, cols, ptr(dst), dst.stride(0),
         stream_ptr())
    return dst


def spmm_csr(indptr, indices, rows, dense, C_, bias, act, out, transpose_out=False, vals=None, accumulate=False,
             out_bf16=None, row_map=None, atomic=False, row_list=None, n_rows_dev=None, segments=None, out_pos=None):
    """out[r] = act(sum_p vals[p] * dense[indices[p]] + bias); ``transpose_out``: out is [C, rows] (the wgrad through
    the transposed CSR), ``accumulate``: out += result.  ``dense`` fp32 or bf16 (fp32 accumulation either way).
    fp32 only: ``row_map`` / ``atomic`` (segment mode).  bf16 only: ``row_list`` / ``n_rows_dev`` (row subset)."""
    if dense.dtype == BF16:
        assert row_map is None and not atomic
        seg_ptr = seg_row = long_rows = None
        if segments is not None:  # (seg_ptr int64 [n_seg + 1], seg_row int32 [n_seg], long_rows int32 [n_long])
            seg_ptr, seg_row, long_rows = segments[:3]
            indptr = seg_ptr
            rows = seg_row.numel() if row_list is None else rows
            if not transpose_out and long_rows.numel() > 0 and out_pos is None:
                from . import _lib
                _lib._launches[0] += 2  # the clear / fix-up passes over the long rows
        call("sbr_spmm_csr_bf16", ptr(indptr), ptr(indices), ptr(vals), int(rows), ptr(dense), dense.stride(0),
             int(C_), ptr(bias), _act(act), ptr(out), out.stride(0) if out is not None else 0, int(transpose_out),
             int(bool(accumulate)), ptr(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0, ptr(row_list),
             ptr(n_rows_dev), ptr(seg_row), ptr(long_rows), long_rows.numel() if long_rows is not None else 0,
             ptr(out_pos), stream_ptr())
        return
    assert segments is None
    assert row_list is None and n_rows_dev is None
    call("sbr_spmm_csr", ptr(indptr), ptr(indices), ptr(vals), int(rows), ptr(dense), dense.stride(0), int(C_),
         ptr(bias), _act(act), ptr(out), out.stride(0) if out is not None else 0, int(transpose_out),
         int(bool(accumulate)), ptr(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0, ptr(row_map),
         int(bool(atomic)), stream_ptr())


def sample_modalities(mods, n_rows, k, n_mods, central, seed, step_dev):
    call("sbr_sample_modalities", ptr(mods), int(n_rows), int(k), int(n_mods), int(central), int(seed),
         ptr(step_dev), stream_ptr())


def tick(counter):
    call("sbr_tick", ptr(counter), stream_ptr())


def step_begin(counter0, counter1, zero_buf, zero_bytes):
    """counters += 1 and the first ``zero_bytes`` of ``zero_buf`` cleared, in one launch"""
    call("sbr_step_begin", ptr(counter0), ptr(counter1), ptr(zero_buf), int(zero_bytes), stream_ptr())


def make_modality_srcs(entries, device) -> torch.Tensor:
    """entries: list of dict(kind, remap, table, grad, codes, max_tags, pad_id) -> uint8 device blob of
    ``sbr_modality_src_t[n]``"""
    arr = (ModalitySrc * len(entries))()
    for i, e in enumerate(entries):
        arr[i].kind = e["kind"]
        arr[i].remap = ptr(e.get("remap"))
        arr[i].table = ptr(e.get("table"))
        arr[i].grad = ptr(e.get("grad"))
        arr[i].codes = ptr(e.get("codes"))
        arr[i].max_tags = int(e.get("max_tags", 0))
        arr[i].pad_id = int(e.get("pad_id", -1))
        arr[i].n_table_rows = int(e["table"].shape[0]) if e.get("table") is not None else 0
        arr[i].key_base = int(e.get("key_base", 0))
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(device)


def row_gather_fwd(srcs, n_mods, idx, mods, k, C_, normalize, p_drop, seed, step_dev, keep_mask, out_bf16=None,
                   out_f32=None, err_flag=None, keep_bits_out=None):
    call("sbr_row_gather_fwd", ptr(srcs), int(n_mods), ptr(idx), ptr(mods), idx.numel(), int(k), int(C_),
         int(bool(normalize)), float(p_drop or 0.0), int(seed), ptr(step_dev), ptr(keep_mask), ptr(out_bf16),
         out_bf16.stride(0) if out_bf16 is not None else 0, ptr(out_f32),
         out_f32.stride(0) if out_f32 is not None else 0, ptr(err_flag), ptr(keep_bits_out), stream_ptr())


def row_gather_bwd(srcs, n_mods, idx, mods, k, C_, normalize, p_drop, seed, step_dev, keep_mask, dx):
    call("sbr_row_gather_bwd", ptr(srcs), int(n_mods), ptr(idx), ptr(mods), idx.numel(), int(k), int(C_),
         int(bool(normalize)), float(p_drop or 0.0), int(seed), ptr(step_dev), ptr(keep_mask), ptr(dx), dx.stride(0),
         stream_ptr())


class GatherPlan:
    """scratch + launches of the sorted-run gather backward for a fixed (n_keys, N)"""

    def __init__(self, n_keys: int, N: int, device, rows_per_warp: int = 8):
        i32 = torch.int32
        self.n_keys, self.N = int(n_keys), int(N)
        self.counts = torch.zeros(n_keys, dtype=i32, device=device)
        self.cursor = torch.zeros(n_keys, dtype=i32, device=device)
        # counts: zero on entry, cleared again by the plan's scan; offsets: + the scan's per-chunk state (include/sibrar_b200.h)
        self.offsets = torch.zeros(n_keys + 1 + (n_keys + 4095) // 4096, dtype=i32, device=device)
        self.row_keys = torch.empty(N, dtype=i32, device=device)
        self.perm = torch.empty(N, dtype=i32, device=device)
        self.sorted_keys = torch.empty(N, dtype=i32, device=device)
        self.rows_per_warp = rows_per_warp

    def build(self, srcs, n_mods, idx, mods, k):
        assert idx.numel() * k == self.N
        call("sbr_gather_plan", ptr(srcs), int(n_mods), ptr(idx), ptr(mods), idx.numel(), int(k), self.n_keys,
             ptr(self.counts), ptr(self.offsets), ptr(self.cursor), ptr(self.row_keys), ptr(self.perm),
             ptr(self.sorted_keys), stream_ptr())

    def backward(self, srcs, n_mods, C_, normalize, p_drop, seed, step_dev, keep_mask, dx, keep_bits=None):
        call("sbr_row_gather_bwd_segmented", ptr(srcs), int(n_mods), self.n_keys, ptr(self.offsets), ptr(self.perm),
             ptr(self.sorted_keys), self.N, int(C_), int(bool(normalize)), float(p_drop or 0.0), int(seed),
             ptr(step_dev), ptr(keep_mask), ptr(dx), dx.stride(0), self.rows_per_warp, ptr(keep_bits), stream_ptr())


TAG_BAG_SMEM_FLOATS = 10240  # csrc/gather.cu SEG_SMEM_FLOATS: gradient matrices up to this size are privatised per block


def make_ref_tables(entries, device) -> torch.Tensor:
    """entries: per modality None (whole-table route) or dict(stamp, pos, list, count[, seg_first, seg_list, seg_count])
    -> uint8 device blob of ``sbr_ref_table_t[n]``"""
    from ._lib import RefTable
    arr = (RefTable * len(entries))()
    for i, e in enumerate(entries):
        if e is None:
            continue
        arr[i].stamp, arr[i].pos, arr[i].list, arr[i].count = ptr(e["stamp"]), ptr(e["pos"]), ptr(e["list"]), ptr(e["count"])
        arr[i].seg_first, arr[i].seg_list = ptr(e.get("seg_first")), ptr(e.get("seg_list"))
        arr[i].seg_count = ptr(e.get("seg_count"))
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)


def mark_referenced(srcs, n_mods, idx, mods, k, epoch_dev, tabs):
    call("sbr_mark_referenced", ptr(srcs), int(n_mods), ptr(idx), ptr(mods), idx.numel(), int(k), ptr(epoch_dev),
         ptr(tabs), stream_ptr())


def gather_rows_bf16(src, lst, count_dev, capacity, dst):
    """dst[slot] = src[lst[slot]] for slot < *count_dev, zero rows up to ``capacity``"""
    assert src.dtype == BF16 and dst.dtype == BF16 and src.shape[1] == dst.shape[1]
    call("sbr_gather_rows_bf16", ptr(src), src.stride(0), ptr(lst), ptr(count_dev), int(capacity), int(src.shape[1]),
         ptr(dst), dst.stride(0), stream_ptr())


def spmm_scatter_wgrad(indptr, indices, vals, unit_list, n_units_dev, max_units, seg_row, pos, dz16, C_, gT):
    call("sbr_spmm_scatter_wgrad", ptr(indptr), ptr(indices), ptr(vals), ptr(unit_list), ptr(n_units_dev),
         int(max_units), ptr(seg_row), ptr(pos), ptr(dz16), dz16.stride(0), int(C_), ptr(gT), gT.stride(0),
         stream_ptr())


def transpose_add_f32(src, dst):
    """dst[c, r] += src[r, c]; src cleared"""
    rows, cols = src.shape
    call("sbr_transpose_add_f32", ptr(src), src.stride(0), ptr(dst), dst.stride(0), rows, cols, stream_ptr())


def mlp2_desc(srcs, n_mods, idx, mods, k, C_, normalize, p_drop, seed, step_dev, keep_mask, err_flag, layers):
    """descriptor of the fused gather + single-branch MLP kernels; ``layers``: list (1 or 2) of
    (w_bf16 [out, ldw], bias or None, in_f, out_f, activation applied to the layer's output)"""
    d = Mlp2Desc()
    d.srcs, d.n_mods, d.idx, d.mods = ptr(srcs), int(n_mods), ptr(idx), ptr(mods)
    d.n_idx, d.k, d.C, d.normalize = idx.numel(), int(k), int(C_), int(bool(normalize))
    d.p_drop, d.seed, d.step_dev = float(p_drop or 0.0), int(seed), ptr(step_dev)
    d.keep_mask, d.err_flag, d.n_layers = ptr(keep_mask), ptr(err_flag), len(layers)
    for i, (w16, bias, in_f, out_f, act) in enumerate(layers):
        assert w16.dtype == BF16 and w16.stride(-1) == 1
        y = d.layers[i]
        y.w_bf16, y.ldw, y.bias = ptr(w16), w16.stride(0), ptr(bias)
        y.in_f, y.out_f, y.act = int(in_f), int(out_f), _act(act)
    d._keep = (srcs, idx, mods, step_dev, keep_mask, err_flag, [(w, b) for w, b, *_ in layers])  # keep-alive
    return d


def mlp2_colstats_rows(n_rows: int) -> int:
    from ._lib import lib
    return int(lib().sbr_mlp2_colstats_rows(int(n_rows)))


def mlp2_fwd(desc, n_rows, C_, z, colstats=None, colstats_rows=0, bn_tail=None):
    """bn_tail: None or dict(counter (uint32 [1], zero), eps, momentum, mean_invstd, running_mean, running_var,
    num_batches_tracked): the BatchNorm statistics are finalised by the last CTA of the kernel"""
    if bn_tail is None:
        call("sbr_mlp2_fwd", C.byref(desc), int(n_rows), int(C_), ptr(z), z.stride(0), ptr(colstats),
             int(colstats_rows), stream_ptr())
        return
    from ._lib import Mlp2BnTail
    t = Mlp2BnTail()
    t.counter, t.eps, t.momentum = ptr(bn_tail["counter"]), float(bn_tail["eps"]), float(bn_tail["momentum"])
    t.mean_invstd, t.running_mean = ptr(bn_tail["mean_invstd"]), ptr(bn_tail.get("running_mean"))
    t.running_var, t.num_batches_tracked = ptr(bn_tail.get("running_var")), ptr(bn_tail.get("num_batches_tracked"))
    call("sbr_mlp2_fwd_bn", C.byref(desc), int(n_rows), int(C_), ptr(z), z.stride(0), ptr(colstats),
         int(colstats_rows), C.byref(t), stream_ptr())


def mlp2_bwd(desc, n_rows, C_, dy, z, bn, grad_w, grad_b, dx):
    """bn: None or dict(mean_invstd, gamma, sums, n_replicas, dgamma, dbeta); grad_w / grad_b: per-layer fp32 tensors
    (None entries are skipped)"""
    b = None
    if bn is not None:
        b = Mlp2Bn()
        b.mean_invstd, b.gamma, b.sums = ptr(bn["mean_invstd"]), ptr(bn["gamma"]), ptr(bn["sums"])
        b.n_replicas, b.dgamma, b.dbeta = int(bn["n_replicas"]), ptr(bn.get("dgamma")), ptr(bn.get("dbeta"))
    from ._lib import c_vp
    gw = (c_vp * 2)(*[ptr(t) for t in (list(grad_w) + [None, None])[:2]])
    gb = (c_vp * 2)(*[ptr(t) for t in (list(grad_b) + [None, None])[:2]])
    call("sbr_mlp2_bwd", C.byref(desc), int(n_rows), int(C_), ptr(dy), dy.stride(0), ptr(z), z.stride(0),
         C.byref(b) if b is not None else None, gw, gb, ptr(dx), dx.stride(0), stream_ptr())


def tag_bag_fwd(codes, max_tags, pad_id, weight, out):
    """EmbeddingBag(mean, padding_idx) of every feature row -> fp32 table [n_rows, C]"""
    call("sbr_tag_bag_fwd", ptr(codes), int(max_tags), int(pad_id), ptr(weight), int(out.shape[0]), int(out.shape[1]),
         ptr(out), stream_ptr())


def tag_bag_bwd(codes, max_tags, pad_id, bag_grad, grad_weight):
    """table gradient [n_rows, C] (cleared) -> += gradient of the tag embedding matrix"""
    call("sbr_tag_bag_bwd", ptr(codes), int(max_tags), int(pad_id), ptr(bag_grad), int(bag_grad.shape[0]),
         int(bag_grad.shape[1]), ptr(grad_weight), int(grad_weight.shape[0]), stream_ptr())


def _y_args(y):
    if y is None:
        return None, None, 0
    return (ptr(y), None, y.stride(0)) if y.dtype == F32 else (None, ptr(y), y.stride(0))


def actgrad_colsum(dy, y, act, rows, cols, out_bf16=None, out_f32=None, colsum=None, zero_dy=False):
    yf, yb, ldy = _y_args(y)
    call("sbr_actgrad_colsum", ptr(dy), dy.stride(0), yf, yb, ldy, _act(act), int(rows), int(cols), ptr(out_bf16),
         out_bf16.stride(0) if out_bf16 is not None else 0, ptr(out_f32),
         out_f32.stride(0) if out_f32 is not None else 0, ptr(colsum), int(zero_dy), stream_ptr())


def bn_finalize(stats, n_rows, C_, mean_invstd, running_mean, running_var, nbt, eps=1e-5, momentum=0.1, n_partials=1):
    call("sbr_bn_finalize", ptr(stats), int(n_partials), int(n_rows), int(C_), float(eps), float(momentum),
         ptr(mean_invstd),
         ptr(running_mean), ptr(running_var), ptr(nbt), stream_ptr())


def bn_eval_coeffs(running_mean, running_var, C_, mean_invstd, eps=1e-5):
    call("sbr_bn_eval_coeffs", ptr(running_mean), ptr(running_var), int(C_), float(eps), ptr(mean_invstd),
         stream_ptr())


def bn_apply(z, mean_invstd, gamma, beta, act, rows, C_, out_bf16=None, out_f32=None):
    call("sbr_bn_apply", ptr(z), z.stride(0), ptr(mean_invstd), ptr(gamma), ptr(beta), _act(act), int(rows), int(C_),
         ptr(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0, ptr(out_f32),
         out_f32.stride(0) if out_f32 is not None else 0, stream_ptr())


def bn_bwd_reduce(dy, y, act, z, mean_invstd, rows, C_, sums):
    yf, yb, ldy = _y_args(y)
    call("sbr_bn_bwd_reduce", ptr(dy), dy.stride(0), yf, yb, ldy, _act(act), ptr(z), z.stride(0), ptr(mean_invstd),
         int(rows), int(C_), ptr(sums), stream_ptr())


def bn_bwd_apply(dy, y, act, z, mean_invstd, gamma, sums, rows, C_, dz_bf16=None, dz_f32=None, dgamma=None,
                 dbeta=None, n_replicas=1):
    yf, yb, ldy = _y_args(y)
    call("sbr_bn_bwd_apply", ptr(dy), dy.stride(0), yf, yb, ldy, _act(act), ptr(z), z.stride(0), ptr(mean_invstd),
         ptr(gamma), ptr(sums), int(n_replicas), int(rows), int(C_), ptr(dz_bf16), dz_bf16.stride(0) if dz_bf16 is not None else 0,
         ptr(dz_f32), dz_f32.stride(0) if dz_f32 is not None else 0, ptr(dgamma), ptr(dbeta), stream_ptr())


def score_loss(eu, ei, B, n, ku, ki, D, agg_max_user, agg_max_item, loss_kind, aggregator_sum, ssm_shift, logits,
               loss_acc, deu=None, dei=None, u_agg=None, i_agg=None):
    lk = loss_kind if isinstance(loss_kind, int) else LOSS[loss_kind]
    call("sbr_score_loss", ptr(eu), ptr(ei), int(B), int(n), int(ku), int(ki), int(D), int(agg_max_user),
         int(agg_max_item), lk, int(aggregator_sum), float(ssm_shift), ptr(logits), ptr(loss_acc), ptr(deu), ptr(dei),
         ptr(u_agg), ptr(i_agg), stream_ptr())


def score_bwd(eu, ei, B, n, ku, ki, D, agg_max_user, agg_max_item, dlogits, deu, dei):
    call("sbr_score_bwd", ptr(eu), ptr(ei), int(B), int(n), int(ku), int(ki), int(D), int(agg_max_user),
         int(agg_max_item), ptr(dlogits), ptr(deu), ptr(dei), stream_ptr())


BN_SUM_REPLICAS = 8


def score_loss_bn(eu, bn_u, ei, bn_i, B, n, D, loss_kind, aggregator_sum, ssm_shift, logits, loss_acc, deu, dei):
    """bn_u / bn_i: None or dict(z, mean_invstd, gamma, beta, sums [BN_SUM_REPLICAS, 2 D] zeroed)"""
    lk = loss_kind if isinstance(loss_kind, int) else LOSS[loss_kind]

    def pack(d):
        if d is None:
            return None
        b = BnInline()
        b.z, b.mean_invstd, b.gamma, b.beta, b.sums = (ptr(d[k]) for k in ("z", "mean_invstd", "gamma", "beta", "sums"))
        return C.byref(b)
    call("sbr_score_loss_bn", ptr(eu), pack(bn_u), ptr(ei), pack(bn_i), int(B), int(n), int(D), lk,
         int(aggregator_sum), float(ssm_shift), ptr(logits), ptr(loss_acc), ptr(deu), ptr(dei), BN_SUM_REPLICAS,
         stream_ptr())


INFONCE_GEMM_MIN_N = 1024          # one group of at least this many rows takes the tensor-core route
INFONCE_GEMM_MAX_BYTES = 24 << 30  # n x n logits (fp32) + weights (bf16)


def infonce_gemm(e, n, D, temperature, weight, loss_acc, de, accumulate):
    """symmetric InfoNCE of ONE group of n rows (e fp32 [n, 2, D]) with the n x n logits and both gradient products on
    the tensor cores: split-bf16 logits GEMM (K = 3 D) -> row / column log-sum-exp + loss -> softmax-weight matrix ->
    dE0 = W E1, dE1 = W^T E0 (written / accumulated into ``de`` [n, 2, D])"""
    dev = e.device
    D8 = pad8(D)
    a3 = torch.empty((n, 3 * D8), dtype=BF16, device=dev)
    b3 = torch.empty((n, 3 * D8), dtype=BF16, device=dev)
    call("sbr_infonce_split", ptr(e), int(n), int(D), ptr(a3), ptr(b3), stream_ptr())
    L = torch.empty((n, n), dtype=F32, device=dev)
    gemm(a3, b3, n, n, 3 * D8, out_f32=L, alpha=1.0 / float(temperature))
    lse = torch.empty((2, n), dtype=F32, device=dev)
    n_chunks = max(1, min(64, n // 256))
    part = torch.empty((n_chunks, n, 2), dtype=F32, device=dev)
    call("sbr_infonce_lse", ptr(L), int(n), float(weight) / float(n), ptr(lse[0]), ptr(lse[1]), ptr(part), int(n_chunks),
         ptr(loss_acc), stream_ptr())
    if de is None:
        return
    W = torch.empty((n, n), dtype=BF16, device=dev)
    call("sbr_infonce_weights", ptr(L), int(n), ptr(lse[0]), ptr(lse[1]), ptr(W), stream_ptr())
    alpha = float(weight) / (float(n) * float(temperature))
    d3 = de.view(n, 2, D)
    # dE0 = alpha W E1 (B = the hi part of E1, [K = n, N = D] MN-major);  dE1 = alpha W^T E0 (A = W read MN-major)
    gemm(W, b3, n, D, n, b_mn=True, ldb=3 * D8, out_f32=d3[:, 0, :], alpha=alpha, atomic_out=bool(accumulate))
    gemm(W, a3, n, D, n, a_mn=True, b_mn=True, lda=n, ldb=3 * D8, out_f32=d3[:, 1, :], alpha=alpha,
         atomic_out=bool(accumulate))


def infonce(e, G, n, D, temperature, weight, loss_acc, de, accumulate, lse_ws=None):
    import os
    if (G == 1 and n >= INFONCE_GEMM_MIN_N and n % 8 == 0 and D % 8 == 0 and 6 * n * n <= INFONCE_GEMM_MAX_BYTES
            and os.environ.get("SBR_INFONCE_GEMM", "1") != "0"):
        return infonce_gemm(e, n, D, temperature, weight, loss_acc, de, accumulate)
    if lse_ws is None:
        lse_ws = torch.empty(2 * G * n, dtype=F32, device=e.device)
    call("sbr_infonce", ptr(e), int(G), int(n), int(D), float(temperature), float(weight), ptr(loss_acc), ptr(de),
         int(accumulate), ptr(lse_ws), stream_ptr())


def aggregate(e, rows, k, D, agg_max, out_f32=None, out_bf16=None):
    call("sbr_aggregate", ptr(e), int(rows), int(k), int(D), int(agg_max), ptr(out_f32), ptr(out_bf16),
         out_bf16.stride(0) if out_bf16 is not None else 0, stream_ptr())


ADAM_CHUNK = 1024  # elements per block of the Adam kernel (csrc/util.cu)


class AdamPlan:
    """device-side tables for the one-launch multi-tensor Adam(W)"""

    def __init__(self, entries, device):
        """entries: list of dict(param, grad, exp_avg, exp_avg_sq, shadow(optional bf16 [rows, ld]))"""
        arr = (AdamTensor * len(entries))()
        self.entries = entries  # keeps every tensor whose raw pointer is stored below alive
        c2t, coff = [], []
        for i, e in enumerate(entries):
            p = e["param"]
            arr[i].param, arr[i].grad = ptr(p), ptr(e["grad"])
            arr[i].exp_avg, arr[i].exp_avg_sq = ptr(e["exp_avg"]), ptr(e["exp_avg_sq"])
            sh = e.get("shadow")
            arr[i].shadow_bf16 = ptr(sh)
            arr[i].numel = p.numel()
            arr[i].cols = p.shape[-1] if (sh is not None and p.dim() >= 1) else max(1, p.numel())
            arr[i].shadow_ld = sh.stride(0) if sh is not None else 0
            for off in range(0, p.numel(), ADAM_CHUNK):
                c2t.append(i)
                coff.append(off)
        self.n_tensors = len(entries)
        self.total_chunks = len(c2t)
        self.tensors = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        self.chunk_to_tensor = torch.tensor(c2t, dtype=torch.int32, device=device)
        self.chunk_offset = torch.tensor(coff, dtype=torch.int64, device=device)

    def step(self, lr, beta1, beta2, eps, wd, decoupled, step_dev, grad_scale=1.0):
        call("sbr_adam_step", ptr(self.tensors), self.n_tensors, self.total_chunks, ptr(self.chunk_to_tensor),
             ptr(self.chunk_offset), float(lr), float(beta1), float(beta2), float(eps), float(wd), int(decoupled),
             ptr(step_dev), float(grad_scale), stream_ptr())

    def step_mc(self, comm, grid_blocks, lr, beta1, beta2, eps, wd, decoupled, step_dev, grad_scale=1.0,
                apply_adam=True):
        """data parallel: in-switch all-reduce of the flat gradient buffer + the update, one kernel (sbr_adam_step_mc);
        ``comm`` is a ``_lib.McComm``"""
        call("sbr_adam_step_mc", ptr(self.tensors), self.n_tensors, self.total_chunks, ptr(self.chunk_to_tensor),
             ptr(self.chunk_offset), float(lr), float(beta1), float(beta2), float(eps), float(wd), int(decoupled),
             ptr(step_dev), float(grad_scale), int(bool(apply_adam)), C.byref(comm), int(grid_blocks), stream_ptr())


def topk_n_splits(U: int, I: int, D: int, k: int, n_sms: int = 148) -> int:
    nu = 256 if D <= 256 else 128
    user_tiles = (U + nu - 1) // nu
    item_tiles = (I + 127) // 128
    want = max(1, -(-n_sms // user_tiles))
    splits = min(want, item_tiles, max(1, 1024 // k))
    tps = -(-item_tiles // splits)
    return -(-item_tiles // tps)  # no empty split


def topk_scores_masked(users_bf16, items_bf16, U, I, D, seen_indptr, seen_indices, k, n_splits=None, item_offset=0,
                       return_keys=False):
    """exact masked top-k of users @ items^T.  Returns (vals [U,k] f32, idx [U,k] i32) or packed keys [U,k]."""
    dev = users_bf16.device
    if n_splits is None:
        n_splits = topk_n_splits(U, I, D, k, torch.cuda.get_device_properties(dev).multi_processor_count)
    nbytes = C.c_int64(0)
    call("sbr_topk_workspace_bytes", int(U), int(I), int(D), int(k), int(n_splits), C.byref(nbytes))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    part = torch.empty((n_splits, U, k), dtype=torch.int64, device=dev)
    call("sbr_topk_scores_masked", ptr(users_bf16), users_bf16.stride(0), ptr(items_bf16), items_bf16.stride(0), int(U),
         int(I), int(D), ptr(seen_indptr), ptr(seen_indices), int(k), int(n_splits), int(item_offset), ptr(part),
         ptr(ws), nbytes.value, stream_ptr())
    return topk_merge(part, n_splits, U, k, return_keys=return_keys)


def topk_merge(keys, L, U, k, return_keys=False):
    dev = keys.device
    if return_keys:
        out = torch.empty((U, k), dtype=torch.int64, device=dev)
        call("sbr_topk_merge", ptr(keys), int(L), int(U), int(k), None, None, ptr(out), stream_ptr())
        return out
    vals = torch.empty((U, k), dtype=F32, device=dev)
    idx = torch.empty((U, k), dtype=torch.int32, device=dev)
    call("sbr_topk_merge", ptr(keys), int(L), int(U), int(k), ptr(vals), ptr(idx), None, stream_ptr())
    return vals, idx


_KS_CACHE = {}


def metrics_at_k(topk_idx, tgt_indptr, tgt_indices, ks, n_items, want_item_hits=False):
    U, k = topk_idx.shape
    dev = topk_idx.device
    key = (tuple(sorted(ks)), str(dev))
    ks_dev = _KS_CACHE.get(key)
    if ks_dev is None:  # (cached: a host list -> device copy is a synchronising H2D, not capturable in a CUDA graph)
        ks_dev = _KS_CACHE[key] = torch.tensor(sorted(ks), dtype=torch.int32, device=dev)
    out = torch.zeros((7, len(ks), U), dtype=F32, device=dev)
    hits = torch.zeros((len(ks), n_items), dtype=torch.int32, device=dev) if want_item_hits else None
    call("sbr_metrics_at_k", ptr(topk_idx), int(U), int(k), ptr(tgt_indptr), ptr(tgt_indices), ptr(ks_dev), len(ks),
         ptr(out), ptr(hits), int(n_items), stream_ptr())
    return out, hits


def sample_batch(coo_user, coo_item, train_indptr, train_indices, items_in_split, B, n_neg, seed, step_dev, out_u,
                 out_i):
    call("sbr_sample_batch", ptr(coo_user), ptr(coo_item), coo_user.numel(), ptr(train_indptr), ptr(train_indices),
         ptr(items_in_split), items_in_split.numel(), int(B), int(n_neg), int(seed), ptr(step_dev), ptr(out_u),
         ptr(out_i), stream_ptr())


def sample_epoch_batch(coo_user, coo_item, order, offset, train_indptr, train_indices, items_in_split, B, n_neg, seed,
                       step_dev, out_u, out_i):
    """batch ``order[offset : offset + B]`` of a shuffled epoch + ``n_neg`` 'uniform_recbole' negatives per slot"""
    call("sbr_sample_epoch_batch", ptr(coo_user), ptr(coo_item), coo_user.numel(), ptr(order), int(offset),
         ptr(train_indptr), ptr(train_indices), ptr(items_in_split), items_in_split.numel(), int(B), int(n_neg),
         int(seed), ptr(step_dev), ptr(out_u), ptr(out_i), stream_ptr())


NEG_STRATEGY = {"uniform_recbole": 0, "uniform": 1}


def sample_negatives(coo_user, coo_item, order, offset, train_indptr, train_indices, items_in_split, item_pos, B, n_neg,
                     strategy, seed, step_dev, out_u, out_i):
    """positives (a random train interaction per slot when ``order`` is None, else ``order[offset + b]``) + ``n_neg``
    negatives per slot by ``strategy`` ('uniform_recbole' | 'uniform': data/sampling.py)"""
    if strategy not in NEG_STRATEGY:
        raise ValueError(f'Sampling strategy "{strategy}" not yet supported.')  # data/dataset.py:375
    call("sbr_sample_negatives", ptr(coo_user), ptr(coo_item), coo_user.numel(), ptr(order), int(offset),
         ptr(train_indptr), ptr(train_indices), ptr(items_in_split), items_in_split.numel(), ptr(item_pos), int(B),
         int(n_neg), NEG_STRATEGY[strategy], int(seed), ptr(step_dev), ptr(out_u), ptr(out_i), stream_ptr())


def logit_bias_fwd(logits, u_idx, i_idx, user_bias=None, item_bias=None, global_bias=None):
    B, n = logits.shape
    call("sbr_logit_bias_fwd", ptr(logits), int(B), int(n), ptr(u_idx), ptr(i_idx), ptr(user_bias), ptr(item_bias),
         ptr(global_bias), stream_ptr())


def clamp_min_fwd(x, lo, clamped=None):
    """x = max(x, lo) in place (contiguous fp32); ``clamped`` uint8 like x records where"""
    call("sbr_clamp_min_fwd", ptr(x), x.numel(), float(lo), ptr(clamped), stream_ptr())


def clamp_min_bwd(dx, clamped):
    call("sbr_clamp_min_bwd", ptr(dx), dx.numel(), ptr(clamped), stream_ptr())


def logit_bias_bwd(dlogits, u_idx, i_idx, d_user_bias=None, d_item_bias=None, d_global_bias=None):
    B, n = dlogits.shape
    call("sbr_logit_bias_bwd", ptr(dlogits), int(B), int(n), ptr(u_idx), ptr(i_idx), ptr(d_user_bias), ptr(d_item_bias),
         ptr(d_global_bias), stream_ptr())
