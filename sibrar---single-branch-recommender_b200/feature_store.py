"""Device-resident feature store: built ONCE from ``dataset.{user,item}_features`` and then read by the kernels.

Replaces the per-batch host path of the reference (``data/Feature.py:140-162``: ``np.vectorize`` dict lookups,
``csr.toarray()``, ``torch.tensor(values, device=...)`` with a D2H sync + an H2D copy per modality per step).

HBM layout per feature (row = position in the feature's value table, NOT the entity index):
  * ``remap``  int32 [n_entities]  entity index -> row (-1: entity has no row; identity maps are dropped)
  * dense (VECTOR / CONTINUOUS / DISCRETE, and 'interactions' when dense enough):
        ``x16`` bf16 [n_rows, pad8(d)]  row-major -- read K-major by the forward GEMM and MN-major by the wgrad GEMM
  * bits ('interactions', density >= 0.4 %): ``bits`` int32 [n_rows, 4*ceil(d/128)] and ``bits_t`` of the transposed
        matrix -- 1 bit per element; expanded to bf16 in shared memory by the GEMM (sbr_gemm_bits_bf16)
  * csr ('interactions', sparse route): ``indptr`` int64 / ``indices`` int32 of the matrix and of its transpose
  * CATEGORICAL: ``codes`` int32 [n_rows];  TAG: ``codes`` int32 [n_rows, max_tags] padded with ``pad_id``
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp
import torch

from . import ops


def feature_type(feature) -> str:
    t = feature.feature_definition.type
    return str(getattr(t, "value", t)).lower()


class DeviceFeature:
    def __init__(self, name: str, feature, n_entities: int, device, dense_min_density: float = 0.004,
                 dense_max_bytes: int = 48 << 30):
        self.name = name
        self.type = feature_type(feature)
        self.device = device
        values = feature.values
        idx = np.asarray(feature._indices)
        self.n_rows = int(values.shape[0])
        self.remap = None
        if not (len(idx) == n_entities and np.array_equal(idx, np.arange(n_entities))):
            remap = np.full(n_entities, -1, dtype=np.int32)
            remap[idx] = np.arange(len(idx), dtype=np.int32)
            self.remap = torch.from_numpy(remap).to(device)
        self.x16 = self.csr = self.csr_t = self.codes = self.bits = self.bits_t = None
        self.csr_vals = self.csr_t_vals = self.csr_seg = self.csr_t_seg = None
        self.max_tags, self.pad_id, self.n_cat = 0, -1, 0
        if self.type == "categorical":
            self.kind = "categorical"
            self.n_cat = int(feature.n_unique_categories)
            self.codes = torch.from_numpy(np.asarray(values).astype(np.int32)).to(device)
            self.dim = 0
        elif self.type == "tag":
            self.kind = "tag"
            v = np.asarray(values).astype(np.int32)
            self.codes = torch.from_numpy(np.ascontiguousarray(v)).to(device)
            self.max_tags = int(v.shape[1])
            self.pad_id = int(feature.dim)  # padding index == number of tags (data/Feature.py:254-255)
            self.n_cat = int(feature.dim) + 1
            self.dim = int(feature.dim)
            # tag -> feature rows, cut into segments of <= 128 entries, with the EmbeddingBag(mean) weight 1 / #tags of
            # the row: the backward of a LARGE tag vocabulary is a gather-sum per segment (sbr_spmm_csr, segment mode)
            # instead of ~n_rows * tags_per_row contended atomics onto n_tags rows
            rr, cc = np.nonzero(v != self.pad_id)
            tags = v[rr, cc]
            cnt = np.maximum((v != self.pad_id).sum(1), 1).astype(np.float32)
            order = np.argsort(tags, kind="stable")
            tags, rows_sorted = tags[order], rr[order].astype(np.int32)
            starts = np.flatnonzero(np.r_[True, tags[1:] != tags[:-1]]) if tags.size else np.zeros(0, np.int64)
            ends = np.r_[starts[1:], tags.size] if tags.size else np.zeros(0, np.int64)
            seg_ptr, seg_tag = [0], []
            for a, b in zip(starts.tolist(), ends.tolist()):
                for lo in range(a, b, 128):
                    seg_ptr.append(min(b, lo + 128))
                    seg_tag.append(int(tags[a]))
            self.tag_segments = (torch.from_numpy(np.asarray(seg_ptr, dtype=np.int64)).to(device),
                                 torch.from_numpy(rows_sorted).to(device),
                                 torch.from_numpy((1.0 / cnt[rows_sorted]).astype(np.float32)).to(device),
                                 torch.from_numpy(np.asarray(seg_tag, dtype=np.int32)).to(device))
        elif sp.issparse(values):
            m = values.tocsr()
            m.sort_indices()
            self.dim = int(m.shape[1])
            density = m.nnz / max(1, m.shape[0] * m.shape[1])
            dense_bytes = m.shape[0] * ops.pad8(m.shape[1]) * 2
            ip = torch.from_numpy(m.indptr.astype(np.int64)).to(device)
            ix = torch.from_numpy(m.indices.astype(np.int32)).to(device)
            binary = m.nnz == 0 or bool(np.all(m.data == 1))
            # the reference feeds the stored values (``csr[rows].toarray().float()``, data/Feature.py:147-150): duplicate
            # history rows are counts > 1 in its sampling matrices.  None = all ones (the bit-packed route needs that).
            vals = None if binary else torch.from_numpy(m.data.astype(np.float32)).to(device)
            if (density >= dense_min_density and binary and m.shape[0] * m.shape[1] // 4 <= dense_max_bytes
                    and os.environ.get("SBR_BITS", "1") != "0"):
                # bit-packed multi-hot rows (and the transposed matrix for the wgrad), 1 bit per element in HBM:
                # the projection is a tcgen05 GEMM whose A operand is expanded to bf16 in shared memory
                self.kind = "bits"
                self.bits = ops.pack_bits(m, device)
                self.bits_t = ops.pack_bits(m.T.tocsr(), device)
            elif density >= dense_min_density and dense_bytes <= dense_max_bytes:
                # dense bf16 multi-hot, resident in HBM: the projection becomes a tcgen05 GEMM
                self.kind = "dense"
                self.x16 = ops.csr_to_dense_bf16(ip, ix, m.shape[0], m.shape[1], vals)
            else:
                self.kind = "csr"
                mt = m.T.tocsr()
                mt.sort_indices()
                self.csr = (ip, ix)
                self.csr_t = (torch.from_numpy(mt.indptr.astype(np.int64)).to(device),
                              torch.from_numpy(mt.indices.astype(np.int32)).to(device))
                self.csr_vals = vals
                self.csr_t_vals = None if binary else torch.from_numpy(mt.data.astype(np.float32)).to(device)
                self.csr_seg = csr_segments(m.indptr, device)
                self.csr_t_seg = csr_segments(mt.indptr, device)
        else:
            self.kind = "dense"
            v = np.asarray(values, dtype=np.float32)
            if v.ndim == 1:
                v = v[:, None]
            v = v.reshape(v.shape[0], -1)
            self.dim = int(v.shape[1])
            x = torch.zeros((v.shape[0], ops.pad8(self.dim)), dtype=torch.bfloat16, device=device)
            x[:, :self.dim] = torch.from_numpy(v).to(device)  # one-off H2D + cast at build time
            self.x16 = x

    def nbytes(self) -> int:
        n = 0
        for t in (self.remap, self.x16, self.codes, self.bits, self.bits_t):
            if t is not None:
                n += t.numel() * t.element_size()
        for pair in (self.csr, self.csr_t):
            if pair is not None:
                n += sum(t.numel() * t.element_size() for t in pair)
        return n


CSR_SEGMENT = 256  # entries per warp-sized unit of work of the sparse kernels


def csr_segments(indptr: np.ndarray, device, seg: int = CSR_SEGMENT):
    """rows of a CSR matrix cut into segments of at most ``seg`` entries (None when no row is longer): ``seg_ptr`` int64
    [n_seg + 1] (a refinement of indptr), ``seg_row`` int32 [n_seg] = row | (row has several segments) << 31,
    ``long_rows`` int32 -- see sbr_spmm_csr_bf16; ``seg_first`` int64 [n_rows + 1]: first segment of every row"""
    indptr = np.asarray(indptr, dtype=np.int64)
    lens = np.diff(indptr)
    if lens.size == 0 or int(lens.max()) <= seg:
        return None
    per_row = np.maximum(1, -(-lens // seg))
    seg_row = np.repeat(np.arange(lens.size, dtype=np.int64), per_row)
    first = np.cumsum(per_row) - per_row
    k = np.arange(seg_row.size, dtype=np.int64) - first[seg_row]
    seg_beg = indptr[seg_row] + k * seg
    seg_ptr = np.concatenate([seg_beg, indptr[-1:]])
    multi = per_row[seg_row] > 1
    enc = (seg_row.astype(np.uint32) | (multi.astype(np.uint32) << np.uint32(31))).view(np.int32)
    long_rows = np.flatnonzero(per_row > 1).astype(np.int32)
    seg_first = np.concatenate([[0], np.cumsum(per_row)]).astype(np.int64)  # row -> its first segment
    return (torch.from_numpy(seg_ptr).to(device), torch.from_numpy(enc).to(device),
            torch.from_numpy(long_rows).to(device), torch.from_numpy(seg_first).to(device))


def csr_to_device(m, device):
    m = m.tocsr()
    m.sort_indices()
    return (torch.from_numpy(m.indptr.astype(np.int64)).to(device),
            torch.from_numpy(m.indices.astype(np.int32)).to(device))
