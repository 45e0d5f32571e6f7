"""Kernel-level parity tests (GPU): every C-ABI entry point against a plain fp32/fp64 torch / numpy statement of the
same arithmetic on the same seeded inputs.  Tolerances: bf16-operand GEMMs 2e-2 relative to the output scale (bf16 has
8 mantissa bits; accumulation is fp32), fp32 kernels 1e-5..1e-4, integer / index results bit-exact."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from sibrar_b200 import ops  # noqa: E402

DEV = "cuda"


def _rand_bf16(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(DEV)


def _padded(t, ld):
    out = torch.full((t.shape[0], ld), float("nan"), dtype=t.dtype, device=t.device)
    out[:, :t.shape[1]] = t
    return out[:, :t.shape[1]]  # view with pitch ld; the pad holds NaN on purpose (must never be read)


def _relerr(a, b):
    return (a.double() - b.double()).abs().max().item() / max(1e-12, b.double().abs().max().item())


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 64, 70), (1000, 100, 200), (257, 200, 1000), (4096, 512, 128),
                                   (77, 16, 8), (513, 24, 3706)])
def test_gemm_linear_forward(M, N, K):
    A = _padded(_rand_bf16(M, K, seed=1), ops.pad8(K) + 8)
    W = _padded(_rand_bf16(N, K, seed=2, scale=0.2), ops.pad8(K))
    bias = torch.randn(N, device=DEV)
    o16 = torch.zeros((M, ops.pad8(N)), dtype=torch.bfloat16, device=DEV)
    o32 = torch.zeros((M, N), dtype=torch.float32, device=DEV)
    stats = torch.zeros(2 * N, device=DEV)
    ops.gemm(A, W, M, N, K, bias=bias, act="relu", out_bf16=o16, out_f32=o32, colstats=stats)
    ref = torch.relu(A.float() @ W.float().T + bias)
    torch.cuda.synchronize()
    assert _relerr(o32, ref) < 2e-5 * math.sqrt(K) + 1e-6, _relerr(o32, ref)
    assert _relerr(o16[:, :N].float(), ref) < 1e-2
    assert _relerr(stats[:N], ref.sum(0)) < 1e-4 and _relerr(stats[N:], (ref * ref).sum(0)) < 1e-4


@pytest.mark.parametrize("rows,din,dout,split", [(1000, 64, 64, 1), (5000, 70, 100, 7), (3000, 768, 64, 4),
                                                 (700, 300, 300, 3), (129, 16, 24, 2)])
def test_gemm_wgrad_mn_major(rows, din, dout, split):
    """dW[out, in] += dY^T X with X [rows, in], dY [rows, out] both read MN-major (contraction over rows)."""
    X = _padded(_rand_bf16(rows, din, seed=3), ops.pad8(din))
    dY = _padded(_rand_bf16(rows, dout, seed=4), ops.pad8(dout) + 8)
    dW = torch.ones((dout, din), dtype=torch.float32, device=DEV)
    ops.gemm(X, dY, din, dout, rows, a_mn=True, b_mn=True, out_f32=dW, transpose_out=True, atomic_out=True,
             split_k=split)
    ref = 1.0 + dY.float().T @ X.float()
    torch.cuda.synchronize()
    assert _relerr(dW, ref) < 1e-4, _relerr(dW, ref)


@pytest.mark.parametrize("rows,din,dout", [(1000, 64, 64), (333, 100, 70), (2048, 512, 256), (100, 24, 16)])
def test_gemm_dgrad_b_mn_major_actgrad(rows, din, dout):
    """dX = (dY W) * relu'(Y_prev), column sums of the result = bias gradient of the previous layer."""
    dY = _padded(_rand_bf16(rows, dout, seed=5), ops.pad8(dout))
    W = _padded(_rand_bf16(dout, din, seed=6, scale=0.3), ops.pad8(din))
    Yp = _padded(torch.relu(_rand_bf16(rows, din, seed=7)), ops.pad8(din))
    o16 = torch.zeros((rows, ops.pad8(din)), dtype=torch.bfloat16, device=DEV)
    o32 = torch.zeros((rows, din), dtype=torch.float32, device=DEV)
    stats = torch.zeros(2 * din, device=DEV)
    ops.gemm(dY, W, rows, din, dout, b_mn=True, out_bf16=o16, out_f32=o32, actgrad_y=Yp, actgrad_act="relu",
             colstats=stats)
    ref = (dY.float() @ W.float()) * (Yp.float() > 0)
    torch.cuda.synchronize()
    assert _relerr(o32, ref) < 1e-4, _relerr(o32, ref)
    assert _relerr(o16[:, :din].float(), ref) < 1e-2
    assert _relerr(stats[:din], ref.sum(0)) < 1e-3


@pytest.mark.parametrize("rows,d,out,density", [(300, 200, 64, 0.05), (1000, 6040, 64, 0.04), (257, 130, 24, 0.3),
                                                  (3706, 777, 128, 0.02), (40000, 300, 64, 0.03), (500, 300, 192, 0.05)])
def test_gemm_bits_matches_dense(rows, d, out, density):
    """bit-packed multi-hot A operand == the same GEMM on the dense bf16 matrix (forward and wgrad forms)"""
    import scipy.sparse as sp
    m = sp.random(rows, d, density=density, format="csr", random_state=rows + d)
    m.data[:] = 1
    dense = torch.from_numpy(m.toarray().astype(np.float32)).to(DEV)
    x16 = ops.cast_bf16(dense)
    bits, bits_t = ops.pack_bits(m, DEV), ops.pack_bits(m.T.tocsr(), DEV)
    w = _rand_bf16(out, ops.pad8(d), seed=3)[:, :d]
    bias = torch.randn(out, device=DEV)
    y_ref, y = torch.empty(rows, out, device=DEV), torch.empty(rows, out, device=DEV)
    ops.gemm(x16, w, rows, out, d, bias=bias, act="relu", out_f32=y_ref)
    ops.gemm_bits(bits, w, rows, out, d, bias=bias, act="relu", out_f32=y)
    assert torch.equal(y, y_ref)
    # split-K forward (atomic fp32): same sum, different order
    z = torch.zeros(rows, out, device=DEV)
    ops.gemm_bits(bits, w, rows, out, d, out_f32=z, atomic_out=True, split_k=3)
    assert _relerr(torch.relu(z + bias), y_ref) < 1e-5
    # sliced split-K (one fp32 slice per partition, no atomics) + the reduce pass with bias / activation
    ns = ops.effective_splits(d, 3)
    part = torch.full((ns, rows, out), float("nan"), device=DEV)
    ops.gemm_bits(bits, w, rows, out, d, out_f32=part.view(ns * rows, out), split_k=ns, split_stride=rows * out)
    y2, y2h = torch.empty(rows, out, device=DEV), torch.empty(rows, ops.pad8(out), dtype=torch.bfloat16, device=DEV)
    ops.splitk_reduce(part, ns, rows, out, bias=bias, act="relu", out_f32=y2, out_bf16=y2h)
    assert _relerr(y2, y_ref) < 1e-5 and _relerr(y2h[:, :out].float(), y_ref) < 1e-2
    # wgrad form: dW [out, d] = dZ^T X  computed as (X^T dZ)^T with the transposed bit matrix as A
    dz = _rand_bf16(rows, ops.pad8(out), seed=4)[:, :out]
    g_ref, g = torch.zeros(out, d, device=DEV), torch.zeros(out, d, device=DEV)
    ops.gemm(x16, dz, d, out, rows, a_mn=True, b_mn=True, out_f32=g_ref, transpose_out=True, atomic_out=True, split_k=1)
    ops.gemm_bits(bits_t, dz, d, out, rows, b_mn=True, out_f32=g, transpose_out=True, atomic_out=True, split_k=2)
    assert _relerr(g, g_ref) < 1e-5
    ns = ops.effective_splits(rows, 2)
    part = torch.full((ns, out, d), float("nan"), device=DEV)
    ops.gemm(x16, dz, d, out, rows, a_mn=True, b_mn=True, out_f32=part.view(ns * out, d), transpose_out=True,
             split_k=ns, split_stride=out * d)
    g2 = g_ref.clone()
    ops.splitk_reduce(part, ns, out, d, out_f32=g2, accumulate=True)
    assert _relerr(g2, 2 * g_ref) < 1e-5


def _ref_topk(u, it, seen, k):
    s = u.double() @ it.double().T
    s = s.cpu().numpy()
    if seen is not None:
        ip, ix = seen
        for r in range(s.shape[0]):
            s[r, ix[ip[r]:ip[r + 1]]] = -np.inf
    order = np.lexsort((np.broadcast_to(np.arange(s.shape[1]), s.shape), -s), axis=-1)[:, :k]
    return np.take_along_axis(s, order, 1), order


def _seen_csr(U, I, per_user, seed):
    rng = np.random.default_rng(seed)
    rows = [np.sort(rng.choice(I, size=min(I, int(rng.integers(0, per_user + 1))), replace=False)) for _ in range(U)]
    indptr = np.zeros(U + 1, dtype=np.int64)
    indptr[1:] = np.cumsum([len(r) for r in rows])
    return indptr, (np.concatenate(rows) if indptr[-1] else np.zeros(0)).astype(np.int32)


@pytest.mark.parametrize("U,I,D,k,n_splits,per_user", [
    (300, 1000, 64, 10, 1, 20), (300, 1000, 64, 10, 3, 20), (1000, 5000, 128, 50, None, 100),
    (257, 3706, 64, 100, None, 300), (130, 700, 512, 20, 2, 50), (64, 129, 72, 5, 2, 129), (5, 40, 8, 1, 1, 3),
    (200, 20000, 256, 10, None, 0)])
def test_topk_scores_masked_exact(U, I, D, k, n_splits, per_user):
    """small-integer embeddings -> every score is exact in fp32 whatever the accumulation order -> the top-k
    (values AND positions, ties -> lowest position) must be bit-exact."""
    g = torch.Generator().manual_seed(U + I)
    u = torch.randint(-3, 4, (U, D), generator=g).to(torch.bfloat16).to(DEV)
    it = torch.randint(-3, 4, (I, D), generator=g).to(torch.bfloat16).to(DEV)
    seen = _seen_csr(U, I, per_user, seed=I) if per_user else None
    ip = torch.from_numpy(seen[0]).to(DEV) if seen else None
    ix = torch.from_numpy(seen[1]).to(DEV) if seen else None
    vals, idx = ops.topk_scores_masked(u, it, U, I, D, ip, ix, k, n_splits=n_splits)
    rv, ri = _ref_topk(u.float(), it.float(), seen, k)
    vals, idx = vals.cpu().numpy(), idx.cpu().numpy()
    finite = np.isfinite(rv)
    assert (idx[finite] == ri[finite]).all(), f"{(idx[finite] != ri[finite]).sum()} of {finite.sum()} positions differ"
    assert (vals[finite] == rv[finite]).all()
    assert (idx[~finite] == -1).all() and np.isneginf(vals[~finite]).all()


def test_topk_full_size_properties():
    """BASELINE configs[4] scale (1e6 items): size-independent properties of the fused kernel -- descending scores,
    unique unseen positions, invariance to the item split, and agreement with torch on a sample of users"""
    U, I, D, k = 20000, 1_000_000, 64, 50
    g = torch.Generator().manual_seed(3)
    u = (torch.randn(U, D, generator=g) / math.sqrt(D)).to(torch.bfloat16).to(DEV)
    it = torch.randn(I, D, generator=g).to(torch.bfloat16).to(DEV)
    per = 64
    seen_ix = torch.sort(torch.randint(0, I, (U, per), device=DEV, dtype=torch.int32), dim=1).values
    ip = torch.arange(0, (U + 1) * per, per, dtype=torch.int64, device=DEV)
    v1, i1 = ops.topk_scores_masked(u, it, U, I, D, ip, seen_ix.reshape(-1).contiguous(), k, n_splits=1)
    v4, i4 = ops.topk_scores_masked(u, it, U, I, D, ip, seen_ix.reshape(-1).contiguous(), k, n_splits=4)
    assert torch.equal(i1, i4) and torch.equal(v1, v4)                       # split invariance (exact merge)
    assert bool((v1[:, :-1] >= v1[:, 1:]).all())                             # sorted
    assert bool((i1 >= 0).all()) and bool((i1 < I).all())
    srt = torch.sort(i1, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())                           # unique positions
    hit = (i1.unsqueeze(2) == seen_ix.unsqueeze(1)).any(dim=2)
    assert not bool(hit.any())                                               # no seen item is recommended
    sample = torch.arange(0, U, U // 32, device=DEV)[:32]
    sc = u[sample].float() @ it.float().T
    sc.scatter_(1, seen_ix[sample].long(), float("-inf"))
    rv, ri = torch.topk(sc, k, dim=1)
    assert float((v1[sample] - rv).abs().max()) < 1e-4
    same = (i1[sample] == ri)
    gap_ok = (rv[:, :-1] - rv[:, 1:]) > 1e-5                                 # positions must agree where scores are apart
    safe = torch.ones_like(same)
    safe[:, 1:] &= gap_ok
    safe[:, :-1] &= gap_ok
    assert bool(same[safe].all())


def test_topk_float_scores_close():
    U, I, D, k = 500, 8000, 64, 20
    u = _rand_bf16(U, D, seed=11, scale=D ** -0.5)
    it = _rand_bf16(I, D, seed=12)
    vals, idx = ops.topk_scores_masked(u, it, U, I, D, None, None, k)
    rv, ri = _ref_topk(u.float(), it.float(), None, k)
    vals, idx = vals.cpu().numpy(), idx.cpu().numpy()
    assert np.abs(vals - rv).max() < 1e-4
    gap = np.abs(np.diff(np.concatenate([rv, rv[:, -1:] - 1], 1), axis=1))
    safe = (gap > 1e-4) & (np.abs(np.diff(np.concatenate([rv[:, :1] + 1, rv], 1), axis=1)) > 1e-4)
    assert (idx[safe] == ri[safe]).all()


def test_topk_merge_and_metrics():
    from oracle import sbnet_oracle as O
    import scipy.sparse as sp
    U, I, k = 400, 3000, 20
    rng = np.random.default_rng(0)
    top = np.stack([rng.choice(I, size=k, replace=False) for _ in range(U)]).astype(np.int32)
    tgt = sp.random(U, I, density=0.01, format="csr", random_state=1)
    tgt.sort_indices()
    ks = [1, 5, 10, 20]
    out, hits = ops.metrics_at_k(torch.from_numpy(top).to(DEV), torch.from_numpy(tgt.indptr.astype(np.int64)).to(DEV),
                                 torch.from_numpy(tgt.indices.astype(np.int32)).to(DEV), ks, I, want_item_hits=True)
    ref = O.metrics_at_k(top, tgt, ks, n_items=I)
    out = out.cpu().numpy()
    for mi, m in enumerate(["ndcg", "precision", "recall", "f_score", "hitrate", "ap", "rr"]):
        for ki, kk in enumerate(ks):
            assert np.abs(out[mi, ki] - ref[f"{m}@{kk}"]).max() < 2e-6, (m, kk)
    for ki, kk in enumerate(ks):
        assert hits[ki].sum().item() / I == pytest.approx(ref[f"coverage@{kk}"])


def test_spmm_matches_dense():
    import scipy.sparse as sp
    rows, d, C_ = 700, 900, 64
    m = sp.random(rows, d, density=0.03, format="csr", random_state=3)
    m.data[:] = 1
    m.sort_indices()
    Wt = torch.randn(d, C_, device=DEV)
    b = torch.randn(C_, device=DEV)
    out = torch.empty(rows, C_, device=DEV)
    ip = torch.from_numpy(m.indptr.astype(np.int64)).to(DEV)
    ix = torch.from_numpy(m.indices.astype(np.int32)).to(DEV)
    ops.spmm_csr(ip, ix, rows, Wt, C_, b, "relu", out)
    ref = torch.relu(torch.from_numpy(m.toarray()).float().to(DEV) @ Wt + b)
    assert _relerr(out, ref) < 1e-5
    dense = ops.csr_to_dense_bf16(ip, ix, rows, d)
    assert (dense[:, :d].float().cpu().numpy() == m.toarray()).all()
    outT = torch.zeros(C_, rows, device=DEV)
    ops.spmm_csr(ip, ix, rows, Wt, C_, None, None, outT, transpose_out=True)
    assert _relerr(outT.T, torch.from_numpy(m.toarray()).float().to(DEV) @ Wt) < 1e-5


def test_row_gather_fwd_bwd():
    n_ent, C_, k = 50, 24, 2
    g = torch.Generator().manual_seed(5)
    table = torch.randn(n_ent, C_, generator=g).to(DEV)
    emb = torch.randn(7, C_, generator=g).to(DEV)
    bag = torch.randn(6, C_, generator=g).to(DEV)
    cat = torch.randint(0, 7, (n_ent,), generator=g).int().to(DEV)
    tags = torch.randint(0, 6, (n_ent, 4), generator=g).int().to(DEV)  # pad id = 5
    remap = torch.randperm(n_ent, generator=g).int().to(DEV)
    grads = [torch.zeros_like(table), torch.zeros_like(emb), torch.zeros_like(bag)]
    srcs = ops.make_modality_srcs([
        dict(kind=0, remap=remap, table=table, grad=grads[0]),
        dict(kind=1, remap=None, table=emb, grad=grads[1], codes=cat),
        dict(kind=2, remap=None, table=bag, grad=grads[2], codes=tags, max_tags=4, pad_id=5)], DEV)
    idx = torch.randint(0, n_ent, (33,), generator=g).to(DEV)
    mods = torch.randint(0, 3, (33 * k,), generator=g).to(torch.uint8).to(DEV)
    keep = (torch.rand(33 * k, C_, generator=g) > 0.3).to(torch.uint8).to(DEV)
    step = torch.zeros(1, dtype=torch.int64, device=DEV)
    out = torch.zeros(33 * k, C_, device=DEV)
    ops.row_gather_fwd(srcs, 3, idx, mods, k, C_, True, 0.3, 1, step, keep, out_f32=out)

    t_table, t_emb, t_bag = (x.clone().requires_grad_() for x in (table, emb, bag))
    fi = idx.repeat_interleave(k)
    x0 = t_table[remap[fi].long()]
    x1 = t_emb[cat[fi].long()]
    valid = (tags[fi] != 5).float()
    x2 = (t_bag[tags[fi].long()] * valid[..., None]).sum(1) / valid.sum(1).clamp(min=1)[:, None]
    x = torch.where((mods == 0)[:, None], x0, torch.where((mods == 1)[:, None], x1, x2))
    ref = torch.nn.functional.normalize(x, dim=-1) * keep.float() / 0.7
    assert _relerr(out, ref) < 1e-5
    dx = torch.randn(33 * k, C_, device=DEV)
    ops.row_gather_bwd(srcs, 3, idx, mods, k, C_, True, 0.3, 1, step, keep, dx)
    ref.backward(dx)
    gb = t_bag.grad.clone()
    gb[5] = 0
    grads[2][5] = 0
    for got, want in zip(grads, (t_table.grad, t_emb.grad, gb)):
        assert _relerr(got, want) < 1e-4

    # sorted-run backward (counting sort by (modality, source row) + run reduce): same gradients
    grads2 = [torch.zeros_like(table), torch.zeros_like(emb), torch.zeros_like(bag)]
    srcs2 = ops.make_modality_srcs([
        dict(kind=0, remap=remap, table=table, grad=grads2[0], key_base=0),
        dict(kind=1, remap=None, table=emb, grad=grads2[1], codes=cat, key_base=n_ent),
        dict(kind=2, remap=None, table=bag, grad=grads2[2], codes=tags, max_tags=4, pad_id=5, key_base=n_ent + 7)], DEV)
    for rpw in (1, 4, 32):
        for gr in grads2:
            gr.zero_()
        plan = ops.GatherPlan(n_ent + 7 + n_ent, 33 * k, DEV, rows_per_warp=rpw)
        plan.build(srcs2, 3, idx, mods, k)
        plan.backward(srcs2, 3, C_, True, 0.3, 1, step, keep, dx)
        sk = plan.sorted_keys.cpu().numpy()
        assert (np.diff(sk) >= 0).all() and sorted(plan.perm.cpu().tolist()) == list(range(33 * k))
        grads2[2][5] = 0
        for got, want in zip(grads2, (t_table.grad, t_emb.grad, gb)):
            assert _relerr(got, want) < 1e-4, rpw
    # keep bits stored by the forward (Philox-free backward): same gradients with a hash-generated mask
    out_h = torch.zeros(33 * k, C_, device=DEV)
    bits = torch.zeros(33 * k, (C_ + 7) // 8, dtype=torch.uint8, device=DEV)
    ops.row_gather_fwd(srcs2, 3, idx, mods, k, C_, True, 0.3, 9, step, None, out_f32=out_h, keep_bits_out=bits)
    res = []
    for kb in (None, bits):
        for gr in grads2:
            gr.zero_()
        plan = ops.GatherPlan(n_ent + 7 + n_ent, 33 * k, DEV)
        plan.build(srcs2, 3, idx, mods, k)
        plan.backward(srcs2, 3, C_, True, 0.3, 9, step, None, dx, keep_bits=kb)
        res.append([gr.clone() for gr in grads2])
    for a_, b_ in zip(*res):
        assert _relerr(a_, b_) < 1e-5
    kept = torch.stack([(bits[:, c // 8] >> (c % 8)) & 1 for c in range(C_)], dim=1).bool()
    assert torch.equal(kept, out_h != 0) or float((kept != (out_h != 0)).float().mean()) < 0.01  # (exact zeros in x)

    # Philox dropout: deterministic for (seed, step), keeps ~ (1 - p)
    big = torch.zeros(4000, C_, device=DEV)
    idx2 = torch.randint(0, n_ent, (4000,), generator=g).to(DEV)
    ops.row_gather_fwd(srcs, 1, idx2, None, 1, C_, False, 0.25, 7, step, None, out_f32=big)
    big2 = torch.zeros_like(big)
    ops.row_gather_fwd(srcs, 1, idx2, None, 1, C_, False, 0.25, 7, step, None, out_f32=big2)
    assert torch.equal(big, big2)
    assert abs((big != 0).float().mean().item() - 0.75) < 0.01


@pytest.mark.parametrize("n_rows,n_tags,max_tags,C", [(3706, 18, 6, 64), (50, 7, 3, 24), (1000, 300, 9, 128),
                                                      (200, 40, 5, 600)])
def test_tag_bag_table_matches_embedding_bag(n_rows, n_tags, max_tags, C):
    """sbr_tag_bag_fwd/bwd == nn.EmbeddingBag(mode='mean', padding_idx=-1) forward / weight gradient; rows without
    tags give zeros; the gradient table is cleared by the backward"""
    g = torch.Generator(device="cpu").manual_seed(n_rows + C)
    codes = torch.full((n_rows, max_tags), n_tags, dtype=torch.int32)
    for r in range(n_rows):
        n = int(torch.randint(0, max_tags + 1, (1,), generator=g))
        codes[r, :n] = torch.randperm(n_tags, generator=g)[:n].to(torch.int32)
    bag = torch.nn.EmbeddingBag(n_tags + 1, C, mode="mean", padding_idx=n_tags).double()
    with torch.no_grad():
        bag.weight.copy_(torch.randn(n_tags + 1, C, generator=g))
    want = bag(codes.long())
    dy = torch.randn(n_rows, C, generator=g)
    want.backward(dy.double())
    w = bag.weight.detach().float().to(DEV)
    out = torch.full((n_rows, C), float("nan"), device=DEV)
    ops.tag_bag_fwd(codes.to(DEV), max_tags, n_tags, w, out)
    assert _relerr(out.cpu(), want.detach()) < 1e-6
    empty = (codes == n_tags).all(dim=1)
    assert empty.any() and (out.cpu()[empty] == 0).all()
    gw = torch.zeros(n_tags + 1, C, device=DEV)
    dyd = dy.to(DEV).clone()
    ops.tag_bag_bwd(codes.to(DEV), max_tags, n_tags, dyd, gw)
    assert _relerr(gw.cpu(), bag.weight.grad) < 1e-5
    assert (gw[n_tags] == 0).all() and (dyd == 0).all()


def test_gemm_partial_colstats_are_deterministic():
    """BatchNorm statistics from the GEMM epilogue: per-warp partial rows + ordered finalize give bit-identical
    mean / invstd on every run (the atomic variant depends on the arrival order) and agree with torch"""
    M, N, K = 50000, 64, 64
    x, w = _rand_bf16(M, K, seed=1), _rand_bf16(N, K, seed=2, scale=0.2)
    rows = ops.gemm_colstats_rows(M, N)
    outs = []
    for _ in range(3):
        a32 = torch.empty(M, N, device=DEV)
        part = torch.full((rows, 2 * N), float("nan"), device=DEV)
        ops.gemm(x, w, M, N, K, out_f32=a32, colstats=part, colstats_rows=rows)
        mi = torch.empty(2 * N, device=DEV)
        ops.bn_finalize(part, M, N, mi, None, None, None, n_partials=rows)
        outs.append(mi.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = x.float() @ w.float().T
    assert _relerr(outs[0][:N], ref.mean(0)) < 1e-4
    assert _relerr(outs[0][N:], 1.0 / torch.sqrt(ref.var(0, unbiased=False) + 1e-5)) < 1e-4


@pytest.mark.parametrize("act", [None, "relu", "tanh"])
@pytest.mark.parametrize("rows,C_,y16", [(1000, 40, False), (4100, 512, True), (3000, 128, False), (2049, 256, True)])
def test_batchnorm_fwd_bwd(act, rows, C_, y16):
    """scalar kernels (C = 40) and the vector kernels (C = 128 / 256 / 512, y saved as bf16 or fp32, with and without
    an activation behind the BatchNorm) against torch.nn.functional.batch_norm + autograd"""
    z = (torch.randn(rows, C_, device=DEV) * 2 + 1).requires_grad_()
    gamma = torch.randn(C_, device=DEV).requires_grad_()
    beta = torch.randn(C_, device=DEV).requires_grad_()
    rm, rv = torch.zeros(C_, device=DEV), torch.ones(C_, device=DEV)
    ref = torch.nn.functional.batch_norm(z, rm.clone(), rv.clone(), gamma, beta, True, 0.1, 1e-5)
    ref = getattr(torch, act)(ref) if act else ref
    stats = torch.stack([z.detach().sum(0), (z.detach() ** 2).sum(0)]).reshape(-1).contiguous()
    mi = torch.empty(2 * C_, device=DEV)
    nbt = torch.zeros(1, dtype=torch.int64, device=DEV)
    ops.bn_finalize(stats, rows, C_, mi, rm, rv, nbt)
    y = torch.empty(rows, C_, device=DEV)
    y_b = torch.empty(rows, C_, device=DEV, dtype=torch.bfloat16)
    ops.bn_apply(z.detach(), mi, gamma.detach(), beta.detach(), act, rows, C_, out_f32=y, out_bf16=y_b)
    assert _relerr(y, ref) < 1e-4
    assert torch.equal(y_b, y.to(torch.bfloat16))
    rm2, rv2 = torch.zeros(C_, device=DEV), torch.ones(C_, device=DEV)
    torch.nn.functional.batch_norm(z.detach(), rm2, rv2, None, None, True, 0.1, 1e-5)
    assert _relerr(rm, rm2) < 1e-4 and _relerr(rv, rv2) < 1e-4 and nbt.item() == 1
    dy = torch.randn(rows, C_, device=DEV)
    ref.backward(dy)
    ysaved = y_b if y16 else y
    # (tanh' from a bf16 y is only as exact as the rounding of y: the fp32 copy is the tight check)
    tol = 2e-4 if (not y16 or act != "tanh") else 2e-2
    sums = torch.zeros(2 * C_, device=DEV)
    ops.bn_bwd_reduce(dy, ysaved, act, z.detach(), mi, rows, C_, sums)
    dz = torch.empty(rows, C_, device=DEV)
    dz_b = torch.empty(rows, C_, device=DEV, dtype=torch.bfloat16)
    dg, db = torch.zeros(C_, device=DEV), torch.zeros(C_, device=DEV)
    ops.bn_bwd_apply(dy, ysaved, act, z.detach(), mi, gamma.detach(), sums, rows, C_, dz_f32=dz, dz_bf16=dz_b, dgamma=dg,
                     dbeta=db)
    assert _relerr(dz, z.grad) < tol and _relerr(dg, gamma.grad) < tol and _relerr(db, beta.grad) < tol
    assert torch.equal(dz_b, dz.to(torch.bfloat16))


@pytest.mark.parametrize("D", [24, 64, 128])  # 24: scalar kernel; 64 / 128: the float4 fast path (any slot count)
@pytest.mark.parametrize("loss,agg_u,agg_i,ku,ki,sum_", [("bpr", 0, 0, 1, 1, 0), ("bpr", 0, 1, 2, 2, 0),
                                                         ("bce", 1, 0, 2, 1, 1), ("sampled_softmax", 0, 0, 1, 2, 0),
                                                         ("bpr", 1, 1, 3, 3, 0)])
def test_score_loss(loss, agg_u, agg_i, ku, ki, sum_, D):
    from oracle import sbnet_oracle as O
    B, n = 37, 6
    eu = torch.randn(B, ku, D, device=DEV).requires_grad_()
    ei = torch.randn(B, n, ki, D, device=DEV).requires_grad_()
    logits = torch.empty(B, n, device=DEV)
    acc = torch.zeros(1, dtype=torch.float64, device=DEV)
    deu, dei = torch.empty_like(eu), torch.empty_like(ei)
    shift = math.log(100 / (n - 1)) if loss == "sampled_softmax" else 0.0
    ops.score_loss(eu.detach(), ei.detach(), B, n, ku, ki, D, agg_u, agg_i, loss, sum_, shift, logits, acc, deu, dei)
    u = eu.max(1).values if agg_u else eu.mean(1)
    it = ei.max(2).values if agg_i else ei.mean(2)
    ref_logits = torch.einsum("be,bce->bc", u, it)
    rl, dlog = O.rec_loss(loss, ref_logits.detach().cpu().numpy(), "sum" if sum_ else "mean", 100, n - 1,
                          "uniform" if loss == "sampled_softmax" else "uniform_recbole")
    ref_logits.backward(torch.from_numpy(dlog).float().to(DEV))
    assert _relerr(logits, ref_logits) < 1e-5
    assert abs(acc.item() - rl) < 1e-5 * max(1, abs(rl))
    assert _relerr(deu, eu.grad) < 1e-4 and _relerr(dei, ei.grad) < 1e-4
    # the gradient-only entry (autograd glue): same gradients from the externally computed d loss / d logits
    deu2, dei2 = torch.empty_like(eu), torch.empty_like(ei)
    ops.score_bwd(eu.detach(), ei.detach(), B, n, ku, ki, D, agg_u, agg_i, torch.from_numpy(dlog).float().to(DEV), deu2,
                  dei2)
    assert _relerr(deu2, eu.grad) < 1e-4 and _relerr(dei2, ei.grad) < 1e-4


@pytest.mark.parametrize("loss,D,n,bn_user,bn_item", [("bpr", 64, 11, True, True), ("bce", 32, 4, False, True),
                                                     ("sampled_softmax", 128, 7, True, False), ("bpr", 16, 3, True, True)])
def test_score_loss_bn_matches_unfused(loss, D, n, bn_user, bn_item):
    """fused (inline BatchNorm + BatchNorm-backward sums) == bn_apply -> score_loss -> bn_bwd_reduce"""
    B = 1000
    g = torch.Generator().manual_seed(D + n)
    zu, zi = torch.randn(B, D, generator=g).to(DEV), (torch.randn(B * n, D, generator=g) * 1.5 + 0.3).to(DEV)

    def side(z, use_bn, seed):
        gg = torch.Generator().manual_seed(seed)
        if not use_bn:
            return z, None
        mi = torch.cat([z.mean(0), 1.0 / torch.sqrt(z.var(0, unbiased=False) + 1e-5)]).contiguous()
        gamma, beta = (torch.rand(D, generator=gg) + 0.5).to(DEV), torch.randn(D, generator=gg).to(DEV)
        e = torch.empty_like(z)
        ops.bn_apply(z, mi, gamma, beta, None, z.shape[0], D, out_f32=e)
        return e, dict(z=z, mean_invstd=mi, gamma=gamma, beta=beta,
                       sums=torch.zeros(ops.BN_SUM_REPLICAS * 2 * D, device=DEV))
    eu, bu = side(zu, bn_user, 1)
    ei, bi = side(zi, bn_item, 2)
    lg0, lg1 = torch.empty(B, n, device=DEV), torch.empty(B, n, device=DEV)
    acc0, acc1 = torch.zeros(1, dtype=torch.float64, device=DEV), torch.zeros(1, dtype=torch.float64, device=DEV)
    du0, di0, du1, di1 = (torch.empty_like(t) for t in (eu, ei, eu, ei))
    ops.score_loss(eu, ei, B, n, 1, 1, D, 0, 0, loss, 0, 0.3, lg0, acc0, du0, di0)
    ops.score_loss_bn(None if bu else eu, bu, None if bi else ei, bi, B, n, D, loss, 0, 0.3, lg1, acc1, du1, di1)
    assert _relerr(lg1, lg0) < 1e-5 and abs(acc1.item() - acc0.item()) < 1e-6 * max(1.0, abs(acc0.item()))
    assert _relerr(du1, du0) < 1e-5 and _relerr(di1, di0) < 1e-5
    for z, b, d in ((zu, bu, du0), (zi, bi, di0)):
        if b is None:
            continue
        ref = torch.zeros(2 * D, device=DEV)
        ops.bn_bwd_reduce(d, None, None, z, b["mean_invstd"], z.shape[0], D, ref)
        got = b["sums"].view(ops.BN_SUM_REPLICAS, 2 * D).sum(0)
        assert _relerr(got, ref) < 1e-4


@pytest.mark.parametrize("G,n,D,accumulate", [(9, 5, 24, 1), (1, 70, 16, 1), (300, 11, 64, 1), (2500, 11, 128, 0),
                                              (700, 11, 512, 1), (333, 32, 256, 0), (40, 3, 8, 1), (1, 300, 128, 1)])
def test_infonce(G, n, D, accumulate):
    """one-warp-per-group kernel (n <= 32: the item side at its real widths) and the two-pass kernels (n > 32) against
    the fp64 oracle (train/regularization_losses.py:28-43)"""
    from oracle import sbnet_oracle as O
    e = torch.randn(G, n, 2, D, device=DEV) * (2.0 / D ** 0.5)
    acc = torch.zeros(1, dtype=torch.float64, device=DEV)
    de = torch.ones_like(e)
    ops.infonce(e, G, n, D, 0.7, 0.3, acc, de, accumulate=accumulate)
    en = e.double().cpu().numpy()
    loss, d0, d1 = O.info_nce(en[:, :, 0], en[:, :, 1], 0.7)
    assert abs(acc.item() - 0.3 * loss) < 1e-5 * max(1, abs(loss))
    ref = (1.0 if accumulate else 0.0) + 0.3 * np.stack([d0, d1], axis=2)
    assert np.abs(de.cpu().numpy() - ref).max() < 1e-5


@pytest.mark.parametrize("n,D,scale,temp,accumulate", [(1024, 64, 1.0, 0.5, False), (2048, 128, 1.0, 0.1, True),
                                                       (1536, 40, 0.3, 1.0, True)])
def test_infonce_tensor_core_route(n, D, scale, temp, accumulate, monkeypatch):
    """one large group (the user side: in-batch negatives): logits / gradient products as tcgen05 GEMMs with split-bf16
    operands (ops.infonce_gemm) against the fp64 oracle (train/regularization_losses.py:28-43) at post-BatchNorm
    magnitudes (unit-variance elements: |logit| up to ~40 / T), and against the CUDA-core kernels on the same input"""
    from oracle import sbnet_oracle as O
    torch.manual_seed(5)
    e = torch.randn(1, n, 2, D, device=DEV) * scale
    acc = torch.zeros(1, dtype=torch.float64, device=DEV)
    de = torch.ones_like(e)
    assert n >= ops.INFONCE_GEMM_MIN_N
    ops.infonce(e, 1, n, D, temp, 0.3, acc, de, accumulate=accumulate)
    en = e.double().cpu().numpy()
    loss, d0, d1 = O.info_nce(en[:, :, 0], en[:, :, 1], temp)
    assert abs(acc.item() - 0.3 * loss) < 2e-4 * max(1, abs(loss))
    ref = (1.0 if accumulate else 0.0) + 0.3 * np.stack([d0, d1], axis=2)
    got = de.cpu().numpy()
    gmax = np.abs(0.3 * np.stack([d0, d1], axis=2)).max()
    assert np.abs(got - ref).max() < 1e-2 * gmax  # (bf16 softmax weights and gradient operands; fp32-grade logits)
    monkeypatch.setenv("SBR_INFONCE_GEMM", "0")
    acc2 = torch.zeros(1, dtype=torch.float64, device=DEV)
    de2 = torch.ones_like(e)
    ops.infonce(e, 1, n, D, temp, 0.3, acc2, de2, accumulate=accumulate)
    assert abs(acc2.item() - acc.item()) < 2e-4 * max(1, abs(acc2.item()))
    assert (de2 - de).abs().max().item() < 1e-2 * gmax


@pytest.mark.parametrize("decoupled", [0, 1])
def test_adam(decoupled):
    shapes = [(70, 33), (5000,), (3, 4097), (1,)]
    ps = [torch.randn(*s, device=DEV) for s in shapes]
    ref = [p.clone().requires_grad_() for p in ps]
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)(ref, lr=1e-2, weight_decay=0.1)
    grads = [torch.zeros_like(p) for p in ps]
    shadows = [torch.zeros((70, 40), dtype=torch.bfloat16, device=DEV), None, None, None]
    plan = ops.AdamPlan([dict(param=p, grad=g, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p), shadow=s)
                         for p, g, s in zip(ps, grads, shadows)], DEV)
    step = torch.zeros(1, dtype=torch.int64, device=DEV)
    for it in range(3):
        ops.tick(step)
        for g, r in zip(grads, ref):
            g.copy_(torch.randn_like(g))
            r.grad = g.clone()
        opt.step()
        plan.step(1e-2, 0.9, 0.999, 1e-8, 0.1, decoupled, step)
        for p, r, g in zip(ps, ref, grads):
            assert _relerr(p, r.detach()) < 1e-5
            assert g.abs().max().item() == 0
    assert torch.equal(shadows[0][:, :33], ps[0].to(torch.bfloat16))


def test_samplers():
    import scipy.sparse as sp
    U, I, B, n_neg = 200, 300, 5000, 6
    m = sp.random(U, I, density=0.05, format="csr", random_state=2)
    m.sort_indices()
    coo = m.tocoo()
    d = lambda a, t: torch.from_numpy(a.astype(t)).to(DEV)  # noqa: E731
    items = np.arange(0, I, 2)
    step = torch.ones(1, dtype=torch.int64, device=DEV)
    out_u = torch.empty(B, dtype=torch.int64, device=DEV)
    out_i = torch.empty(B, n_neg + 1, dtype=torch.int64, device=DEV)
    ops.sample_batch(d(coo.row, np.int32), d(coo.col, np.int32), d(m.indptr, np.int64), d(m.indices, np.int32),
                     d(items, np.int32), B, n_neg, 3, step, out_u, out_i)
    u, i = out_u.cpu().numpy(), out_i.cpu().numpy()
    dense = m.toarray() != 0
    assert dense[u, i[:, 0]].all()
    assert not dense[u[:, None], i[:, 1:]].any()
    assert np.isin(i[:, 1:], items).all()
    mods = torch.empty(20000 * 2, dtype=torch.uint8, device=DEV)
    ops.sample_modalities(mods, 20000, 2, 4, -1, 5, step)
    mm = mods.view(-1, 2).cpu().numpy()
    assert (mm[:, 0] != mm[:, 1]).all() and mm.max() == 3
    assert np.abs(np.bincount(mm[:, 0], minlength=4) / 20000 - 0.25).max() < 0.02
    ops.sample_modalities(mods, 20000, 2, 4, 2, 5, step)
    mm = mods.view(-1, 2).cpu().numpy()
    assert (mm[:, 0] == 2).all() and (mm[:, 1] != 2).all()


def test_device_batch_feeder_epoch_semantics():
    """DeviceBatchFeeder: an epoch visits every train interaction exactly once (shuffled), the last batch is short,
    negatives are items of the split that are not train positives of the slot's user, two epochs differ in order"""
    from sibrar_b200.synthetic import SynCorpus
    from sibrar_b200.trainer import DeviceBatchFeeder
    train = SynCorpus("ml1m", "cold_start_item", seed=4, scale=0.05).dataset("train")
    feeder = DeviceBatchFeeder(train, batch_size=1000, device=DEV, seed=3, n_negative_samples=5)
    coo = train.interaction_matrix
    csr = train.user_sampling_matrix.tocsr()
    want_pairs = np.sort(coo.row.astype(np.int64) * train.n_items + coo.col.astype(np.int64))
    in_split = np.zeros(train.n_items, dtype=bool)
    in_split[train.items_in_split] = True
    orders = []
    for _ in range(2):
        got_u, got_i, sizes = [], [], []
        for u, i in feeder.epoch():
            got_u.append(u.cpu().numpy())
            got_i.append(i.cpu().numpy())
            sizes.append(len(u))
        assert len(sizes) == len(feeder) and all(s == 1000 for s in sizes[:-1]) and sizes[-1] == coo.nnz - 1000 * (len(sizes) - 1)
        u, i = np.concatenate(got_u), np.concatenate(got_i)
        assert np.array_equal(np.sort(u * train.n_items + i[:, 0]), want_pairs)  # each interaction exactly once
        neg = i[:, 1:]
        assert in_split[neg].all()
        assert not np.asarray(csr[np.repeat(u, neg.shape[1]), neg.reshape(-1)]).any()  # never a train positive
        # roughly uniform over the split's items: no item takes more than 5x its fair share
        counts = np.bincount(neg.reshape(-1), minlength=train.n_items)[train.items_in_split]
        assert counts.max() < 5 * neg.size / len(train.items_in_split) + 20
        orders.append(u * train.n_items + i[:, 0])
    assert not np.array_equal(orders[0], orders[1])
    assert feeder.epochs_done == 2
    with pytest.raises(ValueError):
        train.negative_sampling_strategy = "popular"
        DeviceBatchFeeder(train, batch_size=10, device=DEV)


def test_uniform_sampler_matches_reference_semantics():
    """'uniform' (data/sampling.py:7-32): n_neg DISTINCT items of the split, never a train positive, every eligible item
    equally likely; on a tiny split the empirical distribution is compared with the reference's own sampler
    (numpy statement of negative_sample_uniform) by a chi-square bound"""
    from sibrar_b200.synthetic import SynCorpus
    from sibrar_b200.trainer import DeviceBatchFeeder
    train = SynCorpus("ml1m", "cold_start_item", seed=4, scale=0.02).dataset("train")
    train.negative_sampling_strategy = "uniform"
    n_neg = 7
    feeder = DeviceBatchFeeder(train, batch_size=4096, device=DEV, seed=5, n_negative_samples=n_neg)
    csr = train.user_sampling_matrix.tocsr()
    in_split = np.zeros(train.n_items, dtype=bool)
    in_split[train.items_in_split] = True
    us, negs = [], []
    for _ in range(6):
        for u, i in feeder.epoch():
            us.append(u.cpu().numpy())
            negs.append(i.cpu().numpy()[:, 1:])
    u, neg = np.concatenate(us), np.concatenate(negs)
    assert in_split[neg].all()
    assert not np.asarray(csr[np.repeat(u, n_neg), neg.reshape(-1)]).any()
    srt = np.sort(neg, axis=1)
    assert (srt[:, 1:] != srt[:, :-1]).all()  # distinct inside a slot (np.random.choice(replace=False))
    # the busiest user: every eligible item about equally often
    busy = np.bincount(u).argmax()
    rows = neg[u == busy].reshape(-1)
    eligible = np.setdiff1d(train.items_in_split, csr[busy].indices)
    cnt = np.bincount(rows, minlength=train.n_items)[eligible]
    expected = rows.size / len(eligible)
    chi2 = ((cnt - expected) ** 2 / expected).sum()
    assert chi2 < len(eligible) + 6 * np.sqrt(2 * len(eligible)), (chi2, len(eligible))


@pytest.mark.parametrize("rows,d,C_", [(700, 900, 64), (333, 1500, 512), (260, 300, 20), (1000, 4000, 128)])
def test_spmm_bf16_values_transpose_accumulate(rows, d, C_):
    """sbr_spmm_csr_bf16 (the CSR 'interactions' route): stored values are used (counts > 1), bf16 dense rows, fp32
    accumulation; the transposed output accumulates (wgrad); row subsets with a device-side count"""
    import scipy.sparse as sp
    m = sp.random(rows, d, density=0.03, format="csr", random_state=rows + C_)
    m.data[:] = np.random.default_rng(1).integers(1, 4, size=m.nnz)
    m.sort_indices()
    Cp = ops.pad8(C_)
    Wt = torch.zeros(d, Cp, device=DEV, dtype=torch.bfloat16)
    Wt[:, :C_] = torch.randn(d, C_, device=DEV).to(torch.bfloat16)
    b = torch.randn(C_, device=DEV)
    ip = torch.from_numpy(m.indptr.astype(np.int64)).to(DEV)
    ix = torch.from_numpy(m.indices.astype(np.int32)).to(DEV)
    vals = torch.from_numpy(m.data.astype(np.float32)).to(DEV)
    X = torch.from_numpy(m.toarray()).float().to(DEV)
    out = torch.empty(rows, C_, device=DEV)
    out16 = torch.full((rows, Cp), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.spmm_csr(ip, ix, rows, Wt, C_, b, "relu", out, vals=vals, out_bf16=out16)
    ref = torch.relu(X @ Wt[:, :C_].float() + b)
    assert _relerr(out, ref) < 1e-5
    assert _relerr(out16[:, :C_].float(), ref) < 5e-3 and (out16[:, C_:] == 0).all()
    # all-ones shortcut
    ops.spmm_csr(ip, ix, rows, Wt, C_, None, None, out)
    assert _relerr(out, (X > 0).float() @ Wt[:, :C_].float()) < 1e-5
    # wgrad through the transposed matrix: dW[:, j] += sum_r X[r, j] dz[r, :], twice (accumulation)
    mt = m.T.tocsr()
    mt.sort_indices()
    ipt = torch.from_numpy(mt.indptr.astype(np.int64)).to(DEV)
    ixt = torch.from_numpy(mt.indices.astype(np.int32)).to(DEV)
    valst = torch.from_numpy(mt.data.astype(np.float32)).to(DEV)
    dz = torch.zeros(rows, Cp, device=DEV, dtype=torch.bfloat16)
    dz[:, :C_] = torch.randn(rows, C_, device=DEV).to(torch.bfloat16)
    gw = torch.zeros(C_, d, device=DEV)
    for _ in range(2):
        ops.spmm_csr(ipt, ixt, d, dz, C_, None, None, gw, transpose_out=True, vals=valst, accumulate=True)
    assert _relerr(gw, 2 * (dz[:, :C_].float().T @ X)) < 1e-5
    # row subset: only the listed rows are computed / written, the count lives on the device
    sel = torch.randperm(rows, device=DEV)[:rows // 3].to(torch.int32)
    lst = torch.cat([sel, torch.full((17,), -7, dtype=torch.int32, device=DEV)])  # tail beyond the count: ignored
    n_dev = torch.tensor([sel.numel()], dtype=torch.int32, device=DEV)
    sub = torch.full((rows, C_), -1.0, device=DEV)
    ops.spmm_csr(ip, ix, lst.numel(), Wt, C_, b, "relu", sub, vals=vals, row_list=lst, n_rows_dev=n_dev)
    mask = torch.zeros(rows, dtype=torch.bool, device=DEV)
    mask[sel.long()] = True
    assert _relerr(sub[mask], ref[mask]) < 1e-5 and (sub[~mask] == -1).all()


@pytest.mark.parametrize("n_rows,n_tags,max_tags,C", [(5000, 853, 7, 512), (900, 60, 4, 64)])
def test_tag_bag_bwd_segment_route(n_rows, n_tags, max_tags, C):
    """EmbeddingBag(mean) weight gradient of a LARGE vocabulary as per-tag gather-sums over <= 128-row segments
    (DeviceFeature.tag_segments + sbr_spmm_csr segment mode) == the per-row scatter kernel"""
    from types import SimpleNamespace

    from sibrar_b200.feature_store import DeviceFeature
    rng = np.random.default_rng(n_rows)
    codes = np.full((n_rows, max_tags), n_tags, dtype=np.int64)
    for r in range(n_rows):
        n = int(rng.integers(0, max_tags + 1))
        codes[r, :n] = rng.permutation(n_tags)[:n]
    codes[:300, 0] = 3  # one very frequent tag: several segments
    feat = SimpleNamespace(feature_definition=SimpleNamespace(name="t", type="tag"), values=codes,
                           _indices=np.arange(n_rows), dim=n_tags)
    df = DeviceFeature("t", feat, n_rows, DEV)
    seg_ptr, seg_rows, seg_vals, seg_tag = df.tag_segments
    assert int((seg_ptr[1:] - seg_ptr[:-1]).max()) <= 128
    dy = torch.randn(n_rows, C, device=DEV)
    want = torch.zeros(n_tags + 1, C, device=DEV)
    ops.tag_bag_bwd(df.codes, df.max_tags, df.pad_id, dy.clone(), want)
    got = torch.zeros(n_tags + 1, C, device=DEV)
    ops.spmm_csr(seg_ptr, seg_rows, seg_tag.numel(), dy, C, None, None, got, vals=seg_vals, row_map=seg_tag, atomic=True)
    assert _relerr(got, want) < 1e-5 and (got[n_tags] == 0).all()


def test_adagrad():
    """sbr_adam_step mode 2 == torch.optim.Adagrad(lr, weight_decay) (train/trainer.py:62-68)"""
    shapes = [(70, 33), (5000,), (3, 4097), (1,)]
    ps = [torch.randn(*s, device=DEV) for s in shapes]
    ref = [p.clone().requires_grad_() for p in ps]
    opt = torch.optim.Adagrad(ref, lr=1e-2, weight_decay=0.1)
    grads = [torch.zeros_like(p) for p in ps]
    shadows = [torch.zeros((70, 40), dtype=torch.bfloat16, device=DEV), None, None, None]
    plan = ops.AdamPlan([dict(param=p, grad=g, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p), shadow=s)
                         for p, g, s in zip(ps, grads, shadows)], DEV)
    step = torch.zeros(1, dtype=torch.int64, device=DEV)
    for it in range(3):
        ops.tick(step)
        for g, r in zip(grads, ref):
            g.copy_(torch.randn_like(g))
            r.grad = g.clone()
        opt.step()
        plan.step(1e-2, 0.9, 0.999, 1e-10, 0.1, 2, step)
        for p, r, g in zip(ps, ref, grads):
            assert _relerr(p, r.detach()) < 1e-5
            assert g.abs().max().item() == 0
    assert torch.equal(shadows[0][:, :33], ps[0].to(torch.bfloat16))


def test_spmm_bf16_segmented_long_rows():
    """rows longer than 256 entries are cut into segments (feature_store.csr_segments): same results as the plain
    kernel for the forward (bias + activation applied once per row, bf16 copy included) and the accumulating wgrad"""
    import scipy.sparse as sp
    from sibrar_b200.feature_store import csr_segments
    rows, d, C_ = 400, 3000, 128
    m = sp.random(rows, d, density=0.01, format="lil", random_state=5)
    rng = np.random.default_rng(5)
    for r, n in ((0, 2900), (7, 700), (399, 257)):  # a popular item, a heavy user, one entry over the limit
        m[r, rng.choice(d, size=n, replace=False)] = 1
    m = m.tocsr()
    m.data[:] = 1
    m.sort_indices()
    seg = csr_segments(m.indptr, DEV)
    assert seg is not None and seg[2].cpu().tolist() == [0, 7, 399]
    assert int((seg[0][1:] - seg[0][:-1]).max()) <= 256
    Wt = torch.randn(d, C_, device=DEV).to(torch.bfloat16)
    b = torch.randn(C_, device=DEV)
    ip = torch.from_numpy(m.indptr.astype(np.int64)).to(DEV)
    ix = torch.from_numpy(m.indices.astype(np.int32)).to(DEV)
    X = torch.from_numpy(m.toarray()).float().to(DEV)
    ref = torch.relu(X @ Wt.float() + b)
    out = torch.full((rows, C_), float("nan"), device=DEV)
    out16 = torch.full((rows, C_), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.spmm_csr(ip, ix, rows, Wt, C_, b, "relu", out, out_bf16=out16, segments=seg)
    assert _relerr(out, ref) < 1e-5 and _relerr(out16.float(), ref) < 5e-3
    mt = m.T.tocsr()
    mt.sort_indices()
    segt = csr_segments(mt.indptr, DEV, seg=4)  # (short segments: many "long" rows on the transposed side too)
    assert segt is not None
    ipt = torch.from_numpy(mt.indptr.astype(np.int64)).to(DEV)
    ixt = torch.from_numpy(mt.indices.astype(np.int32)).to(DEV)
    dz = torch.randn(rows, C_, device=DEV).to(torch.bfloat16)
    gw = torch.zeros(C_, d, device=DEV)
    for _ in range(2):
        ops.spmm_csr(ipt, ixt, d, dz, C_, None, None, gw, transpose_out=True, accumulate=True, segments=segt)
    assert _relerr(gw, 2 * (dz.float().T @ X)) < 1e-5
