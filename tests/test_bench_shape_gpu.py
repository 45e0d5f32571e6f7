"""Parity of ONE FULL TRAIN STEP AT THE BENCHMARKED SHAPE (BASELINE configs[1]: synthetic ML-1M, 6 040 x 3 706, cold-start
item split, C = D = 64, BPR, AdamW) through the path ``bench.py`` times -- ``FusedTrainer(cuda_graph=True)``: split-K
projection GEMMs, the fused score/loss + trailing-BatchNorm kernel, the persistent GEMMs, graph replay with programmatic
dependent launch -- against the numpy oracle run with the kernels' rounding points (``Bf16Emulation``).

The modalities and dropout masks the step drew on the device (Philox) are read back and handed to the oracle; the
gradients are the snapshot the step takes right before the fused AdamW consumes them.  ENTRY-WISE criteria, no direction
escape: logits 5e-3 of max-abs, loss 5e-4 relative, every gradient tensor within 1e-2 of its max-abs.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import sbnet_oracle as O  # noqa: E402
from sibrar_b200 import workloads  # noqa: E402
from sibrar_b200.sbnet import SingleBranchNet  # noqa: E402
from sibrar_b200.synthetic import sample_batch  # noqa: E402
from sibrar_b200.trainer import FusedTrainer  # noqa: E402

DEV = "cuda"
_CORPUS = {}


def _corpus():
    if "c" not in _CORPUS:
        _CORPUS["c"] = workloads.build("ml1m")[0]
    return _CORPUS["c"]


def _unpack_keep(bits: torch.Tensor, C: int) -> np.ndarray:
    """uint8 [N, ceil(C / 8)] keep bits written by the gather kernel -> float32 [N, C] (1 = kept)"""
    b = bits.cpu().numpy()
    return np.unpackbits(b, axis=1, bitorder="little")[:, :C].astype(np.float32)


# (300, 1000: row counts that are NOT multiples of the fused kernels' 128-row tiles -- a partial last tile behind full ones,
# on the user side (300 / 1000 rows) and on the item side (3 300 / 11 000 rows))
@pytest.mark.parametrize("batch,batch_norm", [(256, True), (300, True), (1000, False), (16384, True), (16384, False)])
def test_graph_step_at_bench_shape_matches_oracle(batch, batch_norm):
    corpus = _corpus()
    train = corpus.dataset("train")
    conf = workloads.ml1m_conf(batch_norm=batch_norm)
    learn = dict(workloads.WORKLOADS["ml1m"]["learn"])
    torch.manual_seed(1234)
    model = SingleBranchNet.build_from_conf(conf, train).to(DEV).train()
    tr = FusedTrainer(model, learn, n_negative_samples=workloads.N_NEG, cuda_graph=True)
    tr.snapshot_grads = True
    rng = np.random.default_rng(17)
    n_steps = 5  # 2 eager, capture + first replay, two more replays; the last one is checked
    for s in range(n_steps):
        u, i = sample_batch(train, batch, rng, workloads.N_NEG)
        ut, it = torch.from_numpy(u).to(DEV), torch.from_numpy(i).to(DEV)
        if s == n_steps - 1:
            torch.cuda.synchronize()
            before = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
            tr.read_losses()
        tr.step(ut, it)
    torch.cuda.synchronize()
    model.check_errors()
    key = ((batch,), (batch, 1 + workloads.N_NEG))
    assert "graph" in tr._graphs[key], "the checked step must be a CUDA-graph replay"
    losses = tr.read_losses()
    ent_u, ent_i = model.user_embedding_module, model.item_embedding_module
    mods = {"user": ent_u._ctx[1].cpu().numpy().astype(np.int64).reshape(batch, 1),
            "item": ent_i._ctx[1].cpu().numpy().astype(np.int64).reshape(batch, 1 + workloads.N_NEG, 1)}
    names = {"user": ent_u.mod_names, "item": ent_i.mod_names}
    drop = {"item": _unpack_keep(ent_i.dropout_keep_bits(), 64)}
    assert 0.75 < drop["item"].mean() < 0.85  # p = 0.2
    p64 = {k: v.astype(np.float64) if v.dtype.kind == "f" else v for k, v in before.items()}
    net = O.OracleSBNet(conf, train)
    ref = net.train_step_fwd_bwd(p64, u, i, mods, names, drop, loss_kind="bpr", emu=O.Bf16Emulation())
    logits = tr.logits.cpu().numpy()
    lerr = np.abs(logits - ref["logits"]).max() / np.abs(ref["logits"]).max()
    assert lerr < 5e-3, f"logits differ from the bf16-emulating oracle by {lerr:.2e} of max-abs"
    assert losses["train/loss"] == pytest.approx(ref["loss"], rel=5e-4)
    snap = tr.grads_snapshot.cpu().numpy()
    bad = {}
    for name, prm in model.named_parameters():
        off, numel, shape = tr.grad_offsets[id(prm)]
        got = snap[off:off + numel].reshape(shape)
        want = ref["grads"].get(name)
        if want is None:
            want = np.zeros(shape)
        gmax = float(np.abs(want).max())
        err = float(np.abs(got - want).max())
        if err > 1e-2 * gmax + 1e-7:
            bad[name] = (err, gmax)
    assert not bad, f"gradients differ entry-wise from the bf16-emulating oracle: {bad}"
    # the optimizer consumed exactly these gradients: Adam state after the step is finite and the weights moved
    after = model.state_dict()
    moved = [float((after[k].cpu() - torch.from_numpy(before[k])).abs().max()) for k in before
             if before[k].dtype.kind == "f" and "running" not in k]
    assert max(moved) > 0 and np.isfinite(moved).all()
