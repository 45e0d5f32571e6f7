"""CPU test of the host side of the drop-in (no GPU, no kernels): state_dict keys and shapes equal the reference's,
config parsing / error behaviour mirror the reference, and the whole fused-step / evaluation call sequence runs
against a recording stub of the C ABI that checks every call's argument count and types against the header
prototypes."""
import ctypes

import numpy as np
import pytest
import torch

from sibrar_b200 import _lib, ops
import sibrar_b200.sbnet as sbnet
from sibrar_b200.config import SingleBranchNetConfig, SingleBranchNetEntityConfig, FeatureModuleConfig
from sibrar_b200.sbnet import SingleBranchNet
from sibrar_b200.synthetic import SynCorpus
from tests.golden_util import CASES, load_case, state_dict_of, step_inputs


@pytest.mark.parametrize("name", list(CASES))
def test_state_dict_keys_and_shapes_match_reference(name):
    spec, g, corpus = load_case(name)
    model = SingleBranchNet.build_from_conf(spec["model"], corpus.dataset("train"))
    want = {k: v.shape for k, v in state_dict_of(g, "sd0/").items()}
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}


def test_config_parsing_follows_reference_rules():
    spec = CASES["plain_user_ssm"]["model"]
    cfg = SingleBranchNetConfig.from_dict(spec)
    assert isinstance(cfg.user, FeatureModuleConfig) and isinstance(cfg.item, SingleBranchNetEntityConfig)
    assert not cfg.is_user_sb_module and cfg.is_item_sb_module
    assert cfg.item.eval_modalities == {"title_mpnet", "image_resnet"}
    assert cfg.item.apply_batch_normalization and cfg.item.apply_batch_norm_every == 0


def test_reference_error_behaviour():
    corpus = SynCorpus("ml1m", "random", seed=1, scale=0.02, vector_dim_cap=8)
    base = CASES["ml1m_small"]["model"]

    def build(**item_over):
        conf = {**base, "item": {**base["item"], **item_over}}
        return SingleBranchNet.build_from_conf(conf, corpus.dataset("train"))

    with pytest.raises(ValueError, match="not available"):
        build(train_modalities=["does_not_exist"])
    with pytest.raises(ValueError, match="Cannot use modality"):
        build(train_modalities=["genres"], eval_modalities=["plot_mpnet"])
    with pytest.raises(ValueError, match="Aggregation function"):
        build(aggregation_fn="median")
    with pytest.raises(ValueError, match="at least one feature"):
        build(features=[])
    with pytest.raises(ValueError):
        build(embedding_regularization_type="no_such_type")
    model = build().eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        model.get_item_representations(torch.arange(4))


class _Stub:
    """stands in for the shared library: validates each call against the ctypes prototypes"""

    def __init__(self):
        self.calls = []

    def __call__(self, name, *args):
        protos = _lib._PROTOS[name]
        assert len(args) == len(protos), (name, len(args), len(protos))
        for a, t in zip(args, protos):
            if a is None:
                continue
            if t in (_lib.c_i64, _lib.c_i32, ctypes.c_int, _lib.c_u64):
                assert isinstance(a, (int, np.integer)) and not isinstance(a, bool), (name, t, type(a))
            elif t is _lib.c_f32:
                assert isinstance(a, (float, int)), (name, t, type(a))
        if name == "sbr_topk_workspace_bytes":
            args[-1]._obj.value = 4096
        self.calls.append(name)


@pytest.mark.parametrize("name", list(CASES))
def test_fused_step_and_eval_call_sequence(name, monkeypatch):
    stub = _Stub()
    monkeypatch.setattr(ops, "call", stub)
    monkeypatch.setattr(ops, "stream_ptr", lambda: 0)
    monkeypatch.setattr(torch.cuda, "get_device_properties",
                        lambda d: type("P", (), {"multi_processor_count": 148})())

    def _rt(self):
        if self._runtime is None:
            self._runtime = sbnet._Runtime(torch.device("cpu"))
        return self._runtime
    monkeypatch.setattr(SingleBranchNet, "_rt", _rt)
    from sibrar_b200.evaluator import FullEvaluator
    from sibrar_b200.trainer import FusedTrainer

    spec, g, corpus = load_case(name)
    model = SingleBranchNet.build_from_conf(spec["model"], corpus.dataset("train")).train()
    tr = FusedTrainer(model, dict(lr=spec["lr"], wd=spec["wd"], optimizer=spec["optimizer"], rec_loss=spec["rec_loss"]),
                      n_negative_samples=spec["n_neg"])
    u, i, _, _, _ = step_inputs(g, 0)
    stub.calls.clear()  # construction builds the feature store and the bf16 weight shadows
    tr.step(torch.from_numpy(u), torch.from_numpy(i))
    seq = stub.calls
    assert seq[0] == "sbr_step_begin" and seq[-1] == "sbr_adam_step" and "sbr_tick" not in seq
    score = "sbr_score_loss_bn" if "sbr_score_loss_bn" in seq else "sbr_score_loss"
    # entities whose single-branch net is 1-2 Linear layers of width <= 64 run the fused gather + MLP kernels
    seq = ["sbr_mlp2_fwd" if c == "sbr_mlp2_fwd_bn" else c for c in seq]  # (= the forward + in-kernel BatchNorm finalize)
    fused = "sbr_mlp2_fwd" in seq
    assert score in seq and ("sbr_row_gather_fwd" in seq or fused) and "sbr_row_gather_bwd_segmented" in seq
    assert seq.index("sbr_gather_plan") < seq.index("sbr_row_gather_bwd_segmented")
    if fused:
        assert seq.count("sbr_mlp2_fwd") == seq.count("sbr_mlp2_bwd")
        assert seq.index("sbr_mlp2_fwd") < seq.index(score) < seq.index("sbr_mlp2_bwd")
        last_bwd = len(seq) - 1 - seq[::-1].index("sbr_row_gather_bwd_segmented")
        assert seq.index("sbr_mlp2_bwd") < last_bwd  # dX0 of the fused backward feeds the sorted-run gather backward
    if "sbr_row_gather_fwd" in seq:
        assert seq.count("sbr_gemm_bf16") >= 3  # forward, wgrad, dgrad GEMMs on the tensor cores
    assert ("sbr_infonce" in seq) == (model.user_embedding_module.reg_enabled or model.item_embedding_module.reg_enabled)
    assert seq.index(score) < seq.index("sbr_row_gather_bwd_segmented") < seq.index("sbr_adam_step")
    assert set(tr.read_losses()) >= {"train/loss", "train/rec_loss", "train/reg_loss"}
    stub.calls.clear()
    res = FullEvaluator(dict(top_k=[1, 3, 5], metrics=["ndcg", "recall", "coverage"], calculate_std=True)).evaluate(
        model, corpus.dataset("val"))
    assert "sbr_topk_scores_masked" in stub.calls and "sbr_topk_merge" in stub.calls and "sbr_metrics_at_k" in stub.calls
    assert list(res) == ["coverage@1", "coverage@3", "coverage@5", "ndcg@1", "ndcg@1_std", "ndcg@3", "ndcg@3_std",
                         "ndcg@5", "ndcg@5_std", "recall@1", "recall@1_std", "recall@3", "recall@3_std", "recall@5",
                         "recall@5_std"]
    assert model.training  # evaluate() restores the mode it found


@pytest.mark.parametrize("rows,cols", [(5, 1), (7, 64), (9, 130), (33, 6040)])
def test_pack_bits_layout(rows, cols):
    """host layout of the bit-packed multi-hot operand (`include/sibrar_b200.h`, sbr_gemm_bits_bf16): bit k of row m =
    word k // 32, bit k % 32; 16-byte row pitch (the kernel fetches the words by TMA), whole 64-bit K blocks, zero
    padding -- the dense matrix the reference builds per batch (data/Feature.py:147-150) is recovered bit for bit"""
    import scipy.sparse as sp
    m = sp.random(rows, cols, density=0.3, format="csr", random_state=rows * 1000 + cols)
    m.data[:] = 1
    bits = ops.pack_bits(m, "cpu")
    assert bits.dtype == torch.int32 and bits.shape[0] == rows
    ld_words = bits.shape[1]
    assert ld_words % 4 == 0 and ld_words * 32 >= 64 * ((cols + 63) // 64) and bits.stride(0) == ld_words
    words = bits.numpy().view(np.uint32)
    dense = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(rows, ld_words * 32)
    assert np.array_equal(dense[:, :cols], m.toarray().astype(np.uint32))
    assert not dense[:, cols:].any()
