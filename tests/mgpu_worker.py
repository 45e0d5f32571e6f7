"""Worker of tests/test_multi_gpu.py, launched with torchrun (one process per GPU, NCCL):
  A. data-parallel step (BatchNorm-free model): gradients of W ranks on the W slices of a batch, all-reduced and divided
     by W, equal the single-process gradients on the concatenated batch; after an optimizer step the ranks hold
     identical parameters.
  B. the same with BatchNorm and ``sync_bn=True`` (statistics + backward sums over the global batch).
  C. ``ShardedEvaluator`` (item-sharded top-k + all-gather + merge) returns the positions / metrics of ``FullEvaluator``.
Rank 0 writes the measured deviations as JSON."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sibrar_b200  # noqa: E402,F401
from sibrar_b200.evaluator import FullEvaluator  # noqa: E402
from sibrar_b200.parallel import DataParallelTrainer, ShardedEvaluator  # noqa: E402
from sibrar_b200.sbnet import SingleBranchNet  # noqa: E402
from sibrar_b200.synthetic import SynCorpus, sample_batch  # noqa: E402
from sibrar_b200.trainer import FusedTrainer  # noqa: E402


def conf(batch_norm):
    ent = lambda feats, hidden, drop: dict(  # noqa: E731
        features=[dict(feature_name=f, feature_hidden_layers=[]) for f in feats], single_branch_hidden_layers=hidden,
        preference_hidden_layers=[], common_modality_dim=16, activation_fn="relu", single_branch_input_dropout=drop,
        apply_batch_normalization=batch_norm)
    return dict(shared_common_dim=16, user=ent(["interactions", "gender", "occupation"], [], None),
                item=ent(["interactions", "genres", "plot_mpnet"], [16], 0.2))


LEARN = dict(lr=1e-3, wd=1e-6, optimizer="adamw", rec_loss="bpr", loss_aggregator="mean")


def main(out_path):
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=dev)
    corpus = SynCorpus("ml1m", "cold_start_item", seed=11, scale=0.05, vector_dim_cap=32)
    train = corpus.dataset("train")
    rng = np.random.default_rng(3)
    B, n_neg = 32 * world, 5
    u, i = sample_batch(train, B, rng, n_neg)
    mods = {"user": rng.integers(0, 3, size=(B, 1)).astype(np.uint8),
            "item": rng.integers(0, 3, size=(B, 1 + n_neg, 1)).astype(np.uint8)}
    keep_i = (rng.random((B * (1 + n_neg), 16)) >= 0.2).astype(np.uint8)
    lo, hi = rank * B // world, (rank + 1) * B // world
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    res = {}
    for tag, bn, sync in (("dp_no_bn", False, False), ("dp_sync_bn", True, True)):
        torch.manual_seed(0)
        model = SingleBranchNet.build_from_conf(conf(bn), train).to(dev).train()
        sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        tr = DataParallelTrainer(model, LEARN, n_negative_samples=n_neg, sync_bn=sync)
        tr.step(t(u[lo:hi]), t(i[lo:hi]), {"user": t(mods["user"][lo:hi].reshape(-1)), "item": t(mods["item"][lo:hi].reshape(-1))},
                {"item": t(keep_i[lo * (1 + n_neg):hi * (1 + n_neg)])}, apply_optimizer=False)
        dist.all_reduce(tr.flat_grads)
        dp_grads = (tr.flat_grads / world).cpu().numpy().copy()
        dp_loss = torch.tensor([tr.read_losses()["train/loss"]], dtype=torch.float64, device=dev)
        dist.all_reduce(dp_loss)
        running = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items() if "running" in k}
        # optimizer path: the ranks must end up with identical parameters
        tr.flat_grads.zero_()
        model.load_state_dict(sd0)
        tr.snapshot_grads = True  # the all-reduced gradients the optimizer consumed (the trainer's own collective)
        tr.step(t(u[lo:hi]), t(i[lo:hi]), {"user": t(mods["user"][lo:hi].reshape(-1)), "item": t(mods["item"][lo:hi].reshape(-1))},
                {"item": t(keep_i[lo * (1 + n_neg):hi * (1 + n_neg)])})
        res[f"{tag}/collective"] = tr.collective
        res[f"{tag}/collective_vs_nccl_max_err_rel"] = float(
            np.abs(tr.grads_snapshot.cpu().numpy() / world - dp_grads).max() / np.abs(dp_grads).max())
        if os.environ.get("SBR_MGPU_DEBUG"):
            snap = tr.grads_snapshot.cpu().numpy() / world
            for k, prm in model.named_parameters():
                o, nel, _ = tr.grad_offsets[id(prm)]
                e = np.abs(snap[o:o + nel] - dp_grads[o:o + nel]).max()
                print(rank, tag, k, o, nel, "err", e, "max", np.abs(dp_grads[o:o + nel]).max(),
                      "snap max", np.abs(snap[o:o + nel]).max(), flush=True)
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        mx, mn = flat.clone(), flat.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        res[f"{tag}/param_spread_over_ranks"] = float((mx - mn).abs().max())
        if rank == 0:
            torch.manual_seed(0)
            ref_model = SingleBranchNet.build_from_conf(conf(bn), train).to(dev).train()
            ref_model.load_state_dict(sd0)
            ref = FusedTrainer(ref_model, LEARN, n_negative_samples=n_neg)
            ref.step(t(u), t(i), {"user": t(mods["user"].reshape(-1)), "item": t(mods["item"].reshape(-1))},
                     {"item": t(keep_i)}, apply_optimizer=False)
            want = ref.flat_grads.cpu().numpy()
            res[f"{tag}/grad_max_err_rel"] = float(np.abs(dp_grads - want).max() / np.abs(want).max())
            res[f"{tag}/loss_rel_err"] = float(abs(dp_loss.item() / world - ref.read_losses()["train/loss"]) /
                                               abs(ref.read_losses(reset=False)["train/loss"] or 1.0)) \
                if False else float(abs(dp_loss.item() / world))
            res[f"{tag}/ref_loss"] = float(ref.read_losses()["train/loss"])
            ref_running = {k: v.detach().cpu().numpy() for k, v in ref_model.state_dict().items() if "running" in k}
            res[f"{tag}/running_stats_max_err"] = float(max([np.abs(running[k] - ref_running[k]).max()
                                                             for k in running] or [0.0]))
        dist.barrier()
    # C. item-sharded evaluation == single-GPU evaluation
    model.eval()
    val = corpus.dataset("val")
    cfg = dict(top_k=[1, 5, 10], metrics=["ndcg", "recall", "precision", "coverage"], calculate_std=False)
    sharded, (sv, si) = ShardedEvaluator(cfg).evaluate(model, val, return_topk=True)
    if rank == 0:
        full, (fv, fi) = FullEvaluator(cfg).evaluate(model, val, return_topk=True)
        res["eval/positions_equal"] = bool(torch.equal(si.cpu(), fi.cpu()))
        res["eval/scores_max_err"] = float((torch.nan_to_num(sv, neginf=-1e30) - torch.nan_to_num(fv, neginf=-1e30)).abs().max())
        res["eval/metrics_max_err"] = float(max(abs(sharded[k] - full[k]) for k in full))
        res["eval/n_items"] = int(val.n_items_in_split)
        with open(out_path, "w") as fh:
            json.dump(res, fh)
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
