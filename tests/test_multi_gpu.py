"""Hardware test of the multi-GPU classes (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py
-m gpu`): `DataParallelTrainer` on split halves == `FusedTrainer` on the concatenated batch (BatchNorm-free model, and a
BatchNorm model with sync_bn=True), `ShardedEvaluator` == `FullEvaluator` bit-exact on positions."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_and_sharded_eval_match_single_gpu(tmp_path):
    out = str(tmp_path / "mgpu.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tests", "mgpu_worker.py"), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    print(res)
    # A. BatchNorm-free: the two half-batch steps ARE the full-batch step (fp32 summation order aside)
    assert res["dp_no_bn/grad_max_err_rel"] < 2e-3
    assert abs(res["dp_no_bn/loss_rel_err"] - res["dp_no_bn/ref_loss"]) < 1e-5 * abs(res["dp_no_bn/ref_loss"]) + 1e-6
    assert res["dp_no_bn/param_spread_over_ranks"] == 0.0
    assert res["dp_no_bn/collective_vs_nccl_max_err_rel"] < 1e-5  # the trainer's own collective == NCCL's sum
    # B. BatchNorm with global statistics
    assert res["dp_sync_bn/grad_max_err_rel"] < 5e-3
    assert abs(res["dp_sync_bn/loss_rel_err"] - res["dp_sync_bn/ref_loss"]) < 1e-4 * abs(res["dp_sync_bn/ref_loss"]) + 1e-6
    assert res["dp_sync_bn/running_stats_max_err"] < 1e-5
    assert res["dp_sync_bn/param_spread_over_ranks"] == 0.0
    # C. item-sharded evaluation
    assert res["eval/positions_equal"] and res["eval/scores_max_err"] == 0.0 and res["eval/metrics_max_err"] < 1e-6
