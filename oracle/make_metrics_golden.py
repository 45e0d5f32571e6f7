"""TEST INFRASTRUCTURE ONLY -- pins the metric arithmetic against code the reference itself holds.

``rmet`` (the package ``eval/eval.py:99-102`` calls) is not vendored, but the reference tree contains its own statement of
three of the metrics: ``eval/metrics.py:4-105`` (``recall_at_k_batch``, ``precision_at_k_batch``, ``ndcg_at_k_batch``),
importable with no shim.  This script runs THOSE functions, unmodified, on seeded tie-free logits and binary targets and
writes inputs + per-user outputs to ``tests/golden/metrics_pin.npz``.  ``tests/test_oracle_vs_golden.py`` checks
``oracle/rmet_restated.py`` against it on CPU and ``tests/test_golden_gpu.py`` checks ``sbr_metrics_at_k`` on the GPU.

Run in the build container (needs ``/root/reference``):  ``python -m oracle.make_metrics_golden``.

Cases (rows = users): ordinary users, users with no target at all (NaN -> 0 in the reference), users with more targets
than k, users whose targets are all ranked first (NDCG = 1 exactly), users with every item a target.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KS = (1, 3, 5, 10, 20)


def _reference_metrics():
    path = os.path.join("/root/reference", "eval", "metrics.py")
    spec = importlib.util.spec_from_file_location("ref_eval_metrics", path)
    mod = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(mod)
    return mod


def main():
    import torch
    ref = _reference_metrics()
    rng = np.random.default_rng(2024)
    U, I = 96, 157
    logits = rng.standard_normal((U, I)).astype(np.float32)
    logits += (np.arange(I, dtype=np.float32) * 2.0 ** -12)[None, :]  # tie-free
    targets = (rng.random((U, I)) < 0.04).astype(np.float32)
    targets[0] = 0.                       # no targets: recall / ndcg are NaN -> 0 in the reference
    targets[1] = 0.
    targets[2] = (rng.random(I) < 0.5)    # many more targets than k
    targets[3] = 1.                       # everything is a target
    order4 = np.argsort(-logits[4])
    targets[4] = 0.
    targets[4, order4[:6]] = 1.           # targets ranked first: ndcg == 1 up to k = 5, < 1 never
    targets[5] = 0.
    targets[5, order4[0]] = 1.            # single target (some rank of user 5's own list)
    out = {"logits": logits, "targets": targets.astype(np.uint8), "ks": np.asarray(KS)}
    lt, tt = torch.from_numpy(logits), torch.from_numpy(targets)
    for k in KS:
        idx = lt.topk(k=k).indices
        out[f"topk@{k}"] = idx.numpy()
        out[f"recall@{k}"] = ref.recall_at_k_batch(lt, tt, k=k, aggr_sum=False, idx_topk=idx).numpy()
        out[f"precision@{k}"] = ref.precision_at_k_batch(lt, tt, k=k, aggr_sum=False, idx_topk=idx).numpy()
        out[f"ndcg@{k}"] = ref.ndcg_at_k_batch(lt, tt, k=k, aggr_sum=False, idx_topk=idx).numpy()
    path = os.path.join(ROOT, "tests", "golden", "metrics_pin.npz")
    np.savez_compressed(path, **out)
    print(f"[golden] metrics pin: {len(out)} arrays -> {path} ({os.path.getsize(path) / 1024:.0f} KiB); "
          f"mean ndcg@10 {out['ndcg@10'].mean():.5f}")


if __name__ == "__main__":
    main()
