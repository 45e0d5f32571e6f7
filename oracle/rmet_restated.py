"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the third-party ``rmet`` package used by the reference.

``rmet`` (git+https://github.com/tigxy/recommender-metrics.git, unpinned HEAD, reference ``environment.yml:40``) is not
vendored in the reference tree and is not installable here (no network).  PINNING STATUS:
  * ndcg / recall / precision are PINNED against code the reference itself holds: ``oracle/make_metrics_golden.py`` runs
    the unmodified ``eval/metrics.py:4-105`` and ``tests/test_oracle_vs_golden.py::
    test_metric_restatement_pinned_against_reference_metrics`` compares this file with its per-user vectors;
  * f_score / hitrate / ap / rr / coverage stay PARITY UNPINNED (no statement of them exists in the reference tree;
    standard definitions, listed at the end of this docstring).
This file restates the package's behaviour, anchored on

* the reference's call sites ``eval/eval.py:99-102`` (overall metrics + top-k indices), ``:115-118`` (group
  metrics), ``:141-144`` (distribution metrics from stored top-k), the key contract ``'{metric}@{k}'`` with an
  ``_individual`` variant that ``eval/eval.py:94-96`` strips, optional ``'{prefix}/'``;
* the only in-repo statement of the metric arithmetic, ``eval/metrics.py:4-105``:
  recall = hits / n_targets (NaN -> 0), precision = hits / k, NDCG = DCG / IDCG with discounts
  ``1/log2(arange(2, k+2))``, IDCG from ``y_true.topk(k)``, NaN -> 0, clamp <= 1;
* ``torch.topk(logits, k, largest=True, sorted=True)`` for the ranking (``eval/eval.py:297``).

hitrate = min(hits, 1); f_score = 2pr/(p+r) (0 when p+r == 0); rr = 1/rank of first hit (0 if none);
ap = mean over hits of precision@rank / min(k, n_targets); coverage = |unique top-k items| / n_items.
"""
from __future__ import annotations

import numpy as np
import torch

supported_user_metrics = ["ndcg", "precision", "recall", "f_score", "hitrate", "ap", "rr"]
supported_distribution_metrics = ["coverage"]
supported_metrics = supported_user_metrics + supported_distribution_metrics


class UserFeature:
    def __init__(self, name, labels):
        self.name = name
        self.labels = np.asarray(labels)


def _user_metric(name: str, rel: torch.Tensor, n_targets: torch.Tensor, k: int) -> torch.Tensor:
    """rel: [U, k] 0/1 relevance of the ranked top-k; n_targets: [U]."""
    rel = rel[:, :k].double()
    hits = rel.sum(-1)
    nt = n_targets.double()
    if name == "precision":
        return hits / k
    if name == "recall":
        r = hits / nt
        r[nt == 0] = 0.
        return r
    if name == "hitrate":
        return hits.clamp(max=1.)
    if name == "f_score":
        p = hits / k
        r = hits / nt
        r[nt == 0] = 0.
        f = 2 * p * r / (p + r)
        f[(p + r) == 0] = 0.
        return f
    if name == "ndcg":
        disc = 1. / torch.log2(torch.arange(2, k + 2, dtype=torch.float32)).double()
        dcg = (rel * disc).sum(-1)
        ideal = (torch.arange(k)[None, :] < nt.clamp(max=k)[:, None]).double()
        idcg = (ideal * disc).sum(-1)
        nd = dcg / idcg
        nd[idcg == 0] = 0.
        return nd.clamp(max=1.)
    if name == "rr":
        first = torch.where(rel.sum(-1) > 0, rel.argmax(-1) + 1, torch.zeros_like(hits, dtype=torch.long))
        rr = torch.where(first > 0, 1. / first.double(), torch.zeros_like(hits))
        return rr
    if name == "ap":
        prec_at = rel.cumsum(-1) / torch.arange(1, k + 1).double()
        denom = nt.clamp(max=k)
        ap = (prec_at * rel).sum(-1) / denom
        ap[denom == 0] = 0.
        return ap
    raise ValueError(f"unsupported metric {name}")


def _key(prefix, metric, k, individual=False):
    key = f"{metric}{'_individual' if individual else ''}@{k}"
    return f"{prefix}/{key}" if prefix else key


def calculate(metrics, logits: torch.Tensor = None, targets: torch.Tensor = None, k=10,
              return_aggregated: bool = True, return_individual: bool = False, flatten_results: bool = False,
              flattened_results_prefix: str = None, n_items: int = None, best_logit_indices: torch.Tensor = None,
              return_best_logit_indices: bool = False):
    ks = [k] if isinstance(k, int) else list(k)
    kmax = max(ks)
    metrics = list(metrics)
    if best_logit_indices is None:
        best_logit_indices = torch.topk(logits, kmax, dim=-1, largest=True, sorted=True).indices
    res = {}
    user_m = [m for m in metrics if m in supported_user_metrics]
    if user_m:
        tgt = targets if isinstance(targets, torch.Tensor) else torch.as_tensor(np.asarray(targets))
        rel = torch.gather(tgt.float(), 1, best_logit_indices.to(tgt.device))
        nt = tgt.float().sum(-1)
        for m in user_m:
            for kk in ks:
                v = _user_metric(m, rel, nt, kk).float()
                if return_aggregated:
                    res[_key(flattened_results_prefix, m, kk)] = v.mean().item()
                if return_individual:
                    res[_key(flattened_results_prefix, m, kk, True)] = v
    for m in metrics:
        if m == "coverage":
            for kk in ks:
                uniq = torch.unique(best_logit_indices[:, :kk]).numel()
                res[_key(flattened_results_prefix, m, kk)] = float(uniq) / float(n_items)
    if return_best_logit_indices:
        return res, best_logit_indices
    return res


def calculate_for_feature(group: UserFeature, metrics, logits, targets, k=10, return_individual=False,
                          flatten_results=True, flattened_results_prefix=None, **kw):
    res = {}
    labels = group.labels
    for lbl in np.unique(labels):
        sel = torch.as_tensor(labels == lbl)
        pre = f"{flattened_results_prefix}/" if flattened_results_prefix else ""
        r = calculate(metrics, logits[sel], targets[sel], k=k, return_aggregated=True,
                      return_individual=return_individual, flatten_results=True,
                      flattened_results_prefix=f"{pre}{group.name}_{lbl}")
        res.update(r)
    return res
