"""Full-catalog evaluation on the device: replaces ``evaluate_recommender_algorithm`` + ``FullEvaluator``
(``eval/eval.py:20-227``) for SGD-based algorithms.

Item representations once (all ``items_in_split``, all eval modalities), user representations once, then ONE fused
kernel does scores -> seen-item mask -> exact top-k (``csrc/eval_topk.cu``) followed by the metric kernel; the
[U, I] score matrix, the dense label rows (``data/dataset.py:445-453``) and the dense host masks
(``eval/eval.py:219``) of the reference never exist.  Result keys follow the reference: ``'{metric}@{k}'``
(optionally ``'{name}/...'``), mean over the users of the split, ``_std`` when requested, ``coverage@k`` from all
users' top-k.
"""
from __future__ import annotations

import re
from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from .config import EvalConfig
from .feature_store import csr_to_device

USER_METRICS = ["ndcg", "precision", "recall", "f_score", "hitrate", "ap", "rr"]  # order of sbr_metrics_at_k's output
SUPPORTED = USER_METRICS + ["coverage"]


def _natural_key(s: str):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)]


def _user_labels(feat, users) -> np.ndarray:
    """group label of every user for a categorical user feature (``Feature.__getitem__`` + ``get_labels``,
    data/Feature.py:126-128, 140-162), lower-cased like eval/eval.py:111-112"""
    if hasattr(type(feat), "__getitem__"):
        values = np.asarray(feat[users]).reshape(-1)
    else:
        row_of = {int(e): r for r, e in enumerate(np.asarray(feat._indices).tolist())}
        values = np.asarray(feat.values)[[row_of[int(u)] for u in users]].reshape(-1)
    if hasattr(feat, "get_labels"):
        labels = np.asarray(feat.get_labels(values)).tolist()
    elif getattr(feat, "_unique_values", None) is not None:
        labels = [feat._unique_values[int(v)] for v in values]
    else:
        labels = values.tolist()
    return np.array([lbl.lower() if isinstance(lbl, str) else lbl for lbl in labels])


class FullEvaluator:
    def __init__(self, config, evaluator_name: Optional[str] = None, dataset=None, cuda_graph: bool = False):
        if isinstance(config, dict):
            config = EvalConfig.from_dict(config)
        invalid = set(config.metrics) - set(SUPPORTED)
        if invalid:
            raise ValueError(f"Metric(s) {invalid} are not supported. Select metrics from {SUPPORTED}.")
        self.config, self.name, self.dataset = config, evaluator_name, dataset
        self._dev_cache = None
        # cuda_graph: the whole evaluation (representations, scores + mask + top-k, metrics, their reductions) of one
        # (model, split) is captured on its second call and replayed afterwards; one D2H copy per evaluation
        self.cuda_graph = bool(cuda_graph)
        self._graphs = {}

    def _device_split(self, dataset, device):
        if self._dev_cache is not None and self._dev_cache[0] is dataset and self._dev_cache[1] == device:
            return self._dev_cache[2]
        users = np.asarray(dataset.users_in_split)
        items = np.asarray(dataset.items_in_split)
        seen = dataset.exclude_data[users] if dataset.exclude_data is not None else None
        tgt = dataset.user_sampling_matrix[users][:, items]
        pack = dict(users=torch.from_numpy(users.astype(np.int64)).to(device),
                    items=torch.from_numpy(items.astype(np.int64)).to(device),
                    seen=csr_to_device(seen, device) if seen is not None and seen.nnz > 0 else (None, None),
                    tgt=csr_to_device(tgt, device))
        self._dev_cache = (dataset, device, pack)
        return pack

    @torch.no_grad()
    def evaluate(self, model, dataset=None, return_topk: bool = False) -> Dict[str, float]:
        dataset = dataset or self.dataset
        dev = model.device
        d = self._device_split(dataset, dev)
        n_items = len(dataset.items_in_split)
        ks = sorted(set(int(k) for k in self.config.top_k))
        st = self._graphs.setdefault((id(model), id(dataset)), {"calls": 0}) if self.cuda_graph else None

        def device_part():
            was_training = model.training
            model.eval()
            if hasattr(model, "eval_factors"):
                # models whose score is not a plain dot product of the two representations (bias terms, slot means:
                # sibling.py) hand over factor matrices whose dot product IS the score
                u_repr, i_repr = model.eval_factors(d["users"], d["items"])
            else:
                i_repr = model.get_item_representations(d["items"])
                u_repr = model.get_user_representations(d["users"])
            if was_training:
                model.train()
            return self._device_eval(u_repr, i_repr, d["seen"], d["tgt"], ks, n_items)

        if st is None or st["calls"] < 1:
            packed, topk = device_part()
            if st is not None:
                st["calls"] += 1
        else:
            model.refresh_shadows()
            if "graph" not in st:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    st["out"] = device_part()
                st["graph"] = g
            st["graph"].replay()
            packed, topk = st["out"]
        out = self._host_results(packed, ks, n_items)
        out.update(self.group_metrics(dataset, ks))
        out = {k: out[k] for k in sorted(out, key=_natural_key)}
        return (out, topk) if return_topk else out

    def _device_eval(self, u_repr, i_repr, seen, tgt, ks, n_items, item_offset=0):
        """everything that runs on the device: returns (packed [mean | std | coverage counts] tensor, (vals, idx))"""
        U, D = u_repr.shape
        I = i_repr.shape[0]
        kmax = min(max(ks), I)
        u16 = ops.cast_bf16(u_repr.contiguous())
        i16 = ops.cast_bf16(i_repr.contiguous())
        vals, idx = ops.topk_scores_masked(u16, i16, U, I, u16.shape[1], seen[0], seen[1], kmax,
                                           item_offset=item_offset)
        return self._device_metrics(idx, tgt, ks, n_items), (vals, idx)

    def _device_metrics(self, idx, tgt, ks, n_items):
        kmax = idx.shape[1]
        ks_eff = [k for k in ks if k <= kmax]
        want_cov = "coverage" in self.config.metrics
        m, hits = ops.metrics_at_k(idx, tgt[0], tgt[1], ks_eff, n_items, want_item_hits=want_cov)
        parts = [m.mean(dim=2).reshape(-1), m.std(dim=2, unbiased=False).reshape(-1)]
        parts.append(hits.sum(dim=1).to(torch.float32) if want_cov else torch.zeros(len(ks_eff), device=m.device))
        self.raw, self._kmax = m, kmax
        return torch.cat(parts)

    def _host_results(self, packed, ks, n_items) -> Dict[str, float]:
        ks_eff = [k for k in ks if k <= self._kmax]
        nk = len(ks_eff)
        host = packed.cpu().numpy()  # the one device -> host copy of an evaluation
        mean, std, cov = host[:7 * nk].reshape(7, nk), host[7 * nk:14 * nk].reshape(7, nk), host[14 * nk:]
        pre = f"{self.name}/" if self.name else ""
        res = {}
        for mi, name in enumerate(USER_METRICS):
            if name not in self.config.metrics:
                continue
            for ki, k in enumerate(ks_eff):
                res[f"{pre}{name}@{k}"] = float(mean[mi, ki])
                if self.config.calculate_std:
                    res[f"{pre}{name}@{k}_std"] = float(std[mi, ki])
        if "coverage" in self.config.metrics:
            for ki, k in enumerate(ks_eff):
                res[f"{pre}coverage@{k}"] = float(cov[ki]) / float(n_items)
        return {k: res[k] for k in sorted(res, key=_natural_key)}

    def group_metrics(self, dataset, ks) -> Dict[str, float]:
        """per-group means (+ std) of the user metrics for every categorical user feature in
        ``eval.user_group_features`` (``FullEvaluator._calculate_group_metrics``, eval/eval.py:106-119): keys
        ``'[{name}/]{feature}_{label}/{metric}@{k}'``.  Uses the per-user metric vectors of the last evaluation."""
        if not self.config.calculate_group_metrics or not self.config.user_group_features:
            return {}
        m = self.raw  # [7, n_ks, U]
        ks_eff = [k for k in ks if k <= self._kmax]
        users = np.asarray(dataset.users_in_split)
        pre = f"{self.name}/" if self.name else ""
        res = {}
        for fname in self.config.user_group_features:
            labels = _user_labels(dataset.user_features[fname], users)
            for lbl in np.unique(labels):
                sel = torch.from_numpy(np.nonzero(labels == lbl)[0]).to(m.device)
                sub = m.index_select(2, sel)
                mean = sub.mean(dim=2).cpu().numpy()
                std = sub.std(dim=2, unbiased=False).cpu().numpy() if self.config.calculate_std else None
                for mi, name in enumerate(USER_METRICS):
                    if name not in self.config.metrics:
                        continue
                    for ki, k in enumerate(ks_eff):
                        key = f"{pre}{fname}_{lbl}/{name}@{k}"
                        res[key] = float(mean[mi, ki])
                        if std is not None:
                            res[f"{key}_std"] = float(std[mi, ki])
        return res

    @torch.no_grad()
    def evaluate_representations(self, u_repr, i_repr, seen, tgt, n_items, return_topk=False, item_offset=0):
        ks = sorted(set(int(k) for k in self.config.top_k))
        packed, topk = self._device_eval(u_repr, i_repr, seen, tgt, ks, n_items, item_offset=item_offset)
        out = self._host_results(packed, ks, n_items)
        return (out, topk) if return_topk else out

    def metrics_from_topk(self, idx, tgt, ks, n_items) -> Dict[str, float]:
        return self._host_results(self._device_metrics(idx, tgt, ks, n_items), ks, n_items)


def evaluate_recommender_algorithm(alg, eval_loader_or_dataset, evaluator: FullEvaluator, device="cuda",
                                   return_raw=False, verbose=False):
    """signature-compatible entry (``eval/eval.py:171``); ``eval_loader_or_dataset`` may be a DataLoader whose
    ``.dataset`` is the eval split, or the split itself."""
    dataset = getattr(eval_loader_or_dataset, "dataset", eval_loader_or_dataset)
    results = evaluator.evaluate(alg, dataset)
    if return_raw:
        return results, _raw_metrics(evaluator)
    return results


def _raw_metrics(evaluator: FullEvaluator) -> Dict[str, np.ndarray]:
    return {f"{name}@{k}": evaluator.raw[mi, ki].cpu().numpy()
            for mi, name in enumerate(USER_METRICS)
            for ki, k in enumerate(sorted(set(evaluator.config.top_k)))
            if ki < evaluator.raw.shape[1]}


def gather_recommender_algorithm_results(alg, eval_loader_or_dataset, evaluator: FullEvaluator,
                                         results_path: Optional[str] = None, device="cuda", verbose=False):
    """``run_gather`` entry (``eval/eval.py:258-333``): the top-``max(top_k)`` logits and item positions of every user
    of the split, the user indices, the targets and the (raw) metrics in one dict, optionally pickled.

    Keys as in the reference: ``n_users, n_items, k, topk_item_indices [U, k]`` (column positions inside
    ``items_in_split``, best first), ``topk_logits [U, k]``, ``user_indices [U]``, ``targets [nnz, 2]``, ``metrics``,
    ``raw_metrics``.  Deviations: the logits are those of the bf16 scoring kernel; the first column of ``targets`` is
    the user's position in ``users_in_split`` (the reference concatenates per-batch ``argwhere`` rows, i.e. positions
    inside each evaluation batch); users with fewer than k unmasked items get ``(-inf, -1)`` in the tail."""
    import pickle
    dataset = getattr(eval_loader_or_dataset, "dataset", eval_loader_or_dataset)
    metrics, (vals, idx) = evaluator.evaluate(alg, dataset, return_topk=True)
    users = np.asarray(dataset.users_in_split)
    labels = dataset.user_sampling_matrix[users][:, np.asarray(dataset.items_in_split)].tocoo()
    order = np.lexsort((labels.col, labels.row))
    out = dict(n_users=dataset.n_users_in_split, n_items=dataset.n_items_in_split, k=int(idx.shape[1]),
               topk_item_indices=idx.cpu().numpy().astype(np.int64), topk_logits=vals.cpu().numpy(),
               user_indices=users.astype(np.int64),
               targets=np.stack([labels.row[order], labels.col[order]], axis=1).astype(np.int64),
               metrics=metrics, raw_metrics=_raw_metrics(evaluator))
    if results_path is not None:
        with open(results_path, "wb") as fh:
            pickle.dump(out, fh)
    return out
