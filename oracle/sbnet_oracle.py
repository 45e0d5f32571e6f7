"""TEST INFRASTRUCTURE ONLY -- numpy (float64) restatement of the reference's SBNet hot path.

Never imported by the product package; used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` as the checker / CPU arm.  Pinned against the real reference through the
golden fixtures in ``tests/golden/`` (made by ``oracle/make_golden.py`` from the unmodified reference; checked by
``tests/test_oracle_vs_golden.py``).  The metric arithmetic (third-party ``rmet``) is PARITY UNPINNED, see
``oracle/rmet_restated.py``.

What each function follows (paths relative to the reference root):
  * ``FeatureProj``          -- ``algorithms/sgd_alg.py:1279-1396`` (FeatureEmbedding) + ``data/Feature.py:140-162``
  * ``poly_forward/backward``-- ``modules/polylinear.py:17-77`` (+ ``nn.BatchNorm1d`` train/eval semantics)
  * ``Entity``               -- ``algorithms/sgd_alg.py:1764-2006`` (SingleBranchNetEntity)
  * ``info_nce``             -- ``train/regularization_losses.py:8-43``
  * ``rec_loss``             -- ``train/rec_losses.py:40-113``
  * ``OracleSBNet``          -- ``algorithms/sgd_alg.py:2009-2144`` + step body ``train/trainer.py:204-223``
  * ``adam_step``            -- ``torch.optim.Adam`` / ``AdamW`` as selected at ``train/trainer.py:62-68``
  * ``masked_topk`` / ``metrics_at_k`` -- ``eval/eval.py:205-222`` + ``eval/metrics.py:4-105`` (+ restated rmet)
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

F64 = np.float64
BN_EPS, BN_MOMENTUM, NORM_EPS = 1e-5, 0.1, 1e-12


# ------------------------------------------------------------------------------------------------ activations
def act_fwd(name, x):
    if name is None:
        return x
    if name == "relu":
        return np.maximum(x, 0.)
    if name == "tanh":
        return np.tanh(x)
    if name == "sigmoid":
        return 1. / (1. + np.exp(-x))
    if name == "selu":
        a, s = 1.6732632423543772, 1.0507009873554805
        return s * np.where(x > 0, x, a * (np.exp(x) - 1.))
    raise ValueError(name)


def act_bwd(name, y, dy):
    """derivative expressed through the OUTPUT y (what the kernels save)."""
    if name is None:
        return dy
    if name == "relu":
        return dy * (y > 0)
    if name == "tanh":
        return dy * (1. - y * y)
    if name == "sigmoid":
        return dy * y * (1. - y)
    if name == "selu":
        a, s = 1.6732632423543772, 1.0507009873554805
        return dy * np.where(y > 0, s, y + s * a)
    raise ValueError(name)


# ------------------------------------------------------------------------------------------------ bf16 emulation
def bf16_round(x):
    """round-to-nearest-even to bfloat16, returned as float64 (what the kernels' bf16 stores do)"""
    x32 = np.ascontiguousarray(np.asarray(x, np.float32))
    u = x32.view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32).astype(F64)


class Bf16Emulation:
    """Mirrors the ROUNDING POINTS of the B200 path ('precision'): GEMM operands (weights, activations
    between fused stages, dz) are bf16; accumulation, BatchNorm, normalisation, losses and every reduction are
    fp32/fp64; the CSR ('sparse interactions') route gathers bf16 weight rows / bf16 dz rows with fp32 accumulation, i.e.
    it has the rounding points of the dense route (multi-hot inputs are exact in bf16).  With it the oracle predicts the
    kernels' results to ~1e-3 instead of the ~1e-2 that separates bf16 from the fp32 reference."""

    def __init__(self, csr_route=()):
        self.csr_route = set()  # (kept for callers; no module runs an fp32-exact first layer any more)

    q = staticmethod(bf16_round)


# ------------------------------------------------------------------------------------------------ PolyLinear
def poly_spec(layer_config, bn_every, activation, output_fn):
    """List of ops exactly in the order of ``modules/polylinear.py:50-71``."""
    ops = []
    L = len(layer_config) - 1
    for i in range(L):
        ops.append(("linear", f"linear_{i}"))
        if bn_every > 0 and (i + 1) % bn_every == 0:
            ops.append(("bn", f"batch_norm_{i}"))
        if i < L - 1:
            ops.append(("act", activation))
    if bn_every == -1:
        ops.append(("bn", "batch_norm"))
    if output_fn is not None:
        ops.append(("act", output_fn))
    return ops


def bn_forward(x, p, prefix, training, new_stats):
    g, b = p[prefix + ".weight"].astype(F64), p[prefix + ".bias"].astype(F64)
    if training:
        n = x.shape[0]
        mean = x.mean(0)
        var = x.var(0)  # biased
        xhat = (x - mean) / np.sqrt(var + BN_EPS)
        if new_stats is not None:
            rm, rv = p[prefix + ".running_mean"].astype(F64), p[prefix + ".running_var"].astype(F64)
            new_stats[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean
            new_stats[prefix + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * n / max(n - 1, 1)
            new_stats[prefix + ".num_batches_tracked"] = p[prefix + ".num_batches_tracked"] + 1
        return xhat * g + b, (xhat, np.sqrt(var + BN_EPS), g)
    rm, rv = p[prefix + ".running_mean"].astype(F64), p[prefix + ".running_var"].astype(F64)
    return (x - rm) / np.sqrt(rv + BN_EPS) * g + b, None


def bn_backward(dy, cache, grads, prefix):
    xhat, std, g = cache
    grads[prefix + ".weight"] = grads.get(prefix + ".weight", 0) + (dy * xhat).sum(0)
    grads[prefix + ".bias"] = grads.get(prefix + ".bias", 0) + dy.sum(0)
    return (g / std) * (dy - dy.mean(0) - xhat * (dy * xhat).mean(0))


def poly_forward(x, p, prefix, ops, training, new_stats=None, emu=None, exact_first=False):
    """emu: Bf16Emulation or None.  With emulation the input of every Linear and its weight are bf16-rounded (the
    kernels store inter-stage activations as bf16); ``exact_first`` keeps the first Linear in fp32 (CSR route)."""
    caches = []
    n_lin = 0
    for kind, arg in ops:
        if kind == "linear":
            W, b = p[f"{prefix}.{arg}.weight"].astype(F64), p[f"{prefix}.{arg}.bias"].astype(F64)
            if emu is not None and not (exact_first and n_lin == 0):
                x, W = emu.q(x), emu.q(W)
            n_lin += 1
            caches.append(x)
            x = x @ W.T + b
        elif kind == "bn":
            x, c = bn_forward(x, p, f"{prefix}.{arg}", training, new_stats)
            caches.append(c)
        else:
            x = act_fwd(arg, x)
            caches.append(x)
    return x, caches


def poly_backward(dy, p, prefix, ops, caches, grads, need_dx=True, emu=None, exact_first=False):
    n_lin = sum(1 for k, _ in ops if k == "linear")
    for (kind, arg), c in zip(reversed(ops), reversed(caches)):
        if kind == "linear":
            n_lin -= 1
            W = p[f"{prefix}.{arg}.weight"].astype(F64)
            kw, kb = f"{prefix}.{arg}.weight", f"{prefix}.{arg}.bias"
            grads[kb] = grads.get(kb, 0) + dy.sum(0)  # bias gradient: fp32 column sums before the bf16 store
            if emu is not None and not (exact_first and n_lin == 0):
                dy, W = emu.q(dy), emu.q(W)
            grads[kw] = grads.get(kw, 0) + dy.T @ c
            dy = dy @ W
        elif kind == "bn":
            dy = bn_backward(dy, c, grads, f"{prefix}.{arg}")
        else:
            dy = act_bwd(arg, c, dy)
    return dy


# ------------------------------------------------------------------------------------------------ features
def _ftype(feature):
    return str(getattr(feature.feature_definition.type, "value", feature.feature_definition.type)).lower()


class FeatureProj:
    """One ``FeatureEmbedding`` (sgd_alg.py:1279-1396): raw feature rows -> [rows, out_dim]."""

    def __init__(self, feature, prefix, embedding_dim, pre_layers, activation, post_layers=None):
        self.f, self.prefix, self.act = feature, prefix, activation
        self.type = _ftype(feature)
        self.remap = {int(e): r for r, e in enumerate(np.asarray(feature._indices).tolist())}
        self.pre_ops = self.post_ops = None
        self.out_dim = embedding_dim
        if self.type not in ("categorical", "tag"):
            dim = feature.dim if not isinstance(feature.dim, tuple) else int(np.prod(feature.dim))
            cfg = [int(dim)] + list(pre_layers or []) + ([embedding_dim] if embedding_dim is not None else [])
            self.out_dim = cfg[-1]
            if len(cfg) > 1:
                self.pre_ops = poly_spec(cfg, 0, activation, activation)
        if post_layers:
            cfg = [self.out_dim] + list(post_layers)
            self.out_dim = cfg[-1]
            self.post_ops = poly_spec(cfg, 0, activation, activation)

    def rows(self, ent_idx):
        return np.array([self.remap[int(e)] for e in np.asarray(ent_idx).reshape(-1)], dtype=np.int64)

    def raw(self, rows):
        v = self.f.values[rows]
        if sp.issparse(v):
            v = v.toarray()
        v = np.asarray(v)
        if self.type in ("continuous", "discrete"):
            v = v.reshape(len(rows), 1)
        return v

    def forward(self, ent_idx, p, emu=None):
        """``emu`` (a plain, non single-branch entity on the B200 path): the table of ALL feature rows is computed once
        with the kernels' rounding points and the batch gathers rows of it; the backward sums the row gradients per
        feature row before they enter the chain (same result in exact arithmetic)."""
        rows = self.rows(ent_idx)
        if emu is None or (self.type == "categorical" and self.post_ops is None):
            return self._forward_rows(rows, p, None)
        n = self.f.values.shape[0]
        T, cache = self._forward_rows(np.arange(n), p, emu)
        return T[rows], {"rows": rows, "table": cache, "n": n, "emu": emu}

    def _forward_rows(self, rows, p, emu):
        x = self.raw(rows)
        cache = {"rows": rows}
        if self.type == "categorical":
            E = p[self.prefix + ".embedding_layer.weight"].astype(F64)
            cache["cat"] = x.astype(np.int64)
            y = E[cache["cat"]]
        elif self.type == "tag":
            E = p[self.prefix + ".embedding_layer.weight"].astype(F64)
            pad = E.shape[0] - 1  # padding_idx=-1
            tags = x.astype(np.int64)
            valid = tags != pad
            cnt = valid.sum(1)
            y = (E[tags] * valid[..., None]).sum(1) / np.maximum(cnt, 1)[:, None]
            cache.update(tags=tags, valid=valid, cnt=cnt)
        else:
            y = x.astype(F64)
            if self.pre_ops is not None:
                y, cache["pre"] = poly_forward(y, p, self.prefix + ".pre_embedding_layers.layers", self.pre_ops, True,
                                               emu=emu)
        if self.post_ops is not None:
            y, cache["post"] = poly_forward(y, p, self.prefix + ".post_embedding_layers.layers", self.post_ops, True,
                                            emu=emu)
        return y, cache

    # -- table mode (what the B200 path does): project ALL rows once, gather afterwards; the backward sums the
    #    row gradients per feature row BEFORE the activation derivative / bf16 rounding of the wgrad operand
    def table_forward(self, p, emu):
        n = self.f.values.shape[0]
        x = self.raw(np.arange(n)).astype(F64)
        csr = emu is not None and self.prefix in emu.csr_route
        y, cache = poly_forward(x, p, self.prefix + ".pre_embedding_layers.layers", self.pre_ops, True, emu=emu,
                                exact_first=csr)
        self._table_cache = (cache, csr)
        return y

    def table_backward(self, dT, p, grads, emu):
        cache, csr = self._table_cache
        poly_backward(dT, p, self.prefix + ".pre_embedding_layers.layers", self.pre_ops, cache, grads, emu=emu,
                      exact_first=csr)

    def backward(self, dy, cache, p, grads):
        emu = None
        if "table" in cache:  # table mode: row gradients summed per feature row first
            dT = np.zeros((cache["n"], dy.shape[-1]), F64)
            np.add.at(dT, cache["rows"], dy)
            dy, emu, cache = dT, cache["emu"], cache["table"]
        if self.post_ops is not None:
            dy = poly_backward(dy, p, self.prefix + ".post_embedding_layers.layers", self.post_ops, cache["post"],
                               grads, emu=emu)
        k = self.prefix + ".embedding_layer.weight"
        if self.type == "categorical":
            g = grads.get(k)
            if g is None or np.isscalar(g):
                g = np.zeros(p[k].shape, F64)
            np.add.at(g, cache["cat"], dy)
            grads[k] = g
        elif self.type == "tag":
            g = grads.get(k)
            if g is None or np.isscalar(g):
                g = np.zeros(p[k].shape, F64)
            w = cache["valid"] / np.maximum(cache["cnt"], 1)[:, None]
            np.add.at(g, cache["tags"], dy[:, None, :] * w[..., None])
            g[-1] = 0.  # padding row never receives gradient
            grads[k] = g
        elif self.pre_ops is not None:
            poly_backward(dy, p, self.prefix + ".pre_embedding_layers.layers", self.pre_ops, cache["pre"], grads,
                          emu=emu)


# ------------------------------------------------------------------------------------------------ losses
def _log_softmax(z, axis):
    m = z.max(axis=axis, keepdims=True)
    return z - m - np.log(np.exp(z - m).sum(axis=axis, keepdims=True))


def info_nce(e0, e1, temperature):
    """returns loss, d e0, d e1.  e*: [..., n, D]; contrast along the second-last axis; mean over all rows."""
    logits = e0 @ np.swapaxes(e1, -1, -2) / temperature
    n = logits.shape[-1]
    R = logits.size // n
    ls_r = _log_softmax(logits, -1)
    ls_c = _log_softmax(logits, -2)
    eye = np.eye(n)
    loss = -(ls_r * eye).sum() / R - (ls_c * eye).sum() / R
    dL = (np.exp(ls_r) - eye) / R + (np.exp(ls_c) - eye) / R
    de0 = dL @ e1 / temperature
    de1 = np.swapaxes(dL, -1, -2) @ e0 / temperature
    return loss, de0, de1


def _softplus(x):
    return np.logaddexp(0., x)


def _sigmoid(x):
    return 1. / (1. + np.exp(-x))


def rec_loss(kind, logits, aggregator="mean", n_items=None, neg_train=None, neg_strategy="uniform_recbole"):
    """returns loss (float64), d logits.  Column 0 is the positive (train/rec_losses.py)."""
    B, n = logits.shape
    z = logits.astype(F64)
    if kind == "bpr":
        diff = z[:, :1] - z[:, 1:]
        cnt = diff.size if aggregator == "mean" else 1
        loss = _softplus(-diff).sum() / cnt
        g = -_sigmoid(-diff) / cnt
        d = np.concatenate([g.sum(1, keepdims=True), -g], axis=1)
        return loss, d
    if kind == "bce":
        y = np.zeros_like(z)
        y[:, 0] = 1.
        cnt = z.size if aggregator == "mean" else 1
        loss = (_softplus(z) - y * z).sum() / cnt
        return loss, (_sigmoid(z) - y) / cnt
    if kind == "sampled_softmax":
        z = z.copy()
        if neg_strategy == "uniform":
            z[:, 1:] += math.log(n_items / neg_train)
        ls = _log_softmax(z, -1)
        cnt = B if aggregator == "mean" else 1
        loss = -ls[:, 0].sum() / cnt
        d = np.exp(ls)
        d[:, 0] -= 1.
        return loss, d / cnt
    raise ValueError(kind)


# ------------------------------------------------------------------------------------------------ entity
class Entity:
    """SingleBranchNetEntity (sgd_alg.py:1764-2006)."""

    def __init__(self, name, features, conf, D, val_interactions_available):
        self.name, self.conf, self.D = name, conf, D
        self.prefix = f"{name}_embedding_module"
        feats = [f["feature_name"] for f in conf["features"]]
        avail = set(feats)
        self.train_mods = set(conf.get("train_modalities") or avail)
        ev = set(conf.get("eval_modalities") or self.train_mods)
        if not val_interactions_available:
            ev.discard("interactions")
        self.eval_mods = ev
        self.C = conf["common_modality_dim"]
        self.act = conf.get("activation_fn", "relu")
        self.proj = {}
        for f in conf["features"]:
            fn = f["feature_name"]
            if fn not in self.train_mods:
                continue
            self.proj[fn] = FeatureProj(features[fn], f"{self.prefix}.modality_modules.{fn}", self.C,
                                        f.get("feature_hidden_layers"), self.act)
        self.p_drop = conf.get("single_branch_input_dropout")
        s = 1 if self.p_drop is not None else 0
        use_bn = conf.get("apply_batch_normalization", True)
        every = conf.get("apply_batch_norm_every", 0) if use_bn else 0
        H = [self.C] + list(conf["single_branch_hidden_layers"]) + [D]
        self.sb_prefix = f"{self.prefix}.sb_net.{s}.layers"
        self.sb_ops = poly_spec(H, every, self.act, self.act if conf.get("apply_output_activation", False) else None)
        self.trailing_bn = f"{self.prefix}.sb_net.{s + 1}" if (use_bn and conf.get("apply_batch_norm_every", 0) == 0) \
            else None
        self.normalize = conf.get("normalize_single_branch_input", False)
        self.agg = conf.get("aggregation_fn", "mean")
        self.reg_type = conf.get("embedding_regularization_type", "no_regularization")
        self.temperature = conf.get("regularization_temperature", 1.)
        self.reg_weight = conf.get("regularization_weight", 1.)

    def forward(self, idx, mods, mod_names, p, training, drop_keep=None, new_stats=None, emu=None):
        """idx: int array [...]; mods: int ids [..., k] into mod_names."""
        shape = idx.shape
        k = mods.shape[-1]
        flat_idx = np.repeat(idx.reshape(-1), k)
        flat_mod = mods.reshape(-1)
        N = flat_idx.size
        X = np.zeros((N, self.C), F64)
        c = {"shape": shape, "k": k, "proj": {}, "emu": emu}
        for mid in np.unique(flat_mod):
            sel = np.nonzero(flat_mod == mid)[0]
            fp = self.proj[str(mod_names[mid])]
            if emu is not None and fp.pre_ops is not None:
                rows = fp.rows(flat_idx[sel])
                X[sel] = fp.table_forward(p, emu)[rows]
                c["proj"][int(mid)] = (sel, {"table_rows": rows})
                continue
            y, pc = fp.forward(flat_idx[sel], p)
            X[sel] = y
            c["proj"][int(mid)] = (sel, pc)
        if self.normalize:
            nrm = np.maximum(np.linalg.norm(X, axis=1, keepdims=True), NORM_EPS)
            c["norm"] = (X / nrm, nrm)
            X = X / nrm
        if self.p_drop is not None and training and self.p_drop > 0:
            keep = np.ones_like(X) if drop_keep is None else drop_keep.reshape(N, self.C).astype(F64)
            scale = keep / (1. - self.p_drop)
            c["drop"] = scale
            X = X * scale
        Z, c["sb"] = poly_forward(X, p, self.sb_prefix, self.sb_ops, training, new_stats, emu=emu)
        if self.trailing_bn:
            Z, c["tbn"] = bn_forward(Z, p, self.trailing_bn, training, new_stats)
        E = Z.reshape(shape + (k, self.D))
        c["E"] = E
        reg = 0.
        c["dE_reg"] = None
        if training and self.reg_type != "no_regularization":
            assert k == 2
            loss, d0, d1 = info_nce(E[..., 0, :], E[..., 1, :], self.temperature)
            reg = loss
            c["dE_reg"] = np.stack([d0, d1], axis=-2)
        out = E.mean(-2) if self.agg == "mean" else E.max(-2)
        self.cache = c
        return out, reg

    def backward(self, d_out, p, grads, mod_names, reg_scale=1.):
        c = self.cache
        E, k = c["E"], c["k"]
        if self.agg == "mean":
            dE = np.repeat(d_out[..., None, :], k, axis=-2) / k
        else:
            arg = E.argmax(-2)
            dE = np.zeros_like(E)
            np.put_along_axis(dE, arg[..., None, :], d_out[..., None, :], axis=-2)
        if c["dE_reg"] is not None:
            dE = dE + c["dE_reg"] * (self.reg_weight * reg_scale)
        dZ = dE.reshape(-1, self.D)
        if self.trailing_bn:
            dZ = bn_backward(dZ, c["tbn"], grads, self.trailing_bn)
        emu = c.get("emu")
        dX = poly_backward(dZ, p, self.sb_prefix, self.sb_ops, c["sb"], grads, emu=emu)
        if "drop" in c:
            dX = dX * c["drop"]
        if self.normalize:
            y, nrm = c["norm"]
            dX = (dX - y * (y * dX).sum(1, keepdims=True)) / nrm
        for mid, (sel, pc) in c["proj"].items():
            fp = self.proj[str(mod_names[mid])]
            if "table_rows" in pc:
                dT = np.zeros((fp.f.values.shape[0], self.C), F64)
                np.add.at(dT, pc["table_rows"], dX[sel])
                fp.table_backward(dT, p, grads, emu)
            else:
                fp.backward(dX[sel], pc, p, grads)


class OracleSBNet:
    """SingleBranchNet (sgd_alg.py:2009-2144) + one trainer step (train/trainer.py:204-223)."""

    def __init__(self, conf: dict, dataset):
        from types import SimpleNamespace as NS
        self.D = conf["shared_common_dim"]
        uf, itf = dict(dataset.user_features), dict(dataset.item_features)

        def synth(name, ftype, values, n):
            return NS(feature_definition=NS(name=name, type=ftype, tag_split_sep=None), values=values,
                      dim=(values.shape[1] if ftype == "vector" else 0), _indices=np.arange(n))
        uf["interactions"] = synth("interactions", "vector", dataset.user_sampling_matrix_train, dataset.n_users)
        uf["user_embedding"] = synth("user_embedding", "categorical", np.arange(dataset.n_users), dataset.n_users)
        itf["interactions"] = synth("interactions", "vector", dataset.item_sampling_matrix_train, dataset.n_items)
        itf["item_embedding"] = synth("item_embedding", "categorical", np.arange(dataset.n_items), dataset.n_items)
        self.ent = {}
        for name, feats, cold in (("user", uf, dataset.is_cold_start_user), ("item", itf, dataset.is_cold_start_item)):
            c = conf[name]
            if "features" in c:
                self.ent[name] = Entity(name, feats, c, self.D, not cold)
            else:
                dim = self.D if c["embedding_dim"] == -1 else c["embedding_dim"]
                self.ent[name] = FeatureProj(feats[c["feature_name"]], f"{name}_embedding_module", dim,
                                             c.get("pre_embedding_layers"), c.get("activation_fn", "relu"),
                                             c.get("post_embedding_layers"))

    # -- representations
    def represent(self, name, idx, p, training, mods=None, mod_names=None, drop_keep=None, new_stats=None, emu=None):
        e = self.ent[name]
        if isinstance(e, FeatureProj):
            y, cache = e.forward(idx.reshape(-1), p, emu)
            self._plain_cache = getattr(self, "_plain_cache", {})
            self._plain_cache[name] = cache
            return y.reshape(idx.shape + (-1,)), 0.
        if not training:
            mod_names = sorted(e.eval_mods)
            mods = np.broadcast_to(np.arange(len(mod_names)), idx.shape + (len(mod_names),))
        return e.forward(idx, np.asarray(mods), mod_names, p, training, drop_keep, new_stats, emu=emu)

    def train_step_fwd_bwd(self, p, u, i, mods, mod_names, drop, loss_kind="bpr", aggregator="mean", n_items=None,
                           neg_train=None, neg_strategy="uniform_recbole", emu=None):
        """mods/mod_names/drop: dicts keyed 'user'/'item'.  Returns dict(logits, rec_loss, reg losses, grads,
        new_stats)."""
        new_stats, grads = {}, {}
        ur, ureg = self.represent("user", u, p, True, mods.get("user"), mod_names.get("user"), drop.get("user"),
                                  new_stats, emu)
        ir, ireg = self.represent("item", i, p, True, mods.get("item"), mod_names.get("item"), drop.get("item"),
                                  new_stats, emu)
        logits = np.einsum("be,bce->bc", ur, ir)
        rl, dlog = rec_loss(loss_kind, logits, aggregator, n_items, neg_train, neg_strategy)
        d_ur = np.einsum("bc,bce->be", dlog, ir)
        d_ir = dlog[..., None] * ur[:, None, :]
        res = dict(logits=logits, rec_loss=rl, new_stats=new_stats)
        reg_total = 0.
        for name, d_out, reg, idx in (("user", d_ur, ureg, u), ("item", d_ir, ireg, i)):
            e = self.ent[name]
            if isinstance(e, FeatureProj):
                e.backward(d_out.reshape(-1, d_out.shape[-1]), self._plain_cache[name], p, grads)
                continue
            res[f"{name}_reg_loss"] = reg * e.reg_weight
            reg_total += reg * e.reg_weight
            e.backward(d_out, p, grads, mod_names[name])
        res["reg_loss"] = reg_total
        res["loss"] = rl + reg_total
        res["grads"] = {k: (np.zeros(p[k].shape, F64) if np.isscalar(v) else v) for k, v in grads.items()}
        return res


def adam_step(p, grads, state, lr, wd, step, decoupled=True, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.AdamW (decoupled=True) / Adam (L2 folded into the gradient).  In place on ``p`` / ``state``."""
    b1, b2 = betas
    for k, g in grads.items():
        w = p[k].astype(F64)
        g = g.astype(F64)
        m = state.setdefault("m/" + k, np.zeros_like(w))
        v = state.setdefault("v/" + k, np.zeros_like(w))
        if decoupled:
            w = w * (1. - lr * wd)
        else:
            g = g + wd * w
        m[:] = b1 * m + (1 - b1) * g
        v[:] = b2 * v + (1 - b2) * g * g
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        w = w - (lr / bc1) * m / (np.sqrt(v) / math.sqrt(bc2) + eps)
        p[k] = w


def adagrad_step(p, grads, state, lr, wd, eps=1e-10):
    """torch.optim.Adagrad with its defaults (lr_decay 0, initial accumulator 0), as built by ``train/trainer.py:62-68``:
    coupled weight decay, ``sum += g^2``, ``w -= lr * g / (sqrt(sum) + eps)``.  In place on ``p`` / ``state``."""
    for k, g in grads.items():
        w = p[k].astype(F64)
        g = g.astype(F64) + wd * w
        acc = state.setdefault("sum/" + k, np.zeros_like(w))
        acc[:] = acc + g * g
        p[k] = w - lr * g / (np.sqrt(acc) + eps)


# ------------------------------------------------------------------------------------------------ evaluation
def masked_topk(u_repr, i_repr, exclude_csr, k):
    """scores = u @ i.T, seen -> -inf (eval/eval.py:217-220), top-k with ties broken by LOWEST index.
    Returns (values [U,k] f64, indices [U,k] int64 into items_in_split)."""
    scores = u_repr.astype(F64) @ i_repr.astype(F64).T
    if exclude_csr is not None:
        ex = exclude_csr.tocsr()
        for r in range(scores.shape[0]):
            scores[r, ex.indices[ex.indptr[r]:ex.indptr[r + 1]]] = -np.inf
    order = np.lexsort((np.broadcast_to(np.arange(scores.shape[1]), scores.shape), -scores), axis=-1)[:, :k]
    return np.take_along_axis(scores, order, 1), order


def metrics_at_k(topk_idx, target_csr, ks, n_items=None):
    """per-user metric vectors keyed '{metric}@{k}' (float64) from ranked indices and a target CSR (rows aligned
    with ``topk_idx`` rows).  Definitions: eval/metrics.py:4-105 + oracle/rmet_restated.py."""
    tgt = target_csr.tocsr()
    U, kmax = topk_idx.shape
    rel = np.zeros((U, kmax), F64)
    nt = np.diff(tgt.indptr).astype(F64)
    for r in range(U):
        rel[r] = np.isin(topk_idx[r], tgt.indices[tgt.indptr[r]:tgt.indptr[r + 1]])
    out = {}
    for k in ks:
        hits = rel[:, :k].sum(1)
        disc = 1. / np.log2(np.arange(2, k + 2, dtype=np.float32)).astype(F64)
        dcg = (rel[:, :k] * disc).sum(1)
        idcg = np.array([disc[:int(min(n, k))].sum() for n in nt])
        with np.errstate(divide="ignore", invalid="ignore"):
            prec = hits / k
            rec = np.where(nt > 0, hits / nt, 0.)
            out[f"ndcg@{k}"] = np.minimum(np.where(idcg > 0, dcg / idcg, 0.), 1.)
            out[f"precision@{k}"] = prec
            out[f"recall@{k}"] = rec
            out[f"f_score@{k}"] = np.where(prec + rec > 0, 2 * prec * rec / (prec + rec), 0.)
            out[f"hitrate@{k}"] = np.minimum(hits, 1.)
            # rmet (restated, oracle/rmet_restated.py): ap = sum_r rel_r * precision@r / min(k, n_targets),
            # rr = 1 / rank of the first hit
            prec_at = np.cumsum(rel[:, :k], 1) / np.arange(1, k + 1, dtype=F64)
            denom = np.minimum(nt, k)
            out[f"ap@{k}"] = np.where(denom > 0, (prec_at * rel[:, :k]).sum(1) / np.maximum(denom, 1.), 0.)
            first = np.where(rel[:, :k].sum(1) > 0, rel[:, :k].argmax(1) + 1, 0)
            out[f"rr@{k}"] = np.where(first > 0, 1. / np.maximum(first, 1), 0.)
        if n_items:
            out[f"coverage@{k}"] = len(np.unique(topk_idx[:, :k])) / float(n_items)
    return out
