"""Multi-threaded PyTorch-CPU restatement of the DLRM / two-tower forward — TEST INFRASTRUCTURE,
used only as the timed CPU baseline (`bench.py` cpu_baseline / `--impl reference`) and checked
against oracle/oracle.py (NumPy) in tests/test_oracle.py.

TensorFlow cannot be installed here (no network), so "the reference's own TF-CPU path" is stood
in for by the same op sequence TF would run, on the same host cores: F.embedding (= tf.gather),
torch.stack, torch.bmm + triu mask (= BatchMatMul + boolean_mask), F.linear + relu (= Dense).
File:line citations as in oracle/oracle.py.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


def _act(x: torch.Tensor, name: Optional[str]) -> torch.Tensor:
    if name in (None, "linear"):
        return x
    return {"relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh, "selu": F.selu, "elu": F.elu,
            "gelu": F.gelu}[name](x)


def mlp(x: torch.Tensor, layers: Sequence[dict]) -> torch.Tensor:
    """blocks/mlp.py:275-280 per layer."""
    for l in layers:
        x = torch.addmm(l["bias"], x, l["kernel"]) if l.get("bias") is not None else x @ l["kernel"]
        x = _act(x, l.get("activation"))
    return x


def dlrm_forward(idx: Dict[str, torch.Tensor], dense: Dict[str, torch.Tensor], tables: Dict[str, torch.Tensor],
                 feature_table: Dict[str, str], bottom: List[dict], top: List[dict], head: dict) -> torch.Tensor:
    """Same staging as the reference graph: T gathers -> bottom MLP -> stack (sorted names) ->
    bmm -> boolean mask -> concat [bottom | interactions] -> top MLP -> Dense(1, sigmoid)."""
    emb = {n: F.embedding(idx[n].long().reshape(-1), tables[t]) for n, t in feature_table.items()}
    x = torch.cat([dense[k].reshape(-1, 1).float() for k in sorted(dense)], dim=1)  # aggregation.py:54-66
    emb["bottom_block"] = mlp(x, bottom)
    stacked = torch.stack([emb[k] for k in sorted(emb)], dim=1)  # aggregation.py:101-108
    z = torch.bmm(stacked, stacked.transpose(1, 2))  # interaction.py:102
    Fn = stacked.shape[1]
    mask = torch.triu(torch.ones(Fn, Fn, dtype=torch.bool), diagonal=1)
    inter = z[:, mask]  # interaction.py:107-112
    body = mlp(torch.cat([emb["bottom_block"], inter], dim=1), top)
    return _act(torch.addmm(head["bias"], body, head["kernel"]), head.get("activation", "sigmoid"))


def tower_forward(idx, dense, bags, tables, feature_table, layers, combiner="mean") -> torch.Tensor:
    d = {}
    for n, t in feature_table.items():
        if n in bags:
            v, o = bags[n]
            d[n] = F.embedding_bag(v.long(), tables[t], o[:-1].long(), mode=combiner)
        else:
            d[n] = F.embedding(idx[n].long().reshape(-1), tables[t])
    for k, v in dense.items():
        d[k] = v.reshape(-1, 1).float()
    return mlp(torch.cat([d[k] for k in sorted(d)], dim=1), layers)


def contrastive_logits(q, items, ids, false_neg_score: float, temperature: float = 1.0) -> torch.Tensor:
    """retrieval/base.py:339-396 with in-batch negatives."""
    pos = (q * items).sum(-1, keepdim=True)
    neg = q @ items.t()
    neg = torch.where(ids.reshape(-1, 1) == ids.reshape(1, -1), torch.full_like(neg, false_neg_score), neg)
    return torch.cat([pos, neg], dim=1) / temperature


class DLRMTrainCPU:
    """One TRAINING step of the DLRM on the host, as the reference's `train_step` would run it on TF-CPU (models/base.py:
    1121-1177): autograd over `dlrm_forward`'s op sequence (F.embedding with sparse gradients = tf.gather's IndexedSlices),
    binary cross-entropy on the logits, Adagrad (accum += g^2, w -= lr g / (sqrt(accum) + eps)); embedding rows are
    updated lazily from the coalesced (duplicates summed) slices, touching only the looked-up rows — what Keras'
    `_resource_apply_sparse_duplicate_indices` does.  Timed CPU baseline of `bench.py`'s training record only.
    (torch.optim.Adagrad's sparse path was 100x slower on 10 M-row tables; the rule is applied by hand.)"""

    def __init__(self, tables, feature_table, bottom, top, head, lr=0.01, initial_accumulator_value=0.1, eps=1e-7):
        self.f2t = dict(feature_table)
        self.lr, self.eps = float(lr), float(eps)
        self.tables = {n: t.clone().float().requires_grad_(True) for n, t in tables.items()}
        self.layers = {}
        for tag, ls in (("bottom", bottom), ("top", top), ("head", [head])):
            self.layers[tag] = [{"kernel": l["kernel"].clone().float().requires_grad_(True),
                                 "bias": l["bias"].clone().float().requires_grad_(True), "activation": l.get("activation")} for l in ls]
        self.dense = [p for ls in self.layers.values() for l in ls for p in (l["kernel"], l["bias"])]
        self.acc = {id(p): torch.full_like(p, initial_accumulator_value) for p in list(self.tables.values()) + self.dense}

    def step(self, idx, dense, targets) -> float:
        for p in list(self.tables.values()) + self.dense:
            p.grad = None
        emb = {n: F.embedding(idx[n].long().reshape(-1), self.tables[t], sparse=True) for n, t in self.f2t.items()}
        x = torch.cat([dense[k].reshape(-1, 1).float() for k in sorted(dense)], dim=1)
        emb["bottom_block"] = mlp(x, self.layers["bottom"])
        stacked = torch.stack([emb[k] for k in sorted(emb)], dim=1)
        z = torch.bmm(stacked, stacked.transpose(1, 2))
        Fn = stacked.shape[1]
        mask = torch.triu(torch.ones(Fn, Fn, dtype=torch.bool), diagonal=1)
        body = mlp(torch.cat([emb["bottom_block"], z[:, mask]], dim=1), self.layers["top"])
        h = self.layers["head"][0]
        logits = torch.addmm(h["bias"], body, h["kernel"]).reshape(-1)
        loss = F.binary_cross_entropy_with_logits(logits, targets.float().reshape(-1))
        loss.backward()
        with torch.no_grad():
            for p in self.dense:
                a = self.acc[id(p)]
                a.addcmul_(p.grad, p.grad)
                p.addcdiv_(p.grad, a.sqrt().add_(self.eps), value=-self.lr)
            for t in self.tables.values():
                g = t.grad.coalesce()  # duplicates summed
                rows, v = g.indices()[0], g.values()
                a = self.acc[id(t)]
                ar = a[rows] + v * v
                a[rows] = ar
                t[rows] -= self.lr * v / (ar.sqrt() + self.eps)
        return float(loss.item())
