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
