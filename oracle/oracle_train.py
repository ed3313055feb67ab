"""CPU restatement of ONE TRAINING STEP of the DLRM path — TEST INFRASTRUCTURE (never imported by models_b200/).

The reference's training step (merlin/models/tf/models/base.py:1121-1177, `train_step`) is
    tf.GradientTape over the forward  ->  compute_loss  ->  optimizer.minimize(loss, trainable_variables)
so its algorithm is "autodiff of the forward restated in oracle/oracle_torch.py, then the Keras optimizer rule".
This file restates exactly that with torch.autograd on the CPU (float64 by default: the GPU path is compared with a
tolerance, the reference value should not carry fp32 noise of its own) and plain NumPy optimizer formulas:

  * loss: BinaryOutput's default binary cross-entropy, mean over the batch.  Keras evaluates it on the logits the sigmoid
    activation caches (`_keras_logits`), i.e. max(z,0) - z y + log(1 + exp(-|z|)); the reference's torch backend uses
    nn.BCELoss on the sigmoid outputs (torch/outputs/classification.py:44) — the same function up to fp rounding.
  * embedding gradients: tf.gather's gradient is an IndexedSlices (values (B, D), indices = the batch's ids); the
    optimizers first sum duplicate ids (`_resource_apply_sparse_duplicate_indices`), then update each touched row ONCE.
    `dense_table_grads` returns the dense equivalent (scatter-add), `sparse_update` the row-wise update.
  * SGD / Adagrad (optimizer_v2: accum += g^2; w -= lr g / (sqrt(accum) + eps), accum_0 = 0.1) / Adam
    (lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); w -= lr_t m / (sqrt(v) + eps)); rows are updated lazily
    (LazyAdam, merlin/models/tf/blocks/optimizer.py:342: only the looked-up rows move).

Pinned by tests/golden/ref_torch_dlrm_train.npz: loss and every gradient of the reference's torch DLRMModel, executed in the
build container by oracle/make_golden_from_reference_torch.py (section 9a) — tests/test_oracle.py compares.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def _act(x, name):
    if name in (None, "linear"):
        return x
    if name == "relu":
        return torch.relu(x)
    raise NotImplementedError(name)


def dlrm_loss_and_grads(batch: Dict[str, np.ndarray], tables: Dict[str, np.ndarray], feature_table: Dict[str, str],
                        continuous: Sequence[str], bottom: List[dict], top: List[dict], head: dict, targets: np.ndarray,
                        sample_weight: Optional[np.ndarray] = None, dtype=torch.float64):
    """Forward as oracle/oracle_torch.py:dlrm_forward (same staging, same orders), BCE on the logits, autograd.
    Returns (loss, logits (B,), grads) with grads = {"table/<name>": dense (rows, D), "bottom/kernel_i", "bottom/bias_i",
    "top/...", "head/kernel", "head/bias"}."""
    P = {}
    for n, t in tables.items():
        P[f"table/{n}"] = torch.tensor(np.asarray(t), dtype=dtype, requires_grad=True)
    for tag, layers in (("bottom", bottom), ("top", top)):
        for i, l in enumerate(layers):
            P[f"{tag}/kernel_{i}"] = torch.tensor(np.asarray(l["kernel"]), dtype=dtype, requires_grad=True)
            if l.get("bias") is not None:
                P[f"{tag}/bias_{i}"] = torch.tensor(np.asarray(l["bias"]), dtype=dtype, requires_grad=True)
    P["head/kernel"] = torch.tensor(np.asarray(head["kernel"]), dtype=dtype, requires_grad=True)
    if head.get("bias") is not None:
        P["head/bias"] = torch.tensor(np.asarray(head["bias"]), dtype=dtype, requires_grad=True)

    def mlp(x, tag, layers):
        for i, l in enumerate(layers):
            x = x @ P[f"{tag}/kernel_{i}"]
            if f"{tag}/bias_{i}" in P:
                x = x + P[f"{tag}/bias_{i}"]
            x = _act(x, l.get("activation"))
        return x

    emb = {n: F.embedding(torch.as_tensor(np.asarray(batch[n]).reshape(-1).astype(np.int64)), P[f"table/{t}"])
           for n, t in feature_table.items()}
    x = torch.cat([torch.as_tensor(np.asarray(batch[k], dtype=np.float64).reshape(-1, 1)).to(dtype) for k in sorted(continuous)], dim=1)
    emb["bottom_block"] = mlp(x, "bottom", bottom)
    stacked = torch.stack([emb[k] for k in sorted(emb)], dim=1)
    z = torch.bmm(stacked, stacked.transpose(1, 2))
    Fn = stacked.shape[1]
    mask = torch.triu(torch.ones(Fn, Fn, dtype=torch.bool), diagonal=1)
    body = mlp(torch.cat([emb["bottom_block"], z[:, mask]], dim=1), "top", top)
    logits = (body @ P["head/kernel"]).reshape(-1)
    if "head/bias" in P:
        logits = logits + P["head/bias"].reshape(-1)
    y = torch.as_tensor(np.asarray(targets, dtype=np.float64).reshape(-1)).to(dtype)
    per = torch.clamp(logits, min=0) - logits * y + torch.log1p(torch.exp(-logits.abs()))
    if sample_weight is not None:
        per = per * torch.as_tensor(np.asarray(sample_weight, dtype=np.float64).reshape(-1)).to(dtype)
    loss = per.sum() / y.shape[0]
    loss.backward()
    grads = {k: (v.grad.numpy().copy() if v.grad is not None else np.zeros(tuple(v.shape))) for k, v in P.items()}
    return float(loss.item()), logits.detach().numpy().copy(), grads


def dense_update(opt: str, w: np.ndarray, g: np.ndarray, state: dict, lr: float, beta_1: float = 0.9, beta_2: float = 0.999,
                 epsilon: float = 1e-7, step: int = 1) -> np.ndarray:
    """One Keras optimizer update of a dense variable; `state` holds the slots ("a" / "m", "v"), updated in place."""
    w = np.asarray(w, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    if opt == "sgd":
        return w - lr * g
    if opt == "adagrad":
        state["a"] = state["a"] + g * g
        return w - lr * g / (np.sqrt(state["a"]) + epsilon)
    if opt == "adam":
        state["m"] = beta_1 * state["m"] + (1 - beta_1) * g
        state["v"] = beta_2 * state["v"] + (1 - beta_2) * g * g
        lr_t = lr * np.sqrt(1 - beta_2 ** step) / (1 - beta_1 ** step)
        return w - lr_t * state["m"] / (np.sqrt(state["v"]) + epsilon)
    raise ValueError(opt)


def sparse_update(opt: str, w: np.ndarray, ids: np.ndarray, values: np.ndarray, state: dict, lr: float, **kw) -> np.ndarray:
    """IndexedSlices update: duplicates of an id are summed, every touched row is updated once, other rows (and their
    slots) do not move.  ids outside [0, rows) are dropped (the forward looked up a zero row for them)."""
    w = np.array(w, dtype=np.float64)
    ids = np.asarray(ids).reshape(-1).astype(np.int64)
    ok = (ids >= 0) & (ids < w.shape[0])
    uniq, inv = np.unique(ids[ok], return_inverse=True)
    summed = np.zeros((len(uniq), w.shape[1]))
    np.add.at(summed, inv, np.asarray(values, dtype=np.float64)[ok])
    sub = {k: v[uniq] for k, v in state.items()}
    w[uniq] = dense_update(opt, w[uniq], summed, sub, lr, **kw)
    for k in state:
        state[k][uniq] = sub[k]
    return w
