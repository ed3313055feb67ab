"""Glue between the product (models_b200) and the checker (oracle): pulls the weights of a built
model to the host and evaluates the CPU restatement on the same batch."""
from __future__ import annotations

import numpy as np
import torch

import models_b200 as mm
from oracle import oracle


def to_numpy(t):
    return t.detach().cpu().numpy()


def mlp_layers(mlp):
    """Oracle description of an MLP: one dict per Dense; a BatchNormalization that follows it rides along as
    "batch_norm" (oracle.mlp applies it after the activation, as Keras does at inference)."""
    from models_b200.blocks import BatchNormalization, _Dense

    out = []
    for l in mlp.layers:
        if isinstance(l, _Dense):
            out.append({"kernel": to_numpy(l.kernel), "bias": None if l.bias is None else to_numpy(l.bias),
                        "activation": l.activation})
        elif isinstance(l, BatchNormalization):
            out[-1]["batch_norm"] = {"gamma": to_numpy(l.gamma), "beta": to_numpy(l.beta), "mean": to_numpy(l.moving_mean),
                                     "var": to_numpy(l.moving_variance)}
    return out


def head_layer(out_block):
    d = out_block.to_call
    return {"kernel": to_numpy(d.kernel), "bias": None if d.bias is None else to_numpy(d.bias), "activation": d.activation}


def emb_tables(emb):
    tables = {n: to_numpy(t.embeddings) for n, t in emb.tables.items()}
    f2t = {f: t.table_name for f, t in emb.feature_to_table.items()}
    return tables, f2t


def oracle_dlrm(model, batch, return_intermediates=False):
    body = model.body
    tables, f2t = emb_tables(body.embeddings)
    cont = body.continuous.features if body.continuous is not None else []
    bottom = mlp_layers(body.bottom_block) if body.bottom_block is not None else None
    top = mlp_layers(body.top_block) if body.top_block is not None else None
    return oracle.dlrm_forward(batch, tables, f2t, cont, bottom, top, head_layer(model.prediction),
                               return_intermediates=return_intermediates)


def oracle_dcn(model, batch):
    body = model.body
    tables, f2t = emb_tables(body.input_block.embeddings)
    cont = body.input_block.continuous.features if body.input_block.continuous is not None else []
    cross = [{"kernel": to_numpy(l.dense.kernel), "bias": None if l.dense.bias is None else to_numpy(l.dense.bias)}
             for l in body.cross.cross_layers]
    return oracle.dcn_forward(batch, tables, f2t, cont, cross, mlp_layers(body.deep), head_layer(model.prediction),
                              stacked=body.stacked, branch_order=body.branch_order())


def oracle_deepfm(model, batch):
    body = model.body
    tables, f2t = emb_tables(body.input_block.embeddings)
    f2t = {f: t for f, t in f2t.items() if f in body.fm.cat_names}
    card = {f: body.input_block.embeddings.feature_to_table[f].input_dim for f in body.fm.cat_names}
    wide = body.fm.wide
    return oracle.deepfm_forward(batch, tables, f2t, body.fm.cont_names, card, to_numpy(wide.kernel), to_numpy(wide.bias),
                                 mlp_layers(body.deep), mlp_layers(body.deep_logit), head_layer(model.prediction))


def oracle_tower(tower, batch, l2=False):
    tables, f2t = emb_tables(tower.inputs.embeddings) if tower.inputs.embeddings is not None else ({}, {})
    cont = tower.inputs.continuous.features if tower.inputs.continuous is not None else []
    comb = "mean"
    if tower.inputs.embeddings is not None:
        comb = next(iter(tower.inputs.embeddings.tables.values())).sequence_combiner or "mean"
    return oracle.tower_forward(batch, tables, f2t, cont, mlp_layers(tower.mlp), combiner=comb, l2_normalize=l2)


def device_batch(batch, device):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in batch.items()}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))
