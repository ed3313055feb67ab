"""Re-evaluate the CPU oracle on the inputs stored in a golden fixture and compare with the
stored outputs.  `ref_torch_*.npz` outputs were produced by the reference's own torch backend
(oracle/make_golden_from_reference_torch.py); `oracle_*.npz` by the oracle itself
(tests/golden/make_golden.py) and guard against drift."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from oracle import oracle

RTOL, ATOL = 1e-5, 1e-6


def load(path):
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


def layers_from(z, prefix="", act="relu", n=None):
    out = []
    i = 0
    while f"{prefix}kernel_{i}" in z:
        out.append({"kernel": z[f"{prefix}kernel_{i}"], "bias": z.get(f"{prefix}bias_{i}"), "activation": act})
        i += 1
    return out


def check(path: Path) -> None:
    z = load(path)
    kind = str(z["kind"])
    if kind == "dlrm_interaction":
        names = [str(n) for n in z["names"]]
        feats = {n: z[f"in_{n}"] for n in names}
        stacked = oracle.stack_features(feats, axis=1)
        assert np.array_equal(stacked, z["stacked"]), "sorted-name stack order differs from the reference"
        inter = oracle.dot_interaction(stacked)
        np.testing.assert_allclose(inter, z["interactions"], rtol=RTOL, atol=ATOL)
        # InteractionBlock: cat((inputs["continuous"], interactions)) — bottom/continuous first
        np.testing.assert_allclose(np.concatenate([feats["continuous"], inter], axis=1), z["block_out"], rtol=RTOL, atol=ATOL)
    elif kind == "concat":
        feats = {str(n): z[f"in_{n}"] for n in z["names"]}
        assert np.array_equal(oracle.concat_features(feats), z["out"])
    elif kind == "cross":
        ls = [{"kernel": l["kernel"], "bias": l["bias"]} for l in layers_from(z)]
        np.testing.assert_allclose(oracle.cross_layers(z["x"], ls), z["out"], rtol=1e-4, atol=1e-5)
    elif kind == "mlp":
        np.testing.assert_allclose(oracle.mlp(z["x"], layers_from(z, act="relu")), z["out"], rtol=RTOL, atol=ATOL)
    elif kind == "embedding_bag":
        for mode in ("mean", "sum"):
            got = oracle.embedding_bag(z["table"], z["values"], z["offsets"], mode)
            np.testing.assert_allclose(got, z[f"out_{mode}"], rtol=RTOL, atol=ATOL)
    elif kind == "contrastive":
        # reference torch ContrastiveOutput.contrastive_outputs: [positive | negatives], accidental hits -> MIN_FLOAT
        assert abs(float(z["min_float"]) - oracle.MIN_FLOAT) < 1e-3
        out, tgt = oracle.contrastive_logits(z["query"], z["positive"], z["negative"], z["ids"], z["negative_ids"], True,
                                             float(z["min_float"]))
        np.testing.assert_allclose(out, z["out_downscored"], rtol=RTOL, atol=ATOL)
        assert np.array_equal(out == np.float32(z["min_float"]), z["out_downscored"] == np.float32(z["min_float"]))
        assert np.array_equal(tgt, z["target"])
        if "out_scaled" in z:  # LogitsTemperatureScaler after the rescoring (transforms/bias.py): logits / T
            scaled, _ = oracle.contrastive_logits(z["query"], z["positive"], z["negative"], z["ids"], z["negative_ids"], True,
                                                  float(z["min_float"]), temperature=float(z["temperature"]))
            np.testing.assert_allclose(scaled, z["out_scaled"], rtol=RTOL, atol=ATOL)
        plain, _ = oracle.contrastive_logits(z["query"], z["positive"], z["negative"], None, None, False, float(z["min_float"]))
        np.testing.assert_allclose(plain, z["out_plain"], rtol=RTOL, atol=ATOL)
        resc, valid = oracle.rescore_false_negatives(z["ids"], z["neg2_ids"], z["scores2"], float(z["min_float"]))
        np.testing.assert_allclose(resc, z["rescored2"], rtol=RTOL, atol=ATOL)
        assert np.array_equal(valid, z["valid2"].astype(bool))
    elif kind == "log_uniform":
        # The torch backend normalises by log(R + 1) over R classes, the TF backend (the parity target,
        # outputs/sampling/popularity.py:150-151) by log(R + 2) over R + 1: torch(max_id) == TF(max_id - 1).
        for i, (max_id, min_id, n_sample) in enumerate(z["cases"].tolist()):
            p = oracle.log_uniform_probs(max_id - 1, min_id, unique=False)
            np.testing.assert_allclose(p, z[f"probs_{i}"], rtol=1e-5, atol=1e-7)
            # "sampled at least once": TF computes 1 - (1 - p)^n as -expm1(n * log1p(-p)) (popularity.py:153-158);
            # the torch backend's get_unique_sampling_distr applies log1p to +p, i.e. 1 - (1 + p)^-n — a different
            # quantity.  TF is the parity target: the stored torch vector is only characterised, not matched.
            p64 = p.astype(np.float64)
            np.testing.assert_allclose(z[f"unique_{i}"], 1.0 - (1.0 + p64) ** (-float(n_sample)), rtol=2e-3, atol=2e-5)
            u = oracle.log_uniform_probs(max_id - 1, min_id, unique=True, n_sampled=n_sample)
            np.testing.assert_allclose(u, 1.0 - (1.0 - p64) ** float(n_sample), rtol=2e-3, atol=2e-5)  # p is stored in fp32
    elif kind == "dlrm_block":
        # the reference torch DLRMBlock executed end to end: tables, bottom MLP, sorted stack, interaction,
        # [bottom | interactions], top MLP
        cat = [str(n) for n in z["cat_names"]]
        batch = {k[len("batch_"):]: z[k] for k in z if k.startswith("batch_")}
        tables = {n: z[f"table_{n}"] for n in cat}
        got = oracle.dlrm_forward(batch, tables, {n: n for n in cat}, [str(n) for n in z["cont_names"]],
                                  unpack_layers(z, "bottom"), unpack_layers(z, "top"), None)
        np.testing.assert_allclose(got, z["out"], rtol=1e-4, atol=1e-5)
    elif kind == "mlp_acts":
        for name in [str(n) for n in z["names"]]:
            layers = layers_from(z, prefix=f"{name}_", act=name)
            np.testing.assert_allclose(oracle.mlp(z["x"], layers), z[f"out_{name}"], rtol=1e-4, atol=1e-5, err_msg=name)
    elif kind == "embedding_dims":
        got = [oracle.infer_embedding_dim(int(m) + 1) for m in z["max_id"]]
        assert got == z["dims_default"].tolist()
        got3 = [oracle.infer_embedding_dim(int(m) + 1, multiplier=3.0, ensure_multiple_of_8=False) for m in z["max_id"]]
        assert got3 == z["dims_mult3_plain"].tolist()
        assert int(z["dims_default"][: int(z["n_criteo"])].sum()) == 1024  # d = 1024 + 13 for DCN
    elif kind == "catalog":
        # reference torch EmbeddingTablePrediction: logits = x @ E^T + bias; nn.CrossEntropyLoss on them; top-k
        logits = oracle.catalog_logits(z["x"], z["table"], z["bias"])
        np.testing.assert_allclose(logits, z["logits"], rtol=RTOL, atol=1e-5)
        st = oracle.softmax_ce_stats(logits, z["targets"])
        np.testing.assert_allclose(st[:, 1] - st[:, 2], z["cross_entropy"], rtol=1e-4, atol=1e-5)
        s, ids = oracle.topk(logits, z["topk_scores"].shape[1])
        np.testing.assert_allclose(s, z["topk_scores"], rtol=RTOL, atol=1e-5)
        assert np.array_equal(ids, z["topk_ids"])
    elif kind in ("dlrm_model", "dcn_model"):
        # the reference's torch DLRMModel / DCNModel executed end to end, BinaryOutput (Linear(1) + sigmoid) included
        cat = [str(n) for n in z["cat_names"]]
        batch = {k[len("batch_"):]: z[k] for k in z if k.startswith("batch_")}
        tables = {n: z[f"table_{n}"] for n in cat}
        cont = [str(n) for n in z["cont_names"]]
        head = unpack_layers(z, "head")[0]
        if kind == "dlrm_model":
            got = oracle.dlrm_forward(batch, tables, {n: n for n in cat}, cont, unpack_layers(z, "bottom"),
                                      unpack_layers(z, "top"), head)
        else:
            cross = [{"kernel": l["kernel"], "bias": l["bias"]} for l in unpack_layers(z, "cross")]
            got = oracle.dcn_forward(batch, tables, {n: n for n in cat}, cont, cross, unpack_layers(z, "deep"), head)
        np.testing.assert_allclose(got, z["out"], rtol=1e-4, atol=1e-5)
    elif kind == "dlrm_train":
        # one training step of the reference's torch DLRMModel: its default BinaryOutput loss + torch.autograd through its
        # modules -> loss and the gradient of every variable
        from oracle import oracle_train

        cat = [str(n) for n in z["cat_names"]]
        batch = {k[len("batch_"):]: z[k] for k in z if k.startswith("batch_")}
        tables = {n: z[f"table_{n}"] for n in cat}
        cont = [str(n) for n in z["cont_names"]]
        loss, logits, grads = oracle_train.dlrm_loss_and_grads(batch, tables, {n: n for n in cat}, cont, unpack_layers(z, "bottom"),
                                                                unpack_layers(z, "top"), unpack_layers(z, "head")[0], z["targets"])
        np.testing.assert_allclose(1.0 / (1.0 + np.exp(-logits)), z["out"].reshape(-1), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(loss, float(z["loss"]), rtol=1e-5)
        for name, want in train_grad_items(z):
            np.testing.assert_allclose(grads[name], want, rtol=2e-4, atol=1e-7, err_msg=name)
    elif kind == "two_tower":
        # reference torch towers: TabularInputBlock(continuous + EmbeddingTables(mean combiner), agg="concat") -> MLPBlock,
        # then the backend's retrieval pieces on the tower outputs
        batch = {k[len("batch_"):]: z[k] for k in z if k.startswith("batch_")}
        towers = {}
        for tag in ("query", "item"):
            cols = [str(c) for c in z[f"{tag}_cols"]]
            tables = {c: z[f"{tag}_table_{c}"] for c in cols if f"{tag}_table_{c}" in z}
            cont = [c for c in cols if c not in tables]
            sub = {k: v for k, v in batch.items() if any(k == c or k.startswith(c + "__") for c in cols)}
            layers = [{"kernel": z[f"{tag}_kernel_{i}"], "bias": z[f"{tag}_bias_{i}"], "activation": "relu"} for i in range(2)]
            # the concat the first Dense sees: sorted feature names, embeddings (bags: mean) and continuous columns
            feats = oracle.prepare_features(sub)
            d = {c: oracle.embed_feature(tables[c], feats[c], "mean") for c in tables}
            d.update({c: np.asarray(feats[c], dtype=np.float32) for c in cont})
            np.testing.assert_allclose(oracle.concat_features(d), z[f"{tag}_concat"], rtol=RTOL, atol=ATOL)
            towers[tag] = oracle.tower_forward(sub, tables, {c: c for c in tables}, cont, layers, combiner="mean")
            np.testing.assert_allclose(towers[tag], z[f"{tag}_out"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(oracle.retrieval_scores(towers["query"], towers["item"]), z["inference_scores"], rtol=1e-5, atol=1e-6)
        ids = batch["movieId"]
        logits, targets = oracle.contrastive_logits(towers["query"], towers["item"], towers["item"], ids, ids, downscore=True,
                                                    false_negative_score=float(z["min_float"]))
        np.testing.assert_allclose(logits, z["train_logits"], rtol=1e-5, atol=1e-5)
        assert np.array_equal(targets, z["train_targets"])
    elif kind == "oracle_model":
        spec = json.loads(str(z["spec"]))
        got = run_model_fixture(z, spec)
        np.testing.assert_allclose(got, z["expected"], rtol=RTOL, atol=ATOL)
    else:
        raise AssertionError(f"unknown golden kind {kind!r} in {path}")


def unpack_layers(z, name):
    out = []
    i = 0
    while f"{name}_kernel_{i}" in z:
        b = z[f"{name}_bias_{i}"] if f"{name}_bias_{i}" in z else None
        out.append({"kernel": z[f"{name}_kernel_{i}"], "bias": b, "activation": str(z[f"{name}_act_{i}"])})
        i += 1
    return out


def train_grad_items(z):
    """(oracle gradient name, golden array) pairs of a `dlrm_train` fixture."""
    out = [(f"table/{str(n)}", z[f"grad_table_{str(n)}"]) for n in z["cat_names"]]
    for tag in ("bottom", "top"):
        i = 0
        while f"grad_{tag}_kernel_{i}" in z:
            out.append((f"{tag}/kernel_{i}", z[f"grad_{tag}_kernel_{i}"]))
            out.append((f"{tag}/bias_{i}", z[f"grad_{tag}_bias_{i}"]))
            i += 1
    out.append(("head/kernel", z["grad_head_kernel_0"]))
    out.append(("head/bias", z["grad_head_bias_0"]))
    return out


def run_model_fixture(z, spec):
    batch = {k[len("batch_"):]: z[k] for k in z if k.startswith("batch_")}
    tables = {k[len("table_"):]: z[k] for k in z if k.startswith("table_")}
    f2t = spec["feature_table"]
    if spec["model"] == "dlrm":
        return oracle.dlrm_forward(batch, tables, f2t, spec["continuous"], unpack_layers(z, "bottom"),
                                   unpack_layers(z, "top"), unpack_layers(z, "head")[0])
    if spec["model"] == "dcn":
        cross = [{"kernel": l["kernel"], "bias": l["bias"]} for l in unpack_layers(z, "cross")]
        return oracle.dcn_forward(batch, tables, f2t, spec["continuous"], cross, unpack_layers(z, "deep"),
                                  unpack_layers(z, "head")[0])
    if spec["model"] == "two_tower":
        q = oracle.tower_forward(batch, tables, spec["query_feature_table"], spec["query_continuous"], unpack_layers(z, "query"))
        it = oracle.tower_forward(batch, tables, spec["item_feature_table"], spec["item_continuous"], unpack_layers(z, "item"))
        ids = batch[spec["item_id"]]
        out, _ = oracle.contrastive_logits(q, it, it, ids, ids, True, oracle.MIN_FLOAT, temperature=spec["temperature"])
        return out
    raise AssertionError(spec["model"])
