"""CUDA path against the COMMITTED golden fixtures: outputs of the reference's own torch backend
(tests/golden/ref_torch_*.npz) and the oracle's model-level vectors (oracle_*.npz)."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

import models_b200 as mm
from models_b200 import datasets, ops
from tests.golden import replay

pytestmark = pytest.mark.gpu
G = Path(__file__).parent / "golden"
RTOL, ATOL = 1e-4, 1e-5


def dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def test_reference_torch_dlrm_interaction(device):
    z = replay.load(G / "ref_torch_dlrm_interaction.npz")
    names = [str(n) for n in z["names"]]
    B, D = z[f"in_{names[0]}"].shape
    F = len(names)
    # StackFeatures layout produced by the concat kernel in sorted-name order
    stack = torch.empty((B, F * D), dtype=torch.float32, device=device)
    ops.concat_columns([dev(z[f"in_{n}"], device) for n in sorted(names)], stack)
    assert np.array_equal(stack.cpu().numpy().reshape(B, F, D), z["stacked"])
    out = torch.empty((B, D + F * (F - 1) // 2), dtype=torch.float32, device=device)
    ops.dot_interaction(stack.view(B, F, D), out, prefix=dev(z["in_continuous"], device))
    # tensor-core interaction (3-pass split-bf16): |err| ~ 2^-16 * sum|x_k y_k| -> absolute, not relative
    np.testing.assert_allclose(out.cpu().numpy(), z["block_out"], rtol=RTOL, atol=2e-4)
    np.testing.assert_allclose(out.cpu().numpy()[:, D:], z["interactions"], rtol=RTOL, atol=2e-4)


def test_reference_torch_concat(device):
    z = replay.load(G / "ref_torch_concat.npz")
    names = sorted(str(n) for n in z["names"])
    out = torch.empty(z["out"].shape, dtype=torch.float32, device=device)
    ops.concat_columns([dev(z[f"in_{n}"], device) for n in names], out)
    assert np.array_equal(out.cpu().numpy(), z["out"])


def test_reference_torch_cross_and_mlp(device):
    z = replay.load(G / "ref_torch_cross.npz")
    x0 = dev(z["x"], device)
    x = x0
    i = 0
    while f"kernel_{i}" in z:
        out = torch.empty_like(x0)
        x = ops.dense_fp32(x, dev(z[f"kernel_{i}"], device), dev(z[f"bias_{i}"], device), None, out, x0=x0)
        i += 1
    assert i == 3
    np.testing.assert_allclose(x.cpu().numpy(), z["out"], rtol=RTOL, atol=ATOL)
    m = replay.load(G / "ref_torch_mlp.npz")
    h = dev(m["x"], device)
    i = 0
    while f"kernel_{i}" in m:
        W = dev(m[f"kernel_{i}"], device)
        o = torch.empty((h.shape[0], W.shape[1]), dtype=torch.float32, device=device)
        h = ops.dense_fp32(h, W, dev(m[f"bias_{i}"], device), "relu", o)
        i += 1
    np.testing.assert_allclose(h.cpu().numpy(), m["out"], rtol=RTOL, atol=ATOL)


def test_reference_torch_embedding_bag(device):
    z = replay.load(G / "ref_torch_embedding_bag.npz")
    B = len(z["offsets"]) - 1
    for mode in ("mean", "sum"):
        out = torch.empty((B, z["table"].shape[1]), dtype=torch.float32, device=device)
        ops.gather_bag(dev(z["table"], device), dev(z["values"], device), dev(z["offsets"], device), mode, out)
        np.testing.assert_allclose(out.cpu().numpy(), z[f"out_{mode}"], rtol=1e-6, atol=1e-7)


def _set_mlp(mlp, layers):
    assert len(mlp.dense_layers) == len(layers)
    for l, w in zip(mlp.dense_layers, layers):
        assert l.activation == w["activation"]
        l.set_weights(w["kernel"], w["bias"])


def _set_tables(emb, z):
    for name, t in emb.tables.items():
        t.table = torch.from_numpy(z[f"table_{name}"]).cuda().contiguous()
        assert t.table.shape == (t.input_dim, t.dim)
        t.built = True


def test_golden_dlrm_model(device):
    z = replay.load(G / "oracle_dlrm_criteo_small.npz")
    schema = datasets.criteo_schema({k: min(v, 200) for k, v in datasets.CRITEO_MAX.items()})
    model = mm.DLRMModel(schema, embedding_dim=16, bottom_block=mm.MLPBlock([32, 16]), top_block=mm.MLPBlock([32, 16, 8]))
    _set_tables(model.body.embeddings, z)
    _set_mlp(model.body.bottom_block, replay.unpack_layers(z, "bottom"))
    _set_mlp(model.body.top_block, replay.unpack_layers(z, "top"))
    h = replay.unpack_layers(z, "head")[0]
    model.prediction.to_call.set_weights(h["kernel"], h["bias"])
    batch = {k[len("batch_"):]: dev(z[k], device) for k in z if k.startswith("batch_")}
    for fused in (True, False):
        model.body.fused = fused
        np.testing.assert_allclose(model(batch).cpu().numpy(), z["expected"], rtol=2e-4, atol=2e-6)


def test_golden_dcn_model(device):
    z = replay.load(G / "oracle_dcn_criteo_small.npz")
    schema = datasets.criteo_schema({k: min(v, 200) for k, v in datasets.CRITEO_MAX.items()})
    model = mm.DCNModel(schema, depth=3, deep_block=mm.MLPBlock([32, 16]))
    _set_tables(model.body.input_block.embeddings, z)
    cross = replay.unpack_layers(z, "cross")
    d = cross[0]["kernel"].shape[0]
    for l, w in zip(model.body.cross.cross_layers, cross):
        l.build(d, device)
        l.dense.set_weights(w["kernel"], w["bias"])
    _set_mlp(model.body.deep, replay.unpack_layers(z, "deep"))
    h = replay.unpack_layers(z, "head")[0]
    model.prediction.to_call.set_weights(h["kernel"], h["bias"])
    batch = {k[len("batch_"):]: dev(z[k], device) for k in z if k.startswith("batch_")}
    np.testing.assert_allclose(model(batch).cpu().numpy(), z["expected"], rtol=2e-4, atol=2e-6)


def test_golden_two_tower_model(device):
    z = replay.load(G / "oracle_two_tower_ml1m_small.npz")
    spec = json.loads(str(z["spec"]))
    schema = datasets.movielens_1m_schema()
    model = mm.TwoTowerModel(schema, query_tower=mm.MLPBlock([32, 16]), item_tower=mm.MLPBlock([32, 16]),
                             embedding_options=mm.EmbeddingOptions(embedding_dim_default=16),
                             logits_temperature=spec["temperature"])
    _set_tables(model.body.query.inputs.embeddings, z)
    _set_tables(model.body.item.inputs.embeddings, z)
    _set_mlp(model.body.query.mlp, replay.unpack_layers(z, "query"))
    _set_mlp(model.body.item.mlp, replay.unpack_layers(z, "item"))
    batch = {k[len("batch_"):]: dev(z[k], device) for k in z if k.startswith("batch_")}
    pred = model(batch, training=True)
    np.testing.assert_allclose(pred.outputs.cpu().numpy(), z["expected"], rtol=2e-4, atol=2e-5)


def test_reference_torch_contrastive_logits(device):
    """Outputs of the reference's own torch ContrastiveOutput.contrastive_outputs / rescore_false_negatives
    (tests/golden/ref_torch_contrastive.npz): [positive | negatives] layout, accidental hits == MIN_FLOAT exactly,
    one-hot targets — against the CUDA scorer (tensor-core and exact-fp32 engines)."""
    z = replay.load(G / "ref_torch_contrastive.npz")
    q, pos, ids = dev(z["query"], device), dev(z["positive"], device), dev(z["ids"], device)
    fns = float(z["min_float"])
    out = mm.ContrastiveOutput(negative_samplers="in-batch", false_negative_score=fns)(
        {"query": q, "candidate": pos}, candidate_ids=ids, training=True)
    got = out.outputs.cpu().numpy()
    hit = z["out_downscored"] == np.float32(fns)
    assert np.array_equal(got == np.float32(fns), hit)            # the mask is exact (integer compare)
    np.testing.assert_allclose(got[~hit], z["out_downscored"][~hit], rtol=RTOL, atol=2e-4)
    assert np.array_equal(out.targets.cpu().numpy(), z["target"])
    # LogitsTemperatureScaler applied after the rescoring: every logit (accidental hits included) divided by T
    T = float(z["temperature"])
    scaled = mm.ContrastiveOutput(negative_samplers="in-batch", false_negative_score=fns, logits_temperature=T)(
        {"query": q, "candidate": pos}, candidate_ids=ids, training=True).outputs.cpu().numpy()
    np.testing.assert_allclose(scaled[hit], z["out_scaled"][hit], rtol=1e-6, atol=0)
    np.testing.assert_allclose(scaled[~hit], z["out_scaled"][~hit], rtol=RTOL, atol=2e-4 / T)
    # exact-fp32 engine and a separate negative set (N != B)
    B, N = z["scores2"].shape
    buf = torch.empty((B, 1 + N), dtype=torch.float32, device=device)
    ops.inbatch_scores(q, pos, dev(z["neg2"], device), buf, pos_ids=ids, neg_ids=dev(z["neg2_ids"], device), false_neg_score=fns,
                       tensor_cores=False)
    np.testing.assert_allclose(buf.cpu().numpy()[:, 1:], z["rescored2"], rtol=RTOL, atol=ATOL)
    plain = torch.empty((B, 1 + B), dtype=torch.float32, device=device)
    ops.inbatch_scores(q, pos, pos, plain, downscore=False, tensor_cores=False)
    np.testing.assert_allclose(plain.cpu().numpy(), z["out_plain"], rtol=RTOL, atol=ATOL)


def test_reference_torch_log_uniform_probabilities():
    """mm.log_uniform_sampling_probs follows the TF formula (normaliser log(R + 2)); the reference's torch backend
    computes the same expression one class short — torch(max_id) == ours(max_id - 1) (tests/golden/replay.py)."""
    z = replay.load(G / "ref_torch_log_uniform.npz")
    for i, (max_id, min_id, n_sample) in enumerate(z["cases"].tolist()):
        p = mm.log_uniform_sampling_probs(max_id - 1, min_id, unique=False)
        np.testing.assert_allclose(p, z[f"probs_{i}"], rtol=1e-5, atol=1e-7)
        # unique=True follows TF's 1 - (1 - p)^n; the torch backend's variant (1 - (1 + p)^-n) is not the target
        u = mm.log_uniform_sampling_probs(max_id - 1, min_id, max_num_samples=n_sample, unique=True)
        np.testing.assert_allclose(u, 1.0 - (1.0 - p.astype(np.float64)) ** float(n_sample), rtol=2e-3, atol=2e-5)  # p is stored in fp32


def test_reference_torch_dlrm_block_end_to_end(device):
    """The reference's torch DLRMBlock executed end to end in the build container
    (tests/golden/ref_torch_dlrm_block.npz: schema -> embeddings -> bottom MLP -> sorted stack -> interaction ->
    [bottom | interactions] -> top MLP) against mm.DLRMBlock with the same tables and kernels — fused and staged
    paths, tensor-core and exact-fp32 dense engines."""
    from models_b200 import blocks
    from models_b200.schema import ColumnSchema, Schema

    z = replay.load(G / "ref_torch_dlrm_block.npz")
    cat = [str(n) for n in z["cat_names"]]
    cols = [ColumnSchema(n, tags=("categorical",), dtype="int64", properties={"domain": {"min": 0, "max": int(mx), "name": n}})
            for n, mx in zip(cat, z["cat_max"])]
    cols += [ColumnSchema(str(n), tags=("continuous",), dtype="float32") for n in z["cont_names"]]
    dim = int(z["dim"])
    body = mm.DLRMBlock(Schema(cols), embedding_dim=dim, bottom_block=mm.MLPBlock([32, dim]), top_block=mm.MLPBlock([24, 8]))
    _set_tables(body.embeddings, z)
    _set_mlp(body.bottom_block, replay.unpack_layers(z, "bottom"))
    _set_mlp(body.top_block, replay.unpack_layers(z, "top"))
    body.build(device)  # nothing left to initialise: every variable was assigned above
    batch = {k[len("batch_"):]: dev(z[k], device) for k in z if k.startswith("batch_")}
    for engine in ("auto", "fp32"):
        blocks.set_dense_engine(engine)
        try:
            for fused in (True, False):
                body.fused = fused
                got = body(batch).cpu().numpy()
                assert got.shape == z["out"].shape
                np.testing.assert_allclose(got, z["out"], rtol=2e-4, atol=2e-5)
        finally:
            blocks.set_dense_engine("auto")


def _ref_schema(z, with_target=True):
    from models_b200.schema import ColumnSchema, Schema

    lists = {str(n) for n in z["list_names"]} if "list_names" in z else set()
    cols = []
    for n, mx in zip(z["cat_names"], z["cat_max"]):
        n = str(n)
        props = {"domain": {"min": 0, "max": int(mx), "name": n}}
        if n in lists:
            props["value_count"] = {"min": 1, "max": 4}
        cols.append(ColumnSchema(n, tags=("categorical",), dtype="int64", is_list=n in lists, is_ragged=n in lists, properties=props))
    cols += [ColumnSchema(str(n), tags=("continuous",), dtype="float32") for n in z["cont_names"]]
    if with_target:
        cols.append(ColumnSchema("click", tags=("target", "binary_classification"), dtype="int64"))
    return Schema(cols)


def test_reference_torch_dlrm_model_end_to_end(device):
    """merlin.models.torch DLRMModel (DLRMBlock + BinaryOutput) executed in the build container -> mm.DLRMModel."""
    z = replay.load(G / "ref_torch_dlrm_model.npz")
    dim = int(z["dim"])
    model = mm.DLRMModel(_ref_schema(z), embedding_dim=dim, bottom_block=mm.MLPBlock([32, dim]), top_block=mm.MLPBlock([24, 8]))
    _set_tables(model.body.embeddings, z)
    _set_mlp(model.body.bottom_block, replay.unpack_layers(z, "bottom"))
    _set_mlp(model.body.top_block, replay.unpack_layers(z, "top"))
    h = replay.unpack_layers(z, "head")[0]
    model.prediction.to_call.set_weights(h["kernel"], h["bias"])
    batch = {k[len("batch_"):]: dev(z[k], device) for k in z if k.startswith("batch_")}
    for fused in (True, False):
        model.body.fused = fused
        got = model(batch).cpu().numpy()
        np.testing.assert_allclose(got, z["out"], rtol=2e-4, atol=2e-6)
    cf = model.compile({k: v.cpu().numpy() for k, v in batch.items()})  # the CUDA-graph runtime gives the same numbers
    hb = mm.HostBatch.like({k: v.cpu().numpy() for k, v in batch.items()}, model.input_columns())
    np.testing.assert_allclose(cf(hb).numpy(), z["out"], rtol=2e-4, atol=2e-6)


def test_reference_torch_dcn_model_end_to_end(device):
    """merlin.models.torch DCNModel (embeddings + continuous -> sorted concat -> CrossBlock(3) -> MLP -> BinaryOutput)."""
    z = replay.load(G / "ref_torch_dcn_model.npz")
    dims = {str(n): int(d) for n, d in zip(z["cat_names"], z["emb_dims"])}
    model = mm.DCNModel(_ref_schema(z), depth=3, deep_block=mm.MLPBlock([32, 16]), dim=dims)
    _set_tables(model.body.input_block.embeddings, z)
    cross = replay.unpack_layers(z, "cross")
    d = cross[0]["kernel"].shape[0]
    assert d == sum(dims.values()) + len(z["cont_names"])
    for l, w in zip(model.body.cross.cross_layers, cross):
        l.build(d, device)
        l.dense.set_weights(w["kernel"], w["bias"])
    _set_mlp(model.body.deep, replay.unpack_layers(z, "deep"))
    h = replay.unpack_layers(z, "head")[0]
    model.prediction.to_call.set_weights(h["kernel"], h["bias"])
    batch = {k[len("batch_"):]: dev(z[k], device) for k in z if k.startswith("batch_")}
    np.testing.assert_allclose(model(batch).cpu().numpy(), z["out"], rtol=2e-4, atol=2e-6)


def test_reference_torch_catalog_logits_ce_and_topk(device):
    """merlin.models.torch EmbeddingTablePrediction (x @ E^T + bias) + the backend's nn.CrossEntropyLoss + top-k,
    executed in the build container -> mm.CategoricalOutput: materialised logits, and the fused kernel's
    log-sum-exp statistics / top-k that never materialise them."""
    from models_b200.schema import ColumnSchema

    z = replay.load(G / "ref_torch_catalog.npz")
    n_items, D = z["table"].shape
    col = ColumnSchema("item_id", tags=("categorical", "item_id"), dtype="int64",
                       properties={"domain": {"min": 0, "max": n_items - 1, "name": "item_id"}})
    table = mm.EmbeddingTable(D, col, embeddings_initializer=z["table"])
    out = mm.CategoricalOutput(table)
    out.build(device)
    out.bias.copy_(dev(z["bias"], device))
    x, t = dev(z["x"], device), dev(z["targets"], device)
    np.testing.assert_allclose(out(x).cpu().numpy(), z["logits"], rtol=1e-4, atol=2e-4)
    stats = out.softmax_ce_stats(x, t).cpu().numpy()           # (max, log-sum-exp, target logit) per row
    np.testing.assert_allclose(stats[:, 1] - stats[:, 2], z["cross_entropy"], rtol=1e-4, atol=3e-4)
    scores, ids = out.top_k(x, z["topk_scores"].shape[1])
    np.testing.assert_allclose(scores.cpu().numpy(), z["topk_scores"], rtol=1e-4, atol=2e-4)
    gap = np.abs(np.diff(z["topk_scores"], axis=1)).min(axis=1) > 1e-3
    assert gap.any() and np.array_equal(ids.cpu().numpy()[gap], z["topk_ids"][gap])


def test_reference_torch_dcn_with_ragged_multi_hot_feature(device):
    """merlin.models.torch DCNModel over a schema with a list column fed as `genres__values` / `genres__offsets`
    (the backend's default bag combiner is "mean"): pins the ragged input convention and the bag lookup in context."""
    z = replay.load(G / "ref_torch_dcn_multihot.npz")
    dims = {str(n): int(d) for n, d in zip(z["cat_names"], z["emb_dims"])}
    model = mm.DCNModel(_ref_schema(z), depth=2, deep_block=mm.MLPBlock([16, 8]), dim=dims)
    _set_tables(model.body.input_block.embeddings, z)
    cross = replay.unpack_layers(z, "cross")
    d = cross[0]["kernel"].shape[0]
    for l, w in zip(model.body.cross.cross_layers, cross):
        l.build(d, device)
        l.dense.set_weights(w["kernel"], w["bias"])
    _set_mlp(model.body.deep, replay.unpack_layers(z, "deep"))
    h = replay.unpack_layers(z, "head")[0]
    model.prediction.to_call.set_weights(h["kernel"], h["bias"])
    batch = {k[len("batch_"):]: dev(z[k], device) for k in z if k.startswith("batch_")}
    assert "genres__values" in batch and "genres__offsets" in batch
    np.testing.assert_allclose(model(batch).cpu().numpy(), z["out"], rtol=2e-4, atol=2e-6)
    batch32 = dict(batch, genres__offsets=batch["genres__offsets"].to(torch.int32))  # the loader's int32 offsets
    np.testing.assert_allclose(model(batch32).cpu().numpy(), z["out"], rtol=2e-4, atol=2e-6)


@pytest.mark.parametrize("engine", ["auto", "fp32"])
def test_reference_torch_mlp_activations(device, engine):
    """MLPBlock([24, 9], activation=nn.{Sigmoid,Tanh,SELU,ELU,GELU,ReLU}) of the torch backend -> mm.MLPBlock with the
    same kernels: pins the activation definitions of the GEMM epilogues (erf GELU, SELU/ELU constants)."""
    from models_b200 import blocks

    z = replay.load(G / "ref_torch_mlp_activations.npz")
    x = dev(z["x"], device)
    blocks.set_dense_engine(engine)
    try:
        for name in [str(n) for n in z["names"]]:
            mlp = mm.MLPBlock([24, 9], activation=name)
            _set_mlp(mlp, replay.layers_from(z, prefix=f"{name}_", act=name))
            np.testing.assert_allclose(mlp(x).cpu().numpy(), z[f"out_{name}"], rtol=2e-4, atol=2e-5, err_msg=name)
    finally:
        blocks.set_dense_engine("auto")


def test_reference_torch_two_tower_end_to_end(device):
    """merlin.models.torch towers (TabularInputBlock: continuous + EmbeddingTables(mean) -> sorted concat -> MLPBlock)
    on the ML-1M column names, executed in the build container (ref_torch_two_tower.npz), against mm.TwoTowerModel
    with the same tables and kernels: tower outputs, inference scores (B,1) and the train-mode
    [positive | in-batch negatives] logits with the false-negative rescoring."""
    from models_b200.schema import ColumnSchema, Schema, Tags

    z = replay.load(G / "ref_torch_two_tower.npz")
    dim = int(z["dim"])

    def col(name, tower):
        tags = ("user",) if tower == "query" else ("item",)
        if f"{tower}_table_{name}" in z:
            mx = z[f"{tower}_table_{name}"].shape[0] - 1
            is_list = name == "genres"
            props = {"domain": {"min": 0, "max": int(mx), "name": name}}
            if is_list:
                props["value_count"] = {"min": 1, "max": 4}
            extra = ("user_id",) if name == "userId" else ("item_id",) if name == "movieId" else ()
            return ColumnSchema(name, tags=("categorical",) + tags + extra, dtype="int64", is_list=is_list, is_ragged=is_list,
                                properties=props)
        return ColumnSchema(name, tags=("continuous",) + tags, dtype="float32")

    cols = [col(str(n), "query") for n in z["query_cols"]] + [col(str(n), "item") for n in z["item_cols"]]
    model = mm.TwoTowerModel(Schema(cols), query_tower=mm.MLPBlock([24, 12]), item_tower=mm.MLPBlock([24, 12]),
                             embedding_options=mm.EmbeddingOptions(embedding_dim_default=dim))
    for tag, tower in (("query", model.body.query), ("item", model.body.item)):
        for name, table in tower.inputs.embeddings.tables.items():
            table.table = torch.from_numpy(z[f"{tag}_table_{name}"]).cuda().contiguous()
            assert table.table.shape == (table.input_dim, table.dim)
            table.built = True
        _set_mlp(tower.mlp, [{"kernel": z[f"{tag}_kernel_{i}"], "bias": z[f"{tag}_bias_{i}"], "activation": "relu"} for i in range(2)])
    model.build(device)
    batch = {k[len("batch_"):]: dev(z[k], device) for k in z if k.startswith("batch_")}
    enc = model.body(batch)
    np.testing.assert_allclose(enc["query"].cpu().numpy(), z["query_out"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(enc["item"].cpu().numpy(), z["item_out"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(model(batch).cpu().numpy(), z["inference_scores"], rtol=2e-4, atol=2e-5)
    pred = model(batch, training=True)
    got = pred.predictions.cpu().numpy()
    assert got.shape == z["train_logits"].shape
    hits = z["train_logits"] == z["min_float"]
    assert hits[:, 1:].sum() > z["train_logits"].shape[0]  # the diagonal plus the duplicated item
    assert np.array_equal(got == np.float32(z["min_float"]), hits)
    np.testing.assert_allclose(got, z["train_logits"], rtol=2e-4, atol=5e-5)
    assert np.array_equal(pred.targets.cpu().numpy(), z["train_targets"])
