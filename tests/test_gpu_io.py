"""Checkpoint boundary: `model.save` / `Model.load` (merlin/models/tf/models/base.py:1687-1728),
`.merlin` schema metadata (merlin/models/io.py:26-55), `load_weights` from a {name: array} mapping (a Keras
checkpoint exported as numpy), `EmbeddingTable.from_pretrained` / `to_df` (inputs/embedding.py:283-379)."""
import json

import numpy as np
import pytest
import torch

import models_b200 as mm
from models_b200 import datasets
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def small_criteo(card=300):
    return datasets.criteo_schema({f"C{i}": card + i for i in range(1, 27)})


def _dev(batch, device):
    return {k: torch.from_numpy(v).to(device) for k, v in batch.items()}


def test_dlrm_save_load_round_trip(device, tmp_path):
    mm.set_seed(41)
    schema = small_criteo()
    model = mm.DLRMModel(schema, embedding_dim=64, bottom_block=mm.MLPBlock([128, 64]), top_block=mm.MLPBlock([128, 64, 32]))
    batch, _ = datasets.split_targets(schema, datasets.generate_batch(schema, 513, seed=3, index_law="uniform"))
    want = model(_dev(batch, device))
    model.save(tmp_path / "dlrm")
    # layout on disk: Keras-shaped variables with Keras-style names, schemas as the reference stores them
    manifest = json.loads((tmp_path / "dlrm" / "variables" / "manifest.json").read_text())
    named = {e["name"]: e for e in manifest["variables"] if e["name"]}
    assert set(named) == set(model.weights())
    emb = [n for n in named if n.endswith("embeddings")]
    assert len(emb) == 26 and all(named[n]["shape"][1] == 64 for n in emb)
    assert any(n.endswith("kernel") and named[n]["shape"] == [415, 128] for n in named)  # Dense kernel is (in, out)
    arr = np.load(tmp_path / "dlrm" / "variables" / named[emb[0]]["file"])
    assert arr.dtype == np.float32 and np.array_equal(arr, model.weights()[emb[0]].cpu().numpy())
    inp, out = mm.load_merlin_metadata(tmp_path / "dlrm")
    assert inp.column_names == schema.column_names and len(out) == 1
    # no derived caches (split-bf16 weights, scratch) in the file: only variables
    assert all(e["dtype"] == "float32" for e in manifest["variables"])

    loaded = mm.Model.load(tmp_path / "dlrm")
    got = loaded(_dev(batch, device))
    assert torch.equal(got, want)
    cf = loaded.compile(batch)  # the loaded model is a full citizen: graph capture works
    assert torch.equal(cf(mm.HostBatch.like(batch, loaded.input_columns())).to(device), want)


def test_two_tower_save_load_round_trip(device, tmp_path):
    mm.set_seed(42)
    schema = datasets.movielens_1m_schema()
    model = mm.TwoTowerModel(schema, query_tower=mm.MLPBlock([128, 64]))
    raw = datasets.generate_batch(schema, 256, seed=4)
    batch, _ = datasets.split_targets(schema, raw)
    d = _dev(batch, device)
    want_inf = model(d)
    want_tr = model(d, training=True)
    model.save(tmp_path / "tt")
    loaded = mm.Model.load(tmp_path / "tt")
    assert torch.equal(loaded(d), want_inf)
    got_tr = loaded(d, training=True)
    assert torch.equal(got_tr.outputs, want_tr.outputs) and torch.equal(got_tr.targets, want_tr.targets)


def test_load_weights_from_keras_style_mapping(device):
    """A second model with different random weights takes the first model's arrays through a
    {name: array} mapping with a name map (as an exported Keras checkpoint would arrive) and then
    agrees with it and with the NumPy oracle."""
    schema = small_criteo(200)
    batch, _ = datasets.split_targets(schema, datasets.generate_batch(schema, 300, seed=5, index_law="uniform"))
    d = _dev(batch, device)

    def make(seed):
        mm.set_seed(seed)
        m = mm.DLRMModel(schema, embedding_dim=64, bottom_block=mm.MLPBlock([32, 64]), top_block=mm.MLPBlock([64, 16]))
        m.build(device)
        return m

    a, b = make(1), make(2)
    assert not torch.allclose(a(d), b(d))
    exported = {"keras/" + k.replace("/", "."): v for k, v in a.state_dict().items()}  # foreign naming
    # the same block names are auto-numbered differently in the second model: map by position in weights()
    order_a, order_b = list(a.weights()), list(b.weights())
    to_a = dict(zip(order_b, order_a))
    loaded = b.load_weights(exported, name_map=lambda n: "keras/" + to_a[n].replace("/", "."))
    assert len(loaded) == len(order_b)
    assert torch.equal(b(d), a(d))  # derived split-bf16 weights were rebuilt from the new kernels
    body = a.body
    ref = oracle.dlrm_forward(batch, {f: t.embeddings.cpu().numpy() for f, t in body.embeddings.feature_to_table.items()},
                              {f: f for f in body.embeddings.feature_to_table}, body.continuous.features,
                              H.mlp_layers(body.bottom_block), H.mlp_layers(body.top_block), H.head_layer(a.prediction))
    assert H.rel_err(b(d).cpu().numpy(), ref) < 2e-4
    with pytest.raises(KeyError):
        b.load_weights({}, strict=True)
    assert b.load_weights({}, strict=False) == {}
    bad = dict(exported)
    k0 = next(iter(bad))
    bad[k0] = np.zeros((3, 3), np.float32)
    with pytest.raises(ValueError, match="shape"):
        b.load_weights(bad, name_map=lambda n: "keras/" + to_a[n].replace("/", "."))


def test_embedding_table_from_pretrained_and_to_df(device):
    import pandas as pd

    rng = np.random.default_rng(6)
    w = rng.standard_normal((50, 16)).astype(np.float32)
    t = mm.EmbeddingTable.from_pretrained(pd.DataFrame(w), name="item_id")
    assert t.input_dim == 50 and t.dim == 16
    ids = torch.tensor([0, 49, 7, 7], device=device)
    assert np.array_equal(t(ids).cpu().numpy(), w[[0, 49, 7, 7]])
    df = t.to_df()
    assert df.shape == (50, 16) and np.array_equal(df.to_numpy(), w)
    t2 = mm.EmbeddingTable.from_dataset(df, name="item_id")
    assert np.array_equal(t2.embeddings.cpu().numpy(), w)
    with pytest.raises(ValueError, match="`name` is required"):
        mm.EmbeddingTable.from_pretrained(w)


def test_loader_feeds_models_like_direct_batches(device, tmp_path):
    """parquet -> Loader (one pinned pack + one H2D per batch, background thread) -> model == the same rows fed
    directly; ragged list features arrive as `__values` / `__offsets`; host_batches() feeds the CUDA-graph runtime."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    mm.set_seed(43)
    schema = small_criteo(250)
    raw = datasets.generate_batch(schema, 1300, seed=9, index_law="uniform")
    pq.write_table(pa.table({k: np.asarray(v).reshape(-1) for k, v in raw.items()}), tmp_path / "criteo.parquet", row_group_size=200)
    feats, labels = datasets.split_targets(schema, raw)
    model = mm.DLRMModel(schema, embedding_dim=64, bottom_block=mm.MLPBlock([128, 64]), top_block=mm.MLPBlock([128, 64, 32]))
    loader = mm.Loader(str(tmp_path / "criteo.parquet"), batch_size=512, shuffle=False, schema=schema)
    assert len(loader) == 3 and loader.device.type == "cuda"
    label_name = loader.label_names[0]
    s = 0
    for inputs, targets in loader:
        n = targets.shape[0]
        assert all(v.is_cuda for v in inputs.values()) and inputs["C1"].dtype == torch.int32
        direct = {k: torch.from_numpy(np.asarray(v[s:s + n])).to(device) for k, v in feats.items()}
        assert torch.equal(model(inputs), model(direct))
        assert np.array_equal(targets.cpu().numpy().reshape(-1), np.asarray(raw[label_name]).reshape(-1)[s:s + n])
        s += n
    assert s == 1300
    # packed pinned batches straight into the graph runtime (it owns the H2D)
    full = mm.Loader(str(tmp_path / "criteo.parquet"), batch_size=512, shuffle=False, schema=schema, drop_last=True,
                     feature_columns=model.input_columns())
    hbs = list(full.host_batches())
    cf = model.compile(hbs[0])
    for i, hb in enumerate(hbs):
        direct = {k: torch.from_numpy(np.asarray(v[i * 512:(i + 1) * 512]).astype(hb.spec[k][1])).to(device) for k, v in feats.items()}
        assert torch.equal(cf(hb).to(device), model(direct))

    # two-tower with a ragged multi-hot feature, shuffled
    schema2 = datasets.movielens_1m_schema()
    raw2 = datasets.generate_batch(schema2, 700, seed=10)
    tt = mm.TwoTowerModel(schema2, query_tower=mm.MLPBlock([64, 32]))
    loader2 = mm.Loader(raw2, batch_size=256, shuffle=True, schema=schema2, seed_fn=lambda: 3)
    order = np.random.default_rng(3).permutation(700)
    from models_b200 import topk

    s = 0
    for inputs, _ in loader2:
        n = inputs["userId"].shape[0]
        want = topk.take_rows({k: v for k, v in raw2.items() if k in inputs or k.split("__")[0] + "__offsets" in inputs}, order[s:s + n])
        direct = {k: torch.from_numpy(np.asarray(v)).to(device) for k, v in want.items()}
        assert torch.allclose(tt(inputs), tt(direct), rtol=0, atol=0)
        s += n
    assert s == 700
