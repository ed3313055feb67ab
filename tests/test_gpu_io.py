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
