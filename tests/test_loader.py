"""Loader hand-off (reference merlin/models/tf/loader.py:135-420) — host logic on CPU tensors:
parquet / DataFrame / dict sources, the `name__values` + `name__offsets` ragged convention
(transforms/features.py:190-210), target split by tag, shuffle / drop_last / global sharding, dtype policy."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest
import torch

import models_b200 as mm
from models_b200.schema import ColumnSchema, Schema, Tags

SCHEMA_PBTXT = '''
feature { name: "user_id" type: INT int_domain { name: "user_id" max: 999 is_categorical: true }
          annotation { tag: "categorical" tag: "user_id" } }
feature { name: "item_id" type: INT int_domain { name: "item_id" max: 4999 is_categorical: true }
          annotation { tag: "categorical" tag: "item_id" } }
feature { name: "genres" value_count { min: 0 max: 4 } type: INT int_domain { name: "genres" max: 19 is_categorical: true }
          annotation { tag: "categorical" tag: "item" } }
feature { name: "price" type: FLOAT annotation { tag: "continuous" } }
feature { name: "click" type: INT int_domain { max: 1 } annotation { tag: "target" tag: "binary_classification" } }
'''


def make_frame(n=1000, seed=0):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 5, n)
    genres = [rng.integers(1, 20, l).astype(np.int64).tolist() for l in lens]
    return {"row": np.arange(n, dtype=np.int64), "user_id": rng.integers(0, 1000, n).astype(np.int64),
            "item_id": rng.integers(0, 5000, n).astype(np.int64), "genres": genres,
            "price": rng.random(n).astype(np.float64), "click": rng.integers(0, 2, n).astype(np.int64)}


@pytest.fixture()
def dataset_dir(tmp_path):
    f = make_frame()
    d = tmp_path / "data"
    d.mkdir()
    half = 600
    for i, (s, e) in enumerate([(0, half), (half, 1000)]):  # two parquet files, schema.pbtxt next to them
        pq.write_table(pa.table({k: (v[s:e] if not isinstance(v, list) else pa.array(v[s:e], pa.list_(pa.int64()))) for k, v in f.items()}),
                       d / f"part_{i}.parquet", row_group_size=128)
    (d / "schema.pbtxt").write_text(SCHEMA_PBTXT)
    return d, f


def _collect(loader):
    rows = []
    for inputs, targets in loader:
        rows.append((inputs, targets))
    return rows


def test_parquet_directory_in_order(dataset_dir):
    d, f = dataset_dir
    loader = mm.Loader(str(d), batch_size=256, shuffle=False, device="cpu")
    assert loader.schema is not None and loader.label_names == ["click"]
    assert loader.feature_names == ["user_id", "item_id", "genres", "price"] and len(loader) == 4
    batches = _collect(loader)
    assert [b[0]["user_id"].shape[0] for b in batches] == [256, 256, 256, 232]
    user = torch.cat([b[0]["user_id"] for b in batches]).numpy()
    assert user.dtype == np.int32 and np.array_equal(user, f["user_id"])  # ids narrowed to int32 (domain max 999)
    price = torch.cat([b[0]["price"] for b in batches]).numpy()
    assert price.dtype == np.float32 and np.array_equal(price, f["price"].astype(np.float32))
    click = torch.cat([b[1] for b in batches]).numpy()
    assert np.array_equal(click, f["click"])
    # ragged convention: values + int32 offsets of length B+1 starting at 0
    row = 0
    for inputs, _ in batches:
        v, o = inputs["genres__values"].numpy(), inputs["genres__offsets"].numpy()
        assert o.dtype == np.int32 and o[0] == 0 and o[-1] == len(v) and len(o) == inputs["user_id"].shape[0] + 1
        for r in range(len(o) - 1):
            assert v[o[r]:o[r + 1]].tolist() == f["genres"][row]
            row += 1
        assert "genres" not in inputs and "click" not in inputs and "row" not in inputs
    assert row == 1000
    assert loader.output_schema.column_names == ["user_id", "item_id", "genres", "price", "click"]


def test_shuffle_drop_last_and_epochs(dataset_dir):
    d, f = dataset_dir
    loader = mm.Loader(str(d), batch_size=300, shuffle=True, drop_last=True, device="cpu", feature_columns=["row", "item_id", "genres"])
    assert len(loader) == 3
    e1 = np.concatenate([b[0]["row"].numpy() for b in _collect(loader)])
    e2 = np.concatenate([b[0]["row"].numpy() for b in _collect(loader)])
    assert len(e1) == 900 and len(np.unique(e1)) == 900 and not np.array_equal(e1, np.sort(e1))
    assert not np.array_equal(e1, e2)  # a new permutation every epoch
    for inputs, _ in loader:  # shuffled rows keep their own list values
        v, o, rows = inputs["genres__values"].numpy(), inputs["genres__offsets"].numpy(), inputs["row"].numpy()
        for r in range(0, len(rows), 37):
            assert v[o[r]:o[r + 1]].tolist() == f["genres"][int(rows[r])]
            assert int(inputs["item_id"][r]) == f["item_id"][int(rows[r])]
    fixed = mm.Loader(str(d), batch_size=300, shuffle=True, seed_fn=lambda: 7, device="cpu", feature_columns=["row"])
    a = np.concatenate([b[0]["row"].numpy() for b in _collect(fixed)])
    b = np.concatenate([b[0]["row"].numpy() for b in _collect(fixed)])
    assert np.array_equal(a, b) and len(a) == 1000


def test_global_sharding_is_a_partition(dataset_dir):
    d, f = dataset_dir
    seen = []
    for rank in range(3):
        loader = mm.Loader(str(d), batch_size=128, shuffle=True, seed_fn=lambda: 11, global_size=3, global_rank=rank,
                           device="cpu", feature_columns=["row"])
        rows = np.concatenate([b[0]["row"].numpy() for b in _collect(loader)])
        assert abs(len(rows) - 1000 / 3) < 1 and len(loader) == (len(rows) + 127) // 128
        seen.append(rows)
    allrows = np.concatenate(seen)
    assert len(allrows) == 1000 and len(np.unique(allrows)) == 1000
    with pytest.raises(ValueError, match="global_rank"):
        mm.Loader(str(d), batch_size=8, global_size=2, global_rank=2, device="cpu")


def test_sources_peek_and_sample_batch(dataset_dir):
    import pandas as pd

    d, f = dataset_dir
    schema = Schema.load(str(d / "schema.pbtxt"))
    df = pd.DataFrame({k: v for k, v in f.items()})
    from_df = mm.Loader(df, batch_size=100, shuffle=False, schema=schema, device="cpu")
    from_files = mm.Loader([str(d / "part_0.parquet"), str(d / "part_1.parquet")], batch_size=100, shuffle=False, device="cpu")
    a, ta = from_df.peek()
    b, tb = from_files.peek()
    assert sorted(a) == sorted(b) and all(torch.equal(a[k], b[k]) for k in a) and torch.equal(ta, tb)
    assert torch.equal(from_df.peek()[0]["item_id"], a["item_id"])  # peek does not advance
    # dict of arrays in the model's own convention (what datasets.generate_batch emits)
    raw = {"user_id": f["user_id"], "price": f["price"], "click": f["click"],
           "genres__values": np.concatenate([np.asarray(g, np.int64) for g in f["genres"]]),
           "genres__offsets": np.concatenate([[0], np.cumsum([len(g) for g in f["genres"]])]).astype(np.int32)}
    inputs, targets = mm.sample_batch(raw, batch_size=64, schema=schema.select_by_name(["user_id", "genres", "price", "click"]), device="cpu")
    assert inputs["user_id"].shape[0] == 64 and torch.equal(inputs["genres__values"], b["genres__values"][: int(b["genres__offsets"][64])])
    assert torch.equal(targets, tb[:64])
    only = mm.sample_batch(from_df, include_targets=False)
    assert isinstance(only, dict) and "click" not in only
    with pytest.raises(ValueError, match="specify 'batch_size'"):
        mm.sample_batch(raw)
    no_schema = mm.Loader({"a": np.arange(10), "y": np.arange(10) % 2}, batch_size=4, label_names=["y"], shuffle=False, device="cpu")
    x, y = no_schema.peek()
    assert list(x) == ["a"] and y.tolist() == [0, 1, 0, 1]


def test_loader_errors(tmp_path):
    pq.write_table(pa.table({"a": pa.array([1, None, 3]), "s": pa.array(["x", "y", "z"])}), tmp_path / "bad.parquet")
    with pytest.raises(ValueError, match="nulls"):
        mm.Loader(str(tmp_path / "bad.parquet"), batch_size=2, device="cpu")
    pq.write_table(pa.table({"s": pa.array(["x", "y", "z"])}), tmp_path / "str.parquet")
    with pytest.raises(TypeError, match="not numeric"):
        mm.Loader(str(tmp_path / "str.parquet"), batch_size=2, device="cpu")
    with pytest.raises(ValueError, match="not in the dataset"):
        mm.Loader({"a": np.arange(4)}, batch_size=2, feature_columns=["b"], device="cpu")
    with pytest.raises(ValueError, match="positive"):
        mm.Loader({"a": np.arange(4)}, batch_size=0, device="cpu")
    big = mm.Loader({"id": np.array([0, 2**40], dtype=np.int64)}, batch_size=2, shuffle=False, device="cpu")
    assert big.peek()[0]["id"].dtype == torch.int64  # does not fit int32: kept wide


def test_packed_id_columns(dataset_dir):
    """`id_bytes` (`Model.id_bytes()`): scalar id columns travel as 1/2/3-byte unsigned ids; ragged ones stay int32."""
    d, f = dataset_dir
    loader = mm.Loader(str(d), batch_size=256, shuffle=False, device="cpu", id_bytes={"user_id": 2, "item_id": 3, "genres": 1})
    hb = next(iter(loader.host_batches()))
    assert hb.spec["user_id"] == ((256,), np.dtype(np.uint16)) and hb.spec["item_id"] == ((256, 3), np.dtype(np.uint8))
    assert hb.spec["genres__values"][1] == np.dtype(np.int32)
    assert np.array_equal(hb.columns["user_id"].numpy(), f["user_id"][:256].astype(np.uint16))
    packed = hb.columns["item_id"].numpy().astype(np.int64)
    assert np.array_equal(packed[:, 0] | packed[:, 1] << 8 | packed[:, 2] << 16, f["item_id"][:256])
    inputs, _ = next(iter(loader))
    assert inputs["user_id"].dtype == torch.uint16 and inputs["item_id"].shape == (256, 3)
