import pytest

from models_b200 import datasets
from models_b200.schema import ColumnSchema, Schema, Tags

PBTXT = '''
feature {
  name: "item_id"
  type: INT
  int_domain { name: "item" max: 99 is_categorical: true }
  annotation { tag: "categorical" tag: "item_id" extra_metadata { type_url: "x" value: "\\n\\r\\007{}" } }
}
feature { name: "genres" value_count { min: 1 max: 5 } type: INT int_domain { max: 18 is_categorical: true }
          annotation { tag: "categorical" tag: "item" } }
feature { name: "price" type: FLOAT float_domain { min: 0.0 max: 10.5 } annotation { tag: "continuous" } }
feature { name: "click" type: INT annotation { tag: "binary_classification" tag: "target" } }
'''


def test_text_proto_reader():
    s = Schema.from_proto_text(PBTXT)
    assert s.column_names == ["item_id", "genres", "price", "click"]
    assert s["item_id"].int_domain.max == 99 and s["item_id"].int_domain.name == "item"
    assert {"item_id", "item", "id", "categorical"} <= set(s["item_id"].tags)
    assert s["genres"].is_list and s["genres"].is_ragged and s["genres"].value_count.max == 5
    assert s["price"].float_domain.max == 10.5 and s["price"].int_domain is None
    assert s.select_by_tag(Tags.CATEGORICAL).column_names == ["item_id", "genres"]
    assert s.select_by_tag([Tags.CONTINUOUS, Tags.TARGET]).column_names == ["price", "click"]
    assert s.excluding_by_tag(Tags.TARGET).column_names == ["item_id", "genres", "price"]
    assert s.select_by_tag(Tags.ITEM_ID).first.name == "item_id"
    assert len(s.select_by_name(["price", "nope"])) == 1
    assert (s.select_by_tag(Tags.TARGET) + s.select_by_tag(Tags.CONTINUOUS)).column_names == ["click", "price"]


def test_json_reader():
    js = '{"feature": [{"name": "a", "type": "INT", "intDomain": {"name": "a", "max": "7", "isCategorical": true},' \
         ' "annotation": {"tag": ["categorical", "user_id"]}}]}'
    s = Schema.from_json(js)
    assert s["a"].int_domain.max == 7 and s["a"].has_tag(Tags.USER)


def test_builtin_shapes_match_survey():
    c = datasets.criteo_schema()
    cat = c.select_by_tag(Tags.CATEGORICAL)
    assert len(cat) == 26 and len(c.select_by_tag(Tags.CONTINUOUS)) == 13
    assert sum(col.int_domain.max + 1 for col in cat) == 45_621_194
    tb = datasets.criteo_tb_schema()
    assert sum(col.int_domain.max + 1 for col in tb.select_by_tag(Tags.CATEGORICAL)) == sum(datasets.CRITEO_TB_ROWS)
    m = datasets.movielens_1m_schema()
    assert m.select_by_tag(Tags.USER).column_names[0] == "userId"
    assert m["genres"].is_ragged


def test_duplicate_and_bad_columns():
    with pytest.raises(ValueError, match="duplicate"):
        Schema([ColumnSchema("a"), ColumnSchema("a")])
    with pytest.raises(ValueError, match="is_ragged requires is_list"):
        ColumnSchema("a", is_ragged=True)


def test_generate_batch_follows_reference_generator():
    m = datasets.movielens_1m_schema()
    b = datasets.generate_batch(m, 500, seed=3)
    assert b["userId"].min() >= 1 and b["userId"].max() <= 6040  # clip(lognormal, 1, max)
    assert b["genres__offsets"].shape == (501,) and b["genres__offsets"][0] == 0
    assert b["genres__values"].shape[0] == b["genres__offsets"][-1]
    lens = b["genres__offsets"][1:] - b["genres__offsets"][:-1]
    assert lens.min() >= 1 and lens.max() <= 4
    assert 0 <= b["TE_age_rating"].min() and b["TE_age_rating"].max() < 1
    assert set(b["rating_binary"].tolist()) <= {0, 1}
    u = datasets.generate_batch(datasets.criteo_schema(), 100, index_law="uniform")
    assert u["C6"].max() <= 3 and u["C1"].dtype.name == "int32"
    with pytest.raises(ValueError):
        datasets.generate_batch(m, 4, index_law="bogus")


def test_schema_json_writer_round_trips(tmp_path):
    """Schema.to_json writes the tensorflow-metadata JSON the reference stores under `.merlin/`
    (merlin/models/io.py:26-55); reading it back gives the same columns, tags, domains and list-ness."""
    import models_b200 as mm

    for schema in (datasets.criteo_schema(), datasets.movielens_1m_schema()):
        p = tmp_path / "input_schema.json"
        schema.to_json(p)
        back = mm.Schema.load(str(p))
        assert back.column_names == schema.column_names
        for a, b in zip(schema, back):
            assert set(a.tags) == set(b.tags), a.name
            assert a.is_list == b.is_list and a.is_ragged == b.is_ragged, a.name
            assert (a.int_domain is None) == (b.int_domain is None), a.name
            if a.int_domain is not None:
                assert (a.int_domain.max, a.int_domain.name) == (b.int_domain.max, b.int_domain.name)
    mm.save_merlin_metadata(tmp_path, datasets.criteo_schema(), None)
    inp, out = mm.load_merlin_metadata(tmp_path)
    assert inp.column_names == datasets.criteo_schema().column_names and out is None


def test_inferred_embedding_dims_match_the_reference_utility():
    """mm.infer_embedding_dim (host logic of Embeddings / InputBlockV2 / DCNModel defaults) against the outputs of the
    reference's own merlin/models/utils/schema_utils.py:169-207 executed in the build container
    (tests/golden/ref_torch_embedding_dims.npz)."""
    from pathlib import Path

    import numpy as np

    import models_b200 as mm

    z = np.load(Path(__file__).parent / "golden" / "ref_torch_embedding_dims.npz")
    cols = [ColumnSchema(f"f{i}", tags=(Tags.CATEGORICAL,), dtype="int64", properties={"domain": {"min": 0, "max": int(m), "name": f"f{i}"}})
            for i, m in enumerate(z["max_id"])]
    assert [mm.infer_embedding_dim(c) for c in cols] == z["dims_default"].tolist()
    assert [mm.infer_embedding_dim(c, multiplier=3.0, ensure_multiple_of_8=False) for c in cols] == z["dims_mult3_plain"].tolist()
    # the bundled Criteo schema: default DCN input width = 1024 + 13
    schema = datasets.criteo_schema()
    cat = schema.select_by_tag(Tags.CATEGORICAL)
    assert sum(mm.infer_embedding_dim(c) for c in cat) == 1024
