"""Host-side logic of the FM heads that needs no GPU: the row layout of the wide kernel (sorted feature names, one block of
int_domain.max + 1 rows per categorical feature, one row per continuous feature — CategoryEncoding + ConcatFeatures,
blocks/interaction.py:307-316) and the constructor checks of FMBlock / DeepFMModel."""
import pytest

import models_b200 as mm
from models_b200.schema import ColumnSchema, Schema


def _schema():
    cat = lambda n, mx: ColumnSchema(n, tags=("categorical",), dtype="int64", properties={"domain": {"min": 0, "max": mx, "name": n}})
    return Schema([cat("item_id", 99), cat("item_category", 9), cat("user_id", 49), ColumnSchema("user_age", tags=("continuous",), dtype="float32"),
                   ColumnSchema("click", tags=("target", "binary_classification"), dtype="int64")])


def test_wide_kernel_layout_follows_sorted_feature_names():
    fm = mm.FMBlock(_schema(), factors_dim=16)
    # sorted: item_category (10 rows), item_id (100), user_age (1), user_id (50)
    assert fm.wide_offsets == {"item_category": 0, "item_id": 10, "user_age": 110, "user_id": 111}
    assert fm.wide_width == 161 and fm.dim == 16
    assert fm.cat_names == ["item_id", "item_category", "user_id"] and fm.cont_names == ["user_age"]


def test_deepfm_constructor_mirrors_the_reference():
    schema = _schema()
    model = mm.DeepFMModel(schema, embedding_dim=16, deep_block=mm.MLPBlock([16]), prediction_tasks=mm.BinaryOutput("click"))
    body = model.body
    assert body.fm.embeddings is body.input_block.embeddings  # one set of tables for the FM and the deep tower
    assert [l.units for l in body.deep.dense_layers] == [16]
    assert [(l.units, l.activation, l.use_bias) for l in body.deep_logit.dense_layers] == [(1, "linear", True)]
    assert model.body_width() == 1 and model.prediction.to_call.activation == "sigmoid"
    assert [l.units for l in mm.DeepFMModel(schema, embedding_dim=8).body.deep.dense_layers] == [64]
    with pytest.raises(ValueError, match="embedding_dim"):
        mm.DeepFMModel(schema)
    with pytest.raises(NotImplementedError, match="wide"):
        mm.DeepFMModel(schema, embedding_dim=8, wide_logit_block=mm.MLPBlock([1]))
    only_cont = Schema([c for c in schema if c.name in ("user_age", "click")])
    with pytest.raises(ValueError, match="categorical"):
        mm.FMBlock(only_cont, factors_dim=8)
