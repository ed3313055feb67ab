"""Factorization-machine heads (blocks/interaction.py:205-332, DeepFMModel models/ranking.py:171-279) on the GPU against the
oracle restatement.  The reference's own tests check shapes only (tests/unit/tf/blocks/test_interactions.py:25-48,
tests/unit/tf/models/test_ranking.py:210-238); TensorFlow cannot run here and the torch backend has no FM block, so the
numbers are pinned to oracle/oracle.py (restated from the source, hand-checked in tests/test_oracle.py)."""
import numpy as np
import pytest
import torch

import models_b200 as mm
from models_b200 import datasets
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _schema(cap=300, cont=True):
    s = datasets.criteo_schema({k: min(v, cap) for k, v in datasets.CRITEO_MAX.items()})
    if not cont:
        s = s.select_by_name([c.name for c in s if not c.has_tag(mm.Tags.CONTINUOUS)])
    return s


def test_fm_pairwise_interaction(device):
    x = torch.rand((100, 10, 64), device=device)
    out = mm.FMPairwiseInteraction()(x)
    assert list(out.shape) == [100, 64]
    np.testing.assert_allclose(out.cpu().numpy(), oracle.fm_pairwise(x.cpu().numpy()), rtol=1e-5, atol=1e-6)
    with pytest.raises(AssertionError, match="3-D"):
        mm.FMPairwiseInteraction()(x[0])


@pytest.mark.parametrize("cont", [True, False])
def test_fm_block_matches_oracle(device, cont):
    mm.set_seed(4)
    schema = _schema(cont=cont)
    fm = mm.FMBlock(schema, factors_dim=32)
    batch = datasets.generate_batch(schema, 333, seed=5, index_law="uniform")
    feats, _ = datasets.split_targets(schema, batch)
    out = fm(H.device_batch(feats, device))
    assert tuple(out.shape) == (333, 1)
    fm.wide.set_weights(np.random.default_rng(1).normal(size=(fm.wide_width, 1)).astype(np.float32) * 0.3, np.array([0.25], np.float32))
    out = fm(H.device_batch(feats, device)).cpu().numpy()
    tables, f2t = H.emb_tables(fm.embeddings)
    card = {f: fm.embeddings.feature_to_table[f].input_dim for f in fm.cat_names}
    ref = oracle.fm_block(feats, tables, f2t, fm.cont_names, card, H.to_numpy(fm.wide.kernel), H.to_numpy(fm.wide.bias))
    assert H.rel_err(out, ref) < 1e-5
    with pytest.raises(ValueError, match="factors_dim"):
        mm.FMBlock(schema)
    with pytest.raises(NotImplementedError, match="wide"):
        mm.FMBlock(schema, factors_dim=8, wide_logit_block=mm.MLPBlock([1]))


@pytest.mark.parametrize("cont", [True, False])
@pytest.mark.parametrize("index_dtype", [np.int32, np.int64])
def test_deepfm_model_matches_oracle(device, cont, index_dtype):
    """tests/unit/tf/models/test_ranking.py:210-238 (categorical only / categorical + continuous), with numbers."""
    mm.set_seed(8)
    schema = _schema(cont=cont)
    model = mm.DeepFMModel(schema, embedding_dim=16, deep_block=mm.MLPBlock([16]))
    model.build(device)
    rng = np.random.default_rng(2)
    fm = model.body.fm
    fm.wide.set_weights(rng.normal(size=(fm.wide_width, 1)).astype(np.float32) * 0.2, np.array([-0.1], np.float32))
    model.prediction.to_call.set_weights(np.array([[0.7]], np.float32), np.array([0.05], np.float32))
    batch = datasets.generate_batch(schema, 1000, seed=6, index_law="uniform", index_dtype=index_dtype)
    feats, _ = datasets.split_targets(schema, batch)
    out = model(H.device_batch(feats, device))
    assert out.shape == (1000, 1) and out.dtype == torch.float32
    got = out.cpu().numpy()
    assert np.all((got > 0) & (got < 1))
    assert H.rel_err(got, H.oracle_deepfm(model, feats)) < 2e-4
    # packed ids and the CUDA-graph runtime give the same numbers
    hb = mm.HostBatch.like(feats, model.input_columns(), id_bytes=model.id_bytes())
    cf = model.compile(hb)
    np.testing.assert_allclose(cf(hb).numpy(), got, rtol=1e-6, atol=1e-7)


def test_deepfm_defaults_checks_and_checkpoint(device, tmp_path):
    schema = _schema(cap=50)
    with pytest.raises(ValueError, match="embedding_dim"):
        mm.DeepFMModel(schema)
    with pytest.raises(ValueError, match="needs to be 1"):
        mm.DeepFMModel(schema, embedding_dim=8, deep_logit_block=mm.MLPBlock([2]))
    model = mm.DeepFMModel(schema, embedding_dim=8)  # deep_block defaults to MLPBlock([64])
    assert [l.units for l in model.body.deep.dense_layers] == [64] and model.body.deep_logit.dense_layers[0].activation == "linear"
    batch = datasets.generate_batch(schema, 64, seed=1, index_law="uniform")
    feats, _ = datasets.split_targets(schema, batch)
    out = model(H.device_batch(feats, device)).cpu().numpy()
    model.save(tmp_path / "export")
    loaded = mm.Model.load(tmp_path / "export")
    np.testing.assert_array_equal(loaded(H.device_batch(feats, device)).cpu().numpy(), out)
    bad = dict(feats)
    bad["C1"] = bad["C1"].copy()
    bad["C1"][0] = 10**6
    with pytest.raises(IndexError):
        model(H.device_batch(bad, device))
    from models_b200.schema import ColumnSchema, Schema

    lists = Schema([ColumnSchema("g", tags=("categorical",), dtype="int64", is_list=True, is_ragged=True,
                                 properties={"domain": {"min": 0, "max": 9, "name": "g"}, "value_count": {"min": 1, "max": 3}}),
                    ColumnSchema("click", tags=("target", "binary_classification"), dtype="int64")])
    with pytest.raises(NotImplementedError, match="multi-hot"):
        mm.DeepFMModel(lists, embedding_dim=8)
