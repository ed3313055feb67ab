"""Constructor options of the hot-path blocks that round 1 rejected: MLPBlock(normalization=...) (BatchNormalization at
inference, folded into the next Dense), CrossBlock(low_rank_dim=...), and weight reassignment after graph capture."""
import numpy as np
import pytest
import torch

import models_b200 as mm
from models_b200 import blocks, datasets
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _randomise_bn(mlp, rng):
    for l in mlp.layers:
        if isinstance(l, blocks.BatchNormalization):
            n = l.gamma.numel()
            l.set_weights(gamma=rng.uniform(0.5, 1.5, n).astype(np.float32), beta=rng.normal(0, 0.3, n).astype(np.float32),
                          moving_mean=rng.normal(0, 0.5, n).astype(np.float32), moving_variance=rng.uniform(0.2, 2.0, n).astype(np.float32))


@pytest.mark.parametrize("engine", ["auto", "fp32"])
@pytest.mark.parametrize("dims", [[128, 64], [200, 40, 8], [16]])
def test_mlp_block_batch_norm_matches_keras_inference(device, engine, dims):
    rng = np.random.default_rng(3)
    mm.set_seed(11)
    x = rng.standard_normal((515, 37)).astype(np.float32)
    mlp = mm.MLPBlock(dims, normalization="batch_norm")
    assert [type(l).__name__ for l in mlp.layers] == ["_Dense", "BatchNormalization"] * len(dims)  # blocks/mlp.py:131-135
    mlp.build_from_width(37, device)
    _randomise_bn(mlp, rng)
    blocks.set_dense_engine(engine)
    try:
        got = mlp(torch.from_numpy(x).to(device)).cpu().numpy()
    finally:
        blocks.set_dense_engine("auto")
    ref = oracle.mlp(x, H.mlp_layers(mlp))
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-4)
    # fresh variables: gamma 1, beta 0, mean 0, var 1 -> y = x / sqrt(1 + 1e-3)
    fresh = mm.MLPBlock([8], activation="linear", normalization=blocks.BatchNormalization())
    plain = mm.MLPBlock([8], activation="linear")
    fresh.build_from_width(37, device), plain.build_from_width(37, device)
    plain.dense_layers[0].set_weights(fresh.dense_layers[0].kernel, fresh.dense_layers[0].bias)
    xt = torch.from_numpy(x).to(device)
    np.testing.assert_allclose(fresh(xt).cpu().numpy(), plain(xt).cpu().numpy() / np.sqrt(1.001), rtol=1e-5, atol=1e-6)
    with pytest.raises(NotImplementedError):
        mlp(xt, training=True)
    with pytest.raises(ValueError, match="Normalization needs to be an instance"):
        mm.MLPBlock([4], normalization="layer_norm")


def test_dlrm_model_with_batch_norm_towers(device):
    """BatchNormalization inside both towers: the bottom tower's last one runs as mm_scale_shift, the top tower's
    are folded — the last into the output Dense(1).  Eager, graph, and after reassigning the statistics."""
    rng = np.random.default_rng(5)
    mm.set_seed(3)
    schema = datasets.criteo_schema({k: min(v, 500) for k, v in datasets.CRITEO_MAX.items()})
    model = mm.DLRMModel(schema, embedding_dim=64, bottom_block=mm.MLPBlock([128, 64], normalization="batch_norm"),
                         top_block=mm.MLPBlock([128, 64, 32], normalization="batch_norm"))
    model.build(device)
    _randomise_bn(model.body.bottom_block, rng)
    _randomise_bn(model.body.top_block, rng)
    feats, _ = datasets.split_targets(schema, datasets.generate_batch(schema, 700, seed=8, index_law="uniform"))
    ref = H.oracle_dlrm(model, feats)
    got = model(H.device_batch(feats, device)).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-5)
    hb = mm.HostBatch.like(feats, model.input_columns(), id_bytes=model.id_bytes())
    cf = model.compile(hb)
    np.testing.assert_allclose(cf(hb).numpy(), ref, rtol=2e-4, atol=2e-5)
    # variables reassigned AFTER the capture -> the compiled forward re-captures instead of replaying stale pointers
    _randomise_bn(model.body.top_block, rng)
    model.prediction.to_call.set_weights(rng.standard_normal((32, 1)).astype(np.float32) * 0.3, np.array([0.25], dtype=np.float32))
    ref2 = H.oracle_dlrm(model, feats)
    assert np.abs(ref2 - ref).max() > 1e-3
    np.testing.assert_allclose(cf(hb).numpy(), ref2, rtol=2e-4, atol=2e-5)
    pf = model.pipeline(hb, depth=2)
    model.body.top_block.dense_layers[0].set_weights(model.body.top_block.dense_layers[0].kernel * 0.5,
                                                     model.body.top_block.dense_layers[0].bias)
    ref3 = H.oracle_dlrm(model, feats)
    np.testing.assert_allclose(pf.result(pf.submit(hb)).numpy(), ref3, rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("engine", ["auto", "fp32"])
@pytest.mark.parametrize("d,r,depth", [(64, 16, 3), (1037, 128, 2), (50, 8, 1)])
def test_cross_block_low_rank(device, engine, d, r, depth):
    """CrossBlock(low_rank_dim=r): x_{l+1} = x0 * (dense(dense_u(x_l))) + x_l with dense_u (d, r) without bias and dense
    (r, d) with bias (blocks/cross.py:99, blocks/mlp.py:389-396)."""
    rng = np.random.default_rng(9)
    mm.set_seed(2)
    x = (rng.standard_normal((300, d)) * 0.5).astype(np.float32)
    cb = mm.CrossBlock(depth, low_rank_dim=r)
    blocks.set_dense_engine(engine)
    try:
        got = cb(torch.from_numpy(x).to(device)).cpu().numpy()
    finally:
        blocks.set_dense_engine("auto")
    xl = x.copy()
    for l in cb.cross_layers:
        assert tuple(l.dense_u.kernel.shape) == (d, r) and l.dense_u.bias is None and tuple(l.dense.kernel.shape) == (r, d)
        proj = (xl @ l.dense_u.kernel.cpu().numpy()) @ l.dense.kernel.cpu().numpy() + l.dense.bias.cpu().numpy()
        xl = x * proj + xl
    np.testing.assert_allclose(got, xl, rtol=3e-4, atol=3e-4)
