"""Fused query x catalog scoring (mm_catalog_score) against the oracle: soft-max CE statistics and
top-k without materialising (B, N_I) — SURVEY §8(a) a14."""
import numpy as np
import pytest
import torch

import models_b200 as mm
from models_b200 import datasets, ops
from oracle import oracle

pytestmark = pytest.mark.gpu


def dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


@pytest.mark.parametrize("B,I,D", [(300, 5000, 64), (128, 128, 32), (1000, 70001, 64), (17, 300, 128), (256, 4096, 16)])
@pytest.mark.parametrize("use_bias", [False, True])
def test_catalog_lse_and_topk(device, B, I, D, use_bias):
    rng = np.random.default_rng(I + D)
    x = rng.standard_normal((B, D)).astype(np.float32)
    E = (rng.standard_normal((I, D)) * 0.5).astype(np.float32)
    bias = (rng.standard_normal(I) * 0.3).astype(np.float32) if use_bias else None
    targets = rng.integers(0, I, B).astype(np.int64)
    logits = oracle.catalog_logits(x, E, bias)
    e_split = ops.split_rows(dev(E, device))
    k = 10
    stats, scores, ids = ops.catalog_score(dev(x, device), e_split, I, bias=None if bias is None else dev(bias, device),
                                           targets=dev(targets, device), k=k)
    ref = oracle.softmax_ce_stats(logits, targets)
    tol = 3e-4 * max(1.0, np.sqrt(D / 64))
    np.testing.assert_allclose(stats.cpu().numpy(), ref, rtol=1e-4, atol=tol)
    rv, ri = oracle.topk(logits, k)
    np.testing.assert_allclose(scores.cpu().numpy(), rv, rtol=1e-4, atol=tol)
    got_ids = ids.cpu().numpy()
    # ids may differ from the oracle only where the scores tie within the GEMM tolerance
    same = got_ids == ri
    if not same.all():
        alt = np.take_along_axis(logits, got_ids, axis=1)
        assert np.allclose(alt[~same], rv[~same], atol=2 * tol)
    assert np.all((got_ids >= 0) & (got_ids < I))
    assert np.all(np.diff(scores.cpu().numpy(), axis=1) <= 0)  # descending
    for r in range(B):
        assert len(set(got_ids[r].tolist())) == k


def test_catalog_topk_tie_break_lower_id_first(device):
    """tf.math.top_k: equal values -> lower index first.  Integer-valued operands make the GEMM exact."""
    B, I, D = 130, 1000, 64
    x = np.zeros((B, D), dtype=np.float32)
    x[:, 0] = 1.0
    E = np.zeros((I, D), dtype=np.float32)
    E[:, 0] = np.arange(I) % 7  # every value 0..6 repeats ~143 times
    stats, scores, ids = ops.catalog_score(dev(x, device), ops.split_rows(dev(E, device)), I, k=12)
    rv, ri = oracle.topk(oracle.catalog_logits(x, E), 12)
    assert np.array_equal(scores.cpu().numpy(), rv) and np.array_equal(ids.cpu().numpy(), ri)
    assert ids.cpu().numpy()[0].tolist() == [6 + 7 * j for j in range(12)]
    ref = oracle.softmax_ce_stats(oracle.catalog_logits(x, E), np.zeros(B, dtype=np.int64))
    np.testing.assert_allclose(stats.cpu().numpy()[:, :2], ref[:, :2], rtol=1e-5, atol=1e-5)


def test_categorical_output_block(device):
    mm.set_seed(5)
    col = datasets._cat("item_id", 2999)
    table = mm.EmbeddingTable(64, col)
    out = mm.CategoricalOutput(table)
    rng = np.random.default_rng(9)
    x = rng.standard_normal((200, 64)).astype(np.float32)
    t = rng.integers(0, 3000, 200)
    full = out(dev(x, device)).cpu().numpy()
    E = table.embeddings.cpu().numpy()
    assert full.shape == (200, 3000)
    np.testing.assert_allclose(full, oracle.catalog_logits(x, E, np.zeros(3000, np.float32)), rtol=1e-4, atol=1e-4)
    stats = out.softmax_ce_stats(dev(x, device), dev(t, device)).cpu().numpy()
    np.testing.assert_allclose(stats, oracle.softmax_ce_stats(full, t), rtol=1e-4, atol=1e-4)
    loss = stats[:, 1] - stats[:, 2]  # categorical cross-entropy from logits
    assert np.all(loss >= -1e-5)
    scores, ids = out.top_k(dev(x, device), 5)
    rv, ri = oracle.topk(full, 5)
    np.testing.assert_allclose(scores.cpu().numpy(), rv, rtol=1e-4, atol=1e-4)
    with pytest.raises(ValueError, match="k must be"):
        out.top_k(dev(x, device), 64)
