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


def test_catalog_full_size_sampled_rows(device):
    """BASELINE configs[2] size: 10 M items x D 64.  The (B, 10 M) logits cannot be checked densely, so (a) the log-sum-exp
    and target logit of a sample of query rows are recomputed on the host in float64 over the WHOLE catalog (hash-initialised:
    any block of rows is regenerated on the host), and (b) the fused top-k of those rows is checked against the host's."""
    I, D, B = 10_000_000, 64, 384
    rng = np.random.default_rng(17)
    E = torch.empty((I, D), dtype=torch.float32, device=device)
    ops.init_uniform_hash(E, 77, -0.5, 0.5)
    x = (rng.standard_normal((B, D)) * 0.6).astype(np.float32)
    targets = rng.integers(0, I, B).astype(np.int64)
    targets[:3] = [0, I - 1, 127]  # tile / range edges
    stats, scores, ids = ops.catalog_score(dev(x, device), ops.split_rows(E), I, targets=dev(targets, device), k=8)
    stats, scores, ids = stats.cpu().numpy(), scores.cpu().numpy(), ids.cpu().numpy()
    rows = [0, 1, 2, 127, 128, 200, 383]
    xs = x[rows].astype(np.float64)
    m = np.full(len(rows), -np.inf)
    s = np.zeros(len(rows))
    best_v = np.full((len(rows), 8), -np.inf)
    best_i = np.full((len(rows), 8), -1, dtype=np.int64)
    step = 500_000
    for r0 in range(0, I, step):
        blk = oracle.hash_table_rows(np.arange(r0, min(I, r0 + step)), D, 77, -0.5, 0.5).astype(np.float64)
        lg = xs @ blk.T
        m_new = np.maximum(m, lg.max(axis=1))
        s = s * np.exp(m - m_new) + np.exp(lg - m_new[:, None]).sum(axis=1)
        m = m_new
        cand_v = np.concatenate([best_v, lg], axis=1)
        cand_i = np.concatenate([best_i, np.broadcast_to(np.arange(r0, r0 + lg.shape[1]), lg.shape)], axis=1)
        order = np.argsort(-cand_v, axis=1, kind="stable")[:, :8]
        best_v, best_i = np.take_along_axis(cand_v, order, 1), np.take_along_axis(cand_i, order, 1)
    lse = m + np.log(s)
    tl = np.einsum("bd,bd->b", xs, oracle.hash_table_rows(targets[rows], D, 77, -0.5, 0.5).astype(np.float64))
    np.testing.assert_allclose(stats[rows, 0], m, rtol=0, atol=5e-4)
    np.testing.assert_allclose(stats[rows, 1], lse, rtol=0, atol=5e-4)
    np.testing.assert_allclose(stats[rows, 2], tl, rtol=0, atol=5e-4)
    np.testing.assert_allclose(scores[rows], best_v, rtol=0, atol=5e-4)
    same = ids[rows] == best_i
    assert same.mean() > 0.9 and np.allclose(scores[rows][~same], best_v[~same], atol=1e-3)  # near-ties may swap


@pytest.mark.parametrize("B,N,D", [(100, 77, 32), (300, 1000, 64), (129, 256, 128), (1024, 1024, 128)])
@pytest.mark.parametrize("downscore,logq,temperature", [(True, False, 1.0), (True, True, 0.5), (False, False, 1.0), (False, True, 2.0)])
def test_inbatch_softmax_ce_equals_ce_of_the_materialised_logits(device, B, N, D, downscore, logq, temperature):
    """mm_inbatch_softmax_ce: [max, lse, positive logit] of exactly the logits mm_inbatch_scores produces."""
    rng = np.random.default_rng(B + N)
    q = rng.standard_normal((B, D)).astype(np.float32) * 0.7
    pos = rng.standard_normal((B, D)).astype(np.float32) * 0.7
    neg = pos if N == B else (rng.standard_normal((N, D)).astype(np.float32) * 0.7)
    pos_ids = rng.integers(0, max(4, N // 3), B).astype(np.int64)       # many duplicates -> many accidental hits
    neg_ids = pos_ids if N == B else rng.integers(0, max(4, N // 3), N).astype(np.int64)
    pp = rng.uniform(1e-4, 0.2, B).astype(np.float32) if logq else None
    npb = (pp if N == B else rng.uniform(1e-4, 0.2, N).astype(np.float32)) if logq else None
    logits, _ = oracle.contrastive_logits(q, pos, neg, pos_ids, neg_ids, downscore=downscore, pos_prob=pp, neg_prob=npb,
                                          temperature=temperature)
    ref = oracle.softmax_ce_stats(logits, np.zeros(B, dtype=np.int64))
    d = lambda a: None if a is None else dev(a, device)
    nt = d(pos) if N == B else d(neg)
    pt = nt if N == B else d(pos)
    stats = ops.inbatch_softmax_ce(d(q), pt, nt, pos_ids=d(pos_ids), neg_ids=d(neg_ids), downscore=downscore,
                                   pos_prob=d(pp), neg_prob=d(npb), temperature=temperature).cpu().numpy()
    tol = 4e-4 * max(1.0, np.sqrt(D / 64)) / min(1.0, temperature)
    np.testing.assert_allclose(stats, ref, rtol=2e-4, atol=tol)
    loss = stats[:, 1] - stats[:, 2]
    assert np.all(loss >= -1e-4)


def test_two_tower_fused_loss_matches_logits_path(device):
    """model(batch, training=True, fused_loss=True) returns the cross-entropy statistics of the logits that
    model(batch, training=True) returns — eager and through the CUDA-graph runtime."""
    mm.set_seed(7)
    schema = datasets.movielens_1m_schema()
    model = mm.TwoTowerModel(schema, query_tower=mm.MLPBlock([128, 64]), logits_temperature=0.8)
    batch = datasets.generate_batch(schema, 512, seed=4)
    feats, _ = datasets.split_targets(schema, batch)
    cols = model.input_columns()
    dbatch = {k: dev(feats[k], device) for k in cols}
    logits = model(dbatch, training=True).predictions.cpu().numpy()
    ref = oracle.softmax_ce_stats(logits, np.zeros(512, dtype=np.int64))
    pred = model(dbatch, training=True, fused_loss=True)
    assert pred.targets is None and tuple(pred.predictions.shape) == (512, 3)
    np.testing.assert_allclose(pred.predictions.cpu().numpy(), ref, rtol=2e-4, atol=5e-4)
    hb = mm.HostBatch.like(feats, cols)
    cf = model.compile(hb, training=True, fused_loss=True)
    np.testing.assert_allclose(cf(hb).numpy(), ref, rtol=2e-4, atol=5e-4)
