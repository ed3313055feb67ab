"""Sampled-softmax pieces and the V2 two-tower API: PopularityBasedSamplerV2 (outputs/sampling/popularity.py:24-195),
ContrastiveOutput over an item EmbeddingTable with logQ correction (outputs/contrastive.py:223-344), the v1 scorer's
sampled_softmax_mode (blocks/retrieval/base.py:274,313,431-453) and TwoTowerModelV2 (models/retrieval.py:409-486)."""
import numpy as np
import pytest
import torch

import models_b200 as mm
from models_b200 import datasets
from oracle import oracle

pytestmark = pytest.mark.gpu


def dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


@pytest.mark.parametrize("unique", [True, False])
def test_popularity_sampler_ids_and_distribution(device, unique):
    s = mm.PopularityBasedSamplerV2(max_id=5000, min_id=2, max_num_samples=300, unique=unique, seed=11)
    np.testing.assert_allclose(s.sampling_dist, oracle.log_uniform_probs(5000, 2, unique=unique, n_sampled=300), rtol=1e-6, atol=1e-9)
    ids = s.sample_ids(device).cpu().numpy()
    assert ids.shape == (300,) and ids.min() >= 2 and ids.max() < 5000
    if unique:
        assert len(set(ids.tolist())) == 300
    again = mm.PopularityBasedSamplerV2(max_id=5000, min_id=2, max_num_samples=300, unique=unique, seed=11).sample_ids(device)
    assert np.array_equal(ids, again.cpu().numpy())  # the seed makes the draw reproducible
    # the law: P(k) = (log(k+2) - log(k+1)) / log(R+1); check the head of the histogram of many non-unique draws
    big = mm.PopularityBasedSamplerV2(max_id=1000, min_id=0, max_num_samples=1000, unique=False, seed=3)
    draws = np.concatenate([big.sample_ids(device).cpu().numpy() for _ in range(200)])
    freq = np.bincount(draws, minlength=1000) / draws.size
    k = np.arange(8)
    want = (np.log(k + 2) - np.log(k + 1)) / np.log(1001.0)
    np.testing.assert_allclose(freq[:8], want, rtol=0.08)
    p = s.sampling_probs(dev(np.array([2, 3, 4999], dtype=np.int64), device)).cpu().numpy()
    np.testing.assert_allclose(p, s.sampling_dist[[2, 3, 4999]], rtol=1e-6)


@pytest.mark.parametrize("fused_loss", [False, True])
def test_contrastive_output_sampled_softmax_with_logq(device, fused_loss):
    """to_call = the item EmbeddingTable: positives are its rows at the target ids, negatives the rows of the sampled ids,
    logQ with the sampler's own probabilities: pos -= log(p_pos + 1e-16), neg -= log(p_neg + 1e-16), false negatives
    (a sampled id equal to the row's positive id) -> MIN_FLOAT, everything / T."""
    mm.set_seed(4)
    rng = np.random.default_rng(2)
    n_items, D, B, N = 3000, 64, 200, 128
    table = mm.EmbeddingTable(D, datasets._cat("item_id", n_items - 1))
    table.build(device)
    sampler = mm.PopularityBasedSamplerV2(max_id=n_items, min_id=1, max_num_samples=N, seed=5)
    out = mm.ContrastiveOutput(to_call=table, negative_samplers=sampler, logq_sampling_correction=True, logits_temperature=0.7)
    q = rng.standard_normal((B, D)).astype(np.float32)
    targets = np.minimum(rng.zipf(1.3, B), n_items - 1).astype(np.int64)  # popular ids: collisions with the negatives happen
    pred = out(dev(q, device), training=True, targets=dev(targets, device), fused_loss=fused_loss)
    nid = pred.negative_candidate_ids.cpu().numpy()
    assert nid.shape == (N,) and len(set(nid.tolist())) == N
    E = table.embeddings.cpu().numpy()
    logits, tgt = oracle.contrastive_logits(q, E[targets], E[nid], targets, nid, downscore=True,
                                            pos_prob=sampler.sampling_dist[targets], neg_prob=sampler.sampling_dist[nid], temperature=0.7)
    assert (targets[:, None] == nid[None, :]).any()  # the test does exercise the false-negative path
    if fused_loss:
        np.testing.assert_allclose(pred.predictions.cpu().numpy(), oracle.softmax_ce_stats(logits, np.zeros(B, dtype=np.int64)),
                                   rtol=3e-4, atol=2e-3)
    else:
        np.testing.assert_allclose(pred.predictions.cpu().numpy(), logits, rtol=3e-4, atol=2e-3)
        assert np.array_equal(pred.targets.cpu().numpy(), tgt)
    # inference: the whole catalog
    full = out(dev(q, device)).cpu().numpy()
    np.testing.assert_allclose(full, q @ E.T, rtol=2e-4, atol=2e-4)


def test_item_retrieval_scorer_sampled_softmax_mode(device):
    mm.set_seed(6)
    rng = np.random.default_rng(8)
    n_items, D, B = 1500, 32, 150
    table = mm.EmbeddingTable(D, datasets._cat("item_id", n_items - 1))
    table.build(device)
    sampler = mm.PopularityBasedSamplerV2(max_id=n_items, max_num_samples=64, seed=1)
    scorer = mm.ItemRetrievalScorer(samplers=[sampler], sampled_softmax_mode=True, item_table=table)
    q = rng.standard_normal((B, D)).astype(np.float32)
    targets = rng.integers(0, n_items, B).astype(np.int64)
    E = table.embeddings.cpu().numpy()
    np.testing.assert_allclose(scorer(dev(q, device)).cpu().numpy(), q @ E.T, rtol=2e-4, atol=2e-4)
    pred = scorer.call_outputs(dev(q, device), {}, targets=dev(targets, device))
    nid = pred.negative_candidate_ids.cpu().numpy()
    logits, _ = oracle.contrastive_logits(q, E[targets], E[nid], targets, nid, downscore=True)
    np.testing.assert_allclose(pred.predictions.cpu().numpy(), logits, rtol=2e-4, atol=5e-4)
    with pytest.raises(ValueError, match="item_table"):
        mm.ItemRetrievalScorer(sampled_softmax_mode=True)


def test_two_tower_model_v2(device):
    """Encoder towers (InputBlockV2: sorted concat of embeddings + continuous) + ContrastiveOutput, in-batch negatives."""
    from tests import helpers as H

    mm.set_seed(9)
    schema = datasets.movielens_1m_schema()
    qs, cs = schema.select_by_tag(mm.Tags.USER), schema.select_by_tag(mm.Tags.ITEM)
    query = mm.Encoder(qs, mm.MLPBlock([64, 32]))
    cand = mm.Encoder(cs, mm.MLPBlock([64, 32]), post="l2-norm")
    model = mm.TwoTowerModelV2(query, cand, logits_temperature=0.5)
    feats, _ = datasets.split_targets(schema, datasets.generate_batch(schema, 300, seed=21))
    batch = {k: dev(feats[k], device) for k in model.input_columns()}
    score = model(batch).cpu().numpy()
    pred = model(batch, training=True)

    def tower(enc, l2):
        tables, f2t = H.emb_tables(enc.inputs.embeddings)
        cont = enc.inputs.continuous.features if enc.inputs.continuous is not None else []
        sub = {k: v for k, v in feats.items() if any(k == c.name or k.startswith(c.name + "__") for c in enc.schema)}
        return oracle.tower_forward(sub, tables, f2t, cont, H.mlp_layers(enc.blocks[0]), combiner="mean", l2_normalize=l2)

    qo, co = tower(query, False), tower(cand, True)
    np.testing.assert_allclose(score, oracle.retrieval_scores(qo, co), rtol=2e-4, atol=2e-5)
    ids = feats["movieId"].astype(np.int64)
    logits, tgt = oracle.contrastive_logits(qo, co, co, ids, ids, downscore=True, temperature=0.5)
    np.testing.assert_allclose(pred.predictions.cpu().numpy(), logits, rtol=2e-4, atol=2e-3)
    assert np.array_equal(pred.targets.cpu().numpy(), tgt)
    with pytest.raises(AssertionError):
        mm.TwoTowerModelV2(mm.MLPBlock([8]), cand)
