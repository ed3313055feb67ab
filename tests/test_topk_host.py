"""Host-side pieces of the top-k evaluation step (no GPU): ranking metrics against the NumPy restatement of
merlin/models/tf/metrics/topk.py:48-190 and its hand-computed known answers; ragged-aware row selection."""
import numpy as np
import torch

from models_b200 import topk
from oracle import oracle


def test_metric_known_answers():
    y = np.array([[0, 1, 0, 0], [0, 0, 0, 0], [1, 0, 1, 0]], np.float32)
    rel = np.array([1, 1, 2], np.float32)
    np.testing.assert_allclose(oracle.recall_at(y, rel, 3), [1, 0, 1])
    np.testing.assert_allclose(oracle.precision_at(y, rel, 2), [0.5, 0, 0.5])
    np.testing.assert_allclose(oracle.mrr_at(y, rel, 3), [0.5, 0, 1])
    np.testing.assert_allclose(oracle.ndcg_at(y, rel, 3), [1 / np.log2(3), 0, 1.5 / (1 + 1 / np.log2(3))], rtol=1e-6)
    np.testing.assert_allclose(oracle.average_precision_at(y, rel, 3), [0.5, 0, (1 + 2 / 3) / 2], rtol=1e-6)


def test_metrics_match_oracle_on_random_relevance():
    rng = np.random.default_rng(0)
    y = (rng.random((200, 20)) < 0.15).astype(np.float32)
    rel = np.maximum(y.sum(-1), rng.integers(0, 3, 200)).astype(np.float32)  # relevant items beyond the top 20 exist
    ty, tr = torch.from_numpy(y), torch.from_numpy(rel)
    for k in (1, 5, 10, 20):
        pairs = [(topk.RecallAt(k), oracle.recall_at), (topk.PrecisionAt(k), oracle.precision_at),
                 (topk.AvgPrecisionAt(k), oracle.average_precision_at), (topk.NDCGAt(k), oracle.ndcg_at),
                 (topk.MRRAt(k), oracle.mrr_at)]
        for metric, ref in pairs:
            np.testing.assert_allclose(metric(ty, tr).numpy(), ref(y, rel, k), rtol=1e-5, atol=1e-6, err_msg=metric.label)
    assert topk.RecallAt(10).label == "recall_at_10" and topk.NDCGAt(5).label == "ndcg_at_5"


def test_row_selection_keeps_ragged_features_consistent():
    d = {"item_id": np.array([7, 3, 7, 9, 3]), "price": np.array([1.0, 2.0, 3.0, 4.0, 5.0], np.float32),
         "genres__values": np.array([1, 2, 3, 4, 5, 6, 7]), "genres__offsets": np.array([0, 2, 2, 5, 6, 7], np.int32)}
    u = topk.unique_rows_by_features(d, "item_id")
    assert u["item_id"].tolist() == [7, 3, 9] and u["price"].tolist() == [1.0, 2.0, 4.0]  # first occurrences, original order
    assert u["genres__offsets"].tolist() == [0, 2, 2, 3] and u["genres__values"].tolist() == [1, 2, 6]
    s = topk.slice_rows(d, 1, 4)
    assert s["genres__offsets"].tolist() == [0, 0, 3, 4] and s["genres__values"].tolist() == [3, 4, 5, 6]
    t = topk.take_rows(d, np.array([4, 0, 2]))
    assert t["genres__values"].tolist() == [7, 1, 2, 3, 4, 5] and t["genres__offsets"].tolist() == [0, 1, 3, 6]


def test_oracle_topk_index_orders_ties_like_tf():
    q = np.array([[1.0, 0.0]], np.float32)
    v = np.array([[2.0, 0], [5.0, 0], [2.0, 1], [-1.0, 0]], np.float32)
    s, ids = oracle.topk_index(q, v, np.array([10, 11, 12, 13]), 3)
    assert s.tolist() == [[5.0, 2.0, 2.0]] and ids.tolist() == [[11, 10, 12]]  # ties: lower index first
    assert oracle.topk_targets(np.array([12]), ids).tolist() == [[0.0, 0.0, 1.0]]
