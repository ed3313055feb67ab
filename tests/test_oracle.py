"""The CPU oracle against (a) the known-answer invariants the reference's own tests hold for this
path and (b) the committed golden fixtures (tests/golden/*.npz, generator scripts beside them)."""
import math
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle

GOLDEN = Path(__file__).parent / "golden"


def test_min_float_constant():
    assert abs(oracle.MIN_FLOAT - (-655.04)) < 1e-9  # merlin/models/utils/constants.py:19


def test_hash_uniform_is_deterministic_and_in_range():
    a = oracle.hash_uniform(np.arange(10000), 42)
    b = oracle.hash_uniform(np.arange(10000), 42)
    c = oracle.hash_uniform(np.arange(10000), 43)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.dtype == np.float32 and a.min() >= -0.05 and a.max() < 0.05
    rows = oracle.hash_table_rows([0, 5, 9_999_999], 64, 7)
    assert rows.shape == (3, 64)
    assert np.array_equal(rows[1], oracle.hash_uniform(5 * 64 + np.arange(64), 7))


@pytest.mark.parametrize("card,expected", [(10, 8), (100, 8), (1000, 16), (10_000_000, 120), (4, 8), (35, 8)])
def test_infer_embedding_dim(card, expected):
    # tests/unit/tf/inputs/test_embedding.py:485-553 semantics: ceil(card^0.25 * 2) rounded up to x8
    assert oracle.infer_embedding_dim(card) == expected
    assert oracle.infer_embedding_dim(card, ensure_multiple_of_8=False) == int(math.ceil(card ** 0.25 * 2))


def test_dot_interaction_order_and_width():
    """tests/unit/tf/blocks/test_dlrm.py:36-38 (width F(F-1)/2) + torch/blocks/dlrm.py:65-74
    (triu_indices(F, F, offset=1) enumeration)."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 5, 8)).astype(np.float32)
    out = oracle.dot_interaction(x)
    assert out.shape == (4, 10)
    k = 0
    for i in range(5):
        for j in range(i + 1, 5):
            np.testing.assert_allclose(out[:, k], (x[:, i] * x[:, j]).sum(-1), rtol=1e-5)
            k += 1
    assert oracle.dot_interaction(x, self_interaction=True).shape == (4, 15)


def test_embedding_bag_semantics():
    t = np.arange(20, dtype=np.float32).reshape(5, 4)
    values = np.array([1, 2, -1, 3, 3, 4], dtype=np.int64)
    offsets = np.array([0, 2, 2, 3, 6], dtype=np.int32)  # bags: [1,2], [], [-1], [3,3,4]
    s = oracle.embedding_bag(t, values, offsets, "sum")
    assert np.array_equal(s[0], t[1] + t[2]) and not s[1].any() and not s[2].any()
    m = oracle.embedding_bag(t, values, offsets, "mean")
    np.testing.assert_allclose(m[3], (t[3] * 2 + t[4]) / 3, rtol=1e-6)
    q = oracle.embedding_bag(t, values, offsets, "sqrtn")
    np.testing.assert_allclose(q[0], (t[1] + t[2]) / np.sqrt(2), rtol=1e-6)
    with pytest.raises(IndexError):
        oracle.embedding_lookup(t, np.array([5]))


def test_sorted_key_aggregations():
    d = {"b": np.ones((2, 1)), "C10": np.full((2, 1), 2.0), "C2": np.full((2, 1), 3.0), "bottom_block": np.zeros((2, 1))}
    # ASCII order: 'C10' < 'C2' < 'b' < 'bottom_block'
    assert oracle.concat_features(d).tolist() == [[2.0, 3.0, 1.0, 0.0]] * 2
    st = oracle.stack_features({k: v for k, v in d.items()}, axis=1)
    assert st.shape == (2, 4, 1) and st[0, :, 0].tolist() == [2.0, 3.0, 1.0, 0.0]


def test_contrastive_false_negative_diagonal():
    """tests/unit/tf/outputs/test_contrastive.py:173-206."""
    rng = np.random.default_rng(1)
    q = rng.standard_normal((10, 4)).astype(np.float32)
    it = rng.standard_normal((10, 4)).astype(np.float32)
    ids = np.arange(10)
    out, targets = oracle.contrastive_logits(q, it, it, ids, ids)
    assert out.shape == (10, 11)
    assert np.all(np.diag(out[:, 1:]) == np.float32(oracle.MIN_FLOAT))
    off = out[:, 1:][~np.eye(10, dtype=bool)]
    assert np.all(off != np.float32(oracle.MIN_FLOAT))
    assert targets[:, 0].all() and not targets[:, 1:].any()
    assert oracle.retrieval_scores(q, it).shape == (10, 1)  # :209-223


def test_cross_layer_changes_input_and_keeps_shape():
    """tests/unit/tf/blocks/test_cross.py:25-59."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((6, 10)).astype(np.float32)
    layers = [{"kernel": (rng.standard_normal((10, 10)) * 0.1).astype(np.float32), "bias": np.zeros(10, np.float32)}
              for _ in range(3)]
    out = oracle.cross_layers(x, layers)
    assert out.shape == x.shape and not np.allclose(out, x)
    one = oracle.cross_layers(x, layers[:1])
    np.testing.assert_allclose(one, x * (x @ layers[0]["kernel"]) + x, rtol=1e-5, atol=1e-6)


def test_topk_ties_lower_index_first_and_lse():
    logits = np.array([[1.0, 3.0, 3.0, 2.0], [0.0, 0.0, 0.0, 0.0]], dtype=np.float32)
    v, i = oracle.topk(logits, 2)
    assert i.tolist() == [[1, 2], [0, 1]] and v.tolist() == [[3.0, 3.0], [0.0, 0.0]]
    st = oracle.softmax_ce_stats(logits, np.array([3, 0]))
    np.testing.assert_allclose(st[1], [0.0, np.log(4.0), 0.0], rtol=1e-6)


def test_oracle_torch_port_matches_numpy_oracle():
    """The multi-threaded CPU baseline (oracle/oracle_torch.py) computes the same function."""
    import torch

    from oracle import oracle_torch

    rng = np.random.default_rng(3)
    T, D, B = 5, 8, 64
    names = ["C1", "C10", "C2", "a9", "Z"]
    tables = {n: rng.standard_normal((30, D)).astype(np.float32) for n in names}
    f2t = {n: n for n in names}
    batch = {n: rng.integers(0, 30, B).astype(np.int32) for n in names}
    cont = ["I1", "I10", "I2"]
    for c in cont:
        batch[c] = rng.random(B).astype(np.float32)

    def layers(dims, w):
        out = []
        for d in dims:
            out.append({"kernel": (rng.standard_normal((w, d)) * 0.2).astype(np.float32),
                        "bias": rng.standard_normal(d).astype(np.float32), "activation": "relu"})
            w = d
        return out

    bottom, top = layers([16, D], 3), layers([16, 4], D + 15)
    head = {"kernel": rng.standard_normal((4, 1)).astype(np.float32), "bias": np.zeros(1, np.float32), "activation": "sigmoid"}
    ref = oracle.dlrm_forward(batch, tables, f2t, cont, bottom, top, head)
    tt = lambda ls: [{k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in l.items()} for l in ls]
    got = oracle_torch.dlrm_forward({n: torch.from_numpy(batch[n]) for n in names},
                                    {c: torch.from_numpy(batch[c]) for c in cont},
                                    {n: torch.from_numpy(t) for n, t in tables.items()}, f2t, tt(bottom), tt(top), tt([head])[0])
    np.testing.assert_allclose(got.numpy(), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", sorted(p.name for p in GOLDEN.glob("*.npz")))
def test_golden_fixture(name):
    """Committed golden vectors: oracle outputs must not drift (oracle_*.npz, made by
    tests/golden/make_golden.py) and must equal what the reference's own torch backend computed
    (ref_torch_*.npz, made by oracle/make_golden_from_reference_torch.py)."""
    from tests.golden import replay

    replay.check(GOLDEN / name)


def test_training_port_matches_oracle_train():
    """oracle_torch.DLRMTrainCPU (the timed CPU baseline of bench.py's training record: sparse-gradient autograd +
    torch.optim.Adagrad) against oracle_train (float64 autograd + the Keras update rule restated in NumPy)."""
    import torch

    from oracle import oracle_torch, oracle_train

    rng = np.random.default_rng(0)
    tabs = {"a": rng.normal(size=(7, 8)).astype(np.float32), "b": rng.normal(size=(5, 8)).astype(np.float32)}
    f2t = {"a": "a", "b": "b"}

    def layer(k, n, act):
        return {"kernel": (rng.normal(size=(k, n)) * 0.3).astype(np.float32), "bias": np.zeros(n, np.float32), "activation": act}

    bottom, top, head = [layer(2, 8, "relu")], [layer(8 + 3, 6, "relu")], layer(6, 1, "linear")
    B = 16
    batch = {"a": rng.integers(0, 7, B), "b": rng.integers(0, 5, B), "x": rng.random(B).astype(np.float32), "y": rng.random(B).astype(np.float32)}
    t = rng.integers(0, 2, B)
    loss, _, g = oracle_train.dlrm_loss_and_grads(batch, tabs, f2t, ["x", "y"], bottom, top, head, t)
    tt = lambda ls: [{k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in l.items()} for l in ls]
    cpu = oracle_torch.DLRMTrainCPU({k: torch.from_numpy(v) for k, v in tabs.items()}, f2t, tt(bottom), tt(top), tt([head])[0], lr=0.1)
    got = cpu.step({k: torch.from_numpy(batch[k]) for k in f2t}, {k: torch.from_numpy(batch[k]) for k in ("x", "y")}, torch.from_numpy(t))
    np.testing.assert_allclose(got, loss, rtol=1e-6)
    for name in ("a", "b"):
        uniq = np.unique(batch[name])
        ref = oracle_train.sparse_update("adagrad", tabs[name], uniq, g[f"table/{name}"][uniq], {"a": np.full(tabs[name].shape, 0.1)}, 0.1)
        np.testing.assert_allclose(cpu.tables[name].detach().numpy(), ref, rtol=0, atol=1e-6)
    ref_k = oracle_train.dense_update("adagrad", top[0]["kernel"], g["top/kernel_0"], {"a": np.full(top[0]["kernel"].shape, 0.1)}, 0.1)
    np.testing.assert_allclose(cpu.layers["top"][0]["kernel"].detach().numpy(), ref_k, rtol=0, atol=1e-6)


@pytest.mark.parametrize("opt", ["sgd", "adagrad"])
def test_optimizer_rules_match_torch_optim(opt):
    """The Keras update rules restated in oracle_train coincide with torch.optim for SGD and Adagrad (dense and, with
    duplicate ids summed first, sparse)."""
    import torch

    from oracle import oracle_train

    rng = np.random.default_rng(3)
    w = rng.normal(size=(9, 4))
    ids = np.array([1, 1, 5, 8, 1, 5])
    vals = rng.normal(size=(6, 4))
    p = torch.nn.Parameter(torch.from_numpy(w.copy()))
    o = torch.optim.SGD([p], lr=0.1) if opt == "sgd" else torch.optim.Adagrad([p], lr=0.1, initial_accumulator_value=0.1, eps=1e-7)
    state = {} if opt == "sgd" else {"a": np.full(w.shape, 0.1)}
    ref = w
    for _ in range(2):
        dense = np.zeros_like(w)
        np.add.at(dense, ids, vals)
        p.grad = torch.from_numpy(dense.copy())
        o.step()
        ref = oracle_train.sparse_update(opt, ref, ids, vals, state, 0.1)
        np.testing.assert_allclose(p.detach().numpy(), ref, rtol=1e-12, atol=1e-12)


def test_fm_pairwise_and_fm_block_known_answers():
    """tests/unit/tf/blocks/test_interactions.py:25-36: (100, 10, 64) -> (100, 64); and the FMBlock composition written out
    by hand for three samples (one-hot wide Dense(1) = row lookup; pairwise over the D components of each feature, the
    reference's StackFeatures(axis=-1) layout)."""
    rng = np.random.default_rng(0)
    x = rng.random((100, 10, 64)).astype(np.float32)
    out = oracle.fm_pairwise(x)
    assert out.shape == (100, 64)
    np.testing.assert_allclose(out[3, 5], 0.5 * (x[3, :, 5].sum() ** 2 - (x[3, :, 5] ** 2).sum()), rtol=1e-5)
    tabs = {"a": rng.normal(size=(4, 8)).astype(np.float32), "b": rng.normal(size=(6, 8)).astype(np.float32)}
    batch = {"a": np.array([0, 3, 1]), "b": np.array([5, 5, 2]), "x": np.array([0.1, 0.2, 0.3], np.float32)}
    wk = rng.normal(size=(4 + 6 + 1, 1)).astype(np.float32)
    got = oracle.fm_block(batch, tabs, {"a": "a", "b": "b"}, ["x"], {"a": 4, "b": 6}, wk, np.array([0.5], np.float32))
    want = []
    for i in range(3):
        w = wk[batch["a"][i], 0] + wk[4 + batch["b"][i], 0] + wk[10, 0] * batch["x"][i] + 0.5
        p = sum(0.5 * (t[batch[n][i]].sum() ** 2 - (t[batch[n][i]] ** 2).sum()) for t, n in ((tabs["a"], "a"), (tabs["b"], "b")))
        want.append(w + p)
    np.testing.assert_allclose(got.reshape(-1), np.array(want), rtol=1e-5)
