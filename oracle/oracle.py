"""CPU restatement (NumPy, fp32) of the Merlin Models hot path — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg
may import this package.  The product (`models_b200/`) never does and has no CPU path.

Each function restates the reference lines it cites (paths relative to
/root/reference/merlin/models/).  The arithmetic itself lives in TensorFlow (`tensorflow>=2.9,<2.13`,
requirements/tensorflow.txt:1 — not vendored, not installable here: no network), so the ops are
restated from their documented semantics: tf.gather / Keras Embedding = row copy,
tf.nn.safe_embedding_lookup_sparse = prune ids < 0, segment-reduce, empty row -> 0,
Keras Dense = x @ kernel + bias with kernel (in, out), tf.matmul(transpose_b) etc.

PINNING STATUS
  * pinned by the reference's own known-answer tests (re-stated in tests/test_reference_invariants.py):
    output widths F(F-1)/2 (+D) (tests/unit/tf/blocks/test_dlrm.py:36-52), false-negative diagonal
    == false_negative_score and off-diagonal != (tests/unit/tf/outputs/test_contrastive.py:173-206),
    inference scorer (B,1) (:209-223), shared-table rows identical
    (tests/unit/tf/inputs/test_embedding.py:248-253), inferred dims (:485-553), L2-normalised towers
    (tests/unit/tf/blocks/retrieval/test_two_tower.py:94-107);
  * pinned against the reference's OWN torch backend executed in the build container
    (oracle/make_golden_from_reference_torch.py -> tests/golden/ref_torch_*.npz): sorted-name concat / stack
    order, DLRM interaction ordering/values and the [bottom | interactions] concat, the WHOLE DLRMBlock end to end
    (tables -> bottom MLP -> stack -> interaction -> top MLP; ref_torch_dlrm_block.npz) and the torch DLRMModel /
    DCNModel with BinaryOutput (ref_torch_dlrm_model.npz, ref_torch_dcn_model.npz), weight-tied catalog logits +
    cross-entropy + top-k (EmbeddingTablePrediction, ref_torch_catalog.npz), embedding-bag combiners,
    MLP, DCN-v2 cross, the [positive | negatives] logits layout with false-negative rescoring and one-hot
    targets (ContrastiveOutput.contrastive_outputs, rescore_false_negatives, InBatchNegativeSampler), the
    log-uniform sampling distribution (the torch backend is one class short of the TF formula and its
    "unique" variant computes a different quantity: characterised in tests/golden/replay.py, TF is the target);
  * the remaining TF-only details (logQ correction arithmetic, safe_embedding_lookup_sparse's pruning of
    ids < 0 / empty-bag zeros, the legacy InputBlock dict -> first-Dense order, top-k metric formulas) are
    restated from source and are PARITY UNPINNED by executed reference code: TensorFlow cannot be run here.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

MIN_FLOAT = float(np.finfo(np.float16).min) / 100.0  # utils/constants.py:19  -> -655.04

F32 = np.float32


# ------------------------------------------------------------------------------------------------
# deterministic table init shared with the CUDA initialiser (models_b200/csrc/cabi.cu)
# ------------------------------------------------------------------------------------------------
def hash_uniform(index, seed: int, lo: float = -0.05, hi: float = 0.05) -> np.ndarray:
    """w[i] = lo + (hi-lo) * u24(splitmix64(seed + (i+1)*phi)); bit-identical to
    mm_init_uniform_hash.  `index` = flat element indices (any shape)."""
    i = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & (2**64 - 1)) + (i + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * F32(1.0 / 16777216.0)
    span = F32(F32(hi) - F32(lo))
    return (F32(lo) + (span * u).astype(np.float32)).astype(np.float32)


def hash_table_rows(rows, dim: int, seed: int, lo: float = -0.05, hi: float = 0.05) -> np.ndarray:
    """Rows `rows` of a (n_rows, dim) table initialised by mm_init_uniform_hash."""
    rows = np.asarray(rows, dtype=np.uint64).reshape(-1, 1)
    flat = rows * np.uint64(dim) + np.arange(dim, dtype=np.uint64).reshape(1, -1)
    return hash_uniform(flat, seed, lo, hi)


# ------------------------------------------------------------------------------------------------
# a1  input canonicalisation
# ------------------------------------------------------------------------------------------------
def prepare_features(batch: Dict[str, np.ndarray]) -> Dict[str, object]:
    """transforms/features.py:324-379 + :168-234 + utils/tf_utils.py:477-481.
    Scalars (B,) -> (B,1); `name__values` + `name__offsets` -> ragged, returned as the pair
    (values, row_splits) under `name`."""
    out: Dict[str, object] = {}
    for k, v in batch.items():
        if k.endswith("__offsets"):
            continue
        if k.endswith("__values"):
            name = k[: -len("__values")]
            out[name] = (np.asarray(v), np.asarray(batch[name + "__offsets"]))
        else:
            v = np.asarray(v)
            out[k] = v.reshape(-1, 1) if v.ndim == 1 else v
    return out


# ------------------------------------------------------------------------------------------------
# a2/a3  embeddings
# ------------------------------------------------------------------------------------------------
def infer_embedding_dim(cardinality: int, multiplier: float = 2.0, ensure_multiple_of_8: bool = True) -> int:
    """utils/schema_utils.py:169-207 (cardinality = int_domain.max + 1)."""
    size = int(math.ceil(math.pow(cardinality, 0.25) * multiplier))
    if ensure_multiple_of_8:
        size = int(math.ceil(size / 8) * 8)
    return size


def embedding_lookup(table: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """inputs/embedding.py:457-461 / :1142-1146: squeeze trailing 1, gather rows. TF-CPU raises on
    an out-of-range id; so do we."""
    ids = np.asarray(ids)
    if ids.ndim > 1 and ids.shape[-1] == 1:
        ids = ids.reshape(ids.shape[:-1])
    if ids.size and (ids.min() < 0 or ids.max() >= table.shape[0]):
        raise IndexError("indices out of range for embedding table")
    return table[ids]


def embedding_bag(table: np.ndarray, values: np.ndarray, offsets: np.ndarray, combiner: str = "mean") -> np.ndarray:
    """tf.nn.safe_embedding_lookup_sparse as called at inputs/embedding.py:432-441,:1139:
    ids < 0 pruned, empty bag -> zeros, sum left-to-right in fp32; mean = sum / n,
    sqrtn = sum / sqrt(n).  Cross-check: torch/inputs/embedding.py:264-293 (embedding_bag)."""
    if combiner not in ("mean", "sum", "sqrtn"):
        raise ValueError(f"combiner {combiner!r}")
    B = len(offsets) - 1
    out = np.zeros((B, table.shape[1]), dtype=np.float32)
    for b in range(B):
        acc = np.zeros(table.shape[1], dtype=np.float32)
        n = 0
        for i in values[int(offsets[b]): int(offsets[b + 1])]:
            if i < 0:
                continue
            acc = (acc + table[int(i)]).astype(np.float32)
            n += 1
        if n:
            if combiner == "mean":
                acc = (acc / F32(n)).astype(np.float32)
            elif combiner == "sqrtn":
                acc = (acc / np.sqrt(F32(n))).astype(np.float32)
        out[b] = acc
    return out


def sequence_combiner(x: np.ndarray, combiner: str) -> np.ndarray:
    """inputs/embedding.py:1545-1587: mean/sum/max over axis 1 of (B, L, D); padding not masked.
    The sum runs left to right in fp32."""
    if combiner == "max":
        return x.max(axis=1)
    acc = np.zeros((x.shape[0], x.shape[2]), dtype=np.float32)
    for l in range(x.shape[1]):
        acc = (acc + x[:, l]).astype(np.float32)
    if combiner == "mean":
        acc = (acc / F32(x.shape[1])).astype(np.float32)
    elif combiner != "sum":
        raise ValueError(f"combiner {combiner!r}")
    return acc


def embed_feature(table: np.ndarray, feat, combiner: Optional[str]) -> np.ndarray:
    """One feature through EmbeddingTable._call_table (inputs/embedding.py:424-471):
    ragged pair -> bag lookup; (B,1)/(B,) -> row gather; (B,L) dense -> gather + combiner."""
    if isinstance(feat, tuple):
        return embedding_bag(table, feat[0], feat[1], combiner or "mean")
    feat = np.asarray(feat)
    if feat.ndim == 2 and feat.shape[1] > 1:
        return sequence_combiner(embedding_lookup(table, feat), combiner or "mean")
    return embedding_lookup(table, feat).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# a4/a6/a8  aggregations: sorted(name) order, fp32
# ------------------------------------------------------------------------------------------------
def concat_features(d: Dict[str, np.ndarray]) -> np.ndarray:
    """core/aggregation.py:54-66."""
    return np.concatenate([np.asarray(d[k], dtype=np.float32).reshape(len(d[k]), -1) for k in sorted(d)], axis=-1)


def stack_features(d: Dict[str, np.ndarray], axis: int = 1) -> np.ndarray:
    """core/aggregation.py:101-108."""
    return np.stack([np.asarray(d[k], dtype=np.float32) for k in sorted(d)], axis=axis)


# ------------------------------------------------------------------------------------------------
# a5  MLP
# ------------------------------------------------------------------------------------------------
def activation(x: np.ndarray, name: Optional[str]) -> np.ndarray:
    x = x.astype(np.float32)
    if name in (None, "linear"):
        return x
    if name == "relu":
        return np.maximum(x, F32(0))
    if name == "sigmoid":
        return (F32(1) / (F32(1) + np.exp(-x))).astype(np.float32)
    if name == "tanh":
        return np.tanh(x)
    if name == "selu":
        a, s = F32(1.6732632423543772), F32(1.0507009873554805)
        return np.where(x > 0, s * x, s * a * np.expm1(x)).astype(np.float32)
    if name == "elu":
        return np.where(x > 0, x, np.expm1(x)).astype(np.float32)
    if name == "gelu":
        from scipy.special import erf
        return (F32(0.5) * x * (F32(1) + erf(x * F32(0.70710678118654752)))).astype(np.float32)
    raise ValueError(f"activation {name!r}")


def dense(x: np.ndarray, kernel: np.ndarray, bias: Optional[np.ndarray], act: Optional[str]) -> np.ndarray:
    """Keras Dense as used by _Dense.call (blocks/mlp.py:275-280): act(x @ kernel + bias)."""
    y = np.matmul(x.astype(np.float32), kernel.astype(np.float32))
    if bias is not None:
        y = y + bias.astype(np.float32)
    return activation(y.astype(np.float32), act)


def mlp(x, layers: Sequence[dict]) -> np.ndarray:
    """MLPBlock (blocks/mlp.py:35-139): layers = [{"kernel","bias","activation"}...]; dict input is
    concat-aggregated in sorted-key order first (:275-277).  Dropout is identity at inference."""
    if isinstance(x, dict):
        x = concat_features(x)
    for l in layers:
        x = dense(x, l["kernel"], l.get("bias"), l.get("activation"))
        bn = l.get("batch_norm")
        if bn is not None:  # Keras BatchNormalization inference, epsilon 1e-3
            x = ((x - bn["mean"]) / np.sqrt(bn["var"] + F32(1e-3)) * bn["gamma"] + bn["beta"]).astype(np.float32)
    return x


# ------------------------------------------------------------------------------------------------
# a7  DLRM interaction
# ------------------------------------------------------------------------------------------------
def dot_interaction(x: np.ndarray, self_interaction: bool = False) -> np.ndarray:
    """blocks/interaction.py:86-116 with interaction_type=None: Z = X X^T, boolean-mask of the
    (strict) upper triangle in row-major order.  Cross-check: torch/blocks/dlrm.py:65-74
    (triu_indices(F, F, offset=1))."""
    x = x.astype(np.float32)
    z = np.matmul(x, np.transpose(x, (0, 2, 1)))
    F = x.shape[1]
    mask = np.triu(np.ones((F, F), dtype=bool), k=0 if self_interaction else 1)
    return z[:, mask]


def dlrm_forward(batch: Dict[str, np.ndarray], tables: Dict[str, np.ndarray], feature_table: Dict[str, str],
                 continuous: Sequence[str], bottom: Optional[Sequence[dict]], top: Optional[Sequence[dict]],
                 head: Optional[dict], combiner: str = "mean", return_intermediates: bool = False):
    """DLRMModel forward (models/ranking.py:23-92 -> blocks/dlrm.py:32-133 -> outputs/classification.py:114).

    feature_table: categorical feature name -> table name (shared tables, inputs/embedding.py:668-679).
    """
    feats = prepare_features(batch)
    emb = {name: embed_feature(tables[tname], feats[name], combiner) for name, tname in feature_table.items()}
    inter_in = dict(emb)
    bottom_out = None
    if continuous:
        # ContinuousFeatures (inputs/continuous.py:117-138) -> bottom MLP; its _Dense concats sorted
        bottom_out = mlp({n: np.asarray(feats[n], dtype=np.float32) for n in continuous}, bottom)
        inter_in["bottom_block"] = bottom_out  # ParallelBlock merge, core/combinators.py:562-571
    stacked = stack_features(inter_in, axis=1)  # dlrm.py:169-170
    inter = dot_interaction(stacked)
    if top is None:
        body = inter
    elif bottom_out is None:
        body = mlp(inter, top)
    else:
        # WithShortcut + Filter("bottom_block") + concat over sorted
        # {"bottom_block", "sequential_block[_N]"} -> bottom first (dlrm.py:126-130)
        body = mlp(np.concatenate([bottom_out, inter], axis=1), top)
    out = body
    if head is not None:
        out = dense(body, head["kernel"], head.get("bias"), head.get("activation", "sigmoid"))
    if return_intermediates:
        return out, {"stacked": stacked, "interactions": inter, "bottom": bottom_out, "body": body}
    return out


# ------------------------------------------------------------------------------------------------
# a10  DCN-v2
# ------------------------------------------------------------------------------------------------
def cross_layers(x0: np.ndarray, layers: Sequence[dict]) -> np.ndarray:
    """CrossBlock (blocks/cross.py:29-109) of Cross.call (:188-202):
    x_{l+1} = x0 * (x_l @ W_l + b_l) + x_l."""
    x0 = x0.astype(np.float32)
    x = x0
    for l in layers:
        proj = dense(x, l["kernel"], l.get("bias"), None)
        if "kernel_u" in l:  # low-rank: dense(dense_u(x)) (blocks/mlp.py:389-396)
            proj = dense(dense(x, l["kernel_u"], None, None), l["kernel"], l.get("bias"), None)
        x = (x0 * proj + x).astype(np.float32)
    return x


def dcn_forward(batch, tables, feature_table, continuous, cross, deep, head, stacked: bool = True,
                combiner: str = "mean", branch_order=("deep", "cross")) -> np.ndarray:
    """DCNModel (models/ranking.py:95-168): InputBlockV2 concat (sorted names over embeddings +
    continuous, inputs/base.py:216-341) -> CrossBlock -> deep MLP -> BinaryOutput."""
    feats = prepare_features(batch)
    d = {name: embed_feature(tables[tname], feats[name], combiner) for name, tname in feature_table.items()}
    for n in continuous:
        d[n] = np.asarray(feats[n], dtype=np.float32)
    x0 = concat_features(d)
    if stacked:
        body = mlp(cross_layers(x0, cross), deep)
    else:
        # connect_branch(CrossBlock(depth), deep_block, aggregation="concat") (models/ranking.py:159-166): a ParallelBlock
        # keyed by the layers' auto names, concatenated in sorted-key order (core/aggregation.py:54-66).  Both are
        # `sequential_block[_N]`; the deep block exists before CrossBlock(depth) is built, so normally [deep | cross].
        br = {"cross": cross_layers(x0, cross), "deep": mlp(x0, deep)}
        body = np.concatenate([br[branch_order[0]], br[branch_order[1]]], axis=1)
    return dense(body, head["kernel"], head.get("bias"), head.get("activation", "sigmoid"))


def fm_pairwise(x: np.ndarray) -> np.ndarray:
    """FMPairwiseInteraction.call (blocks/interaction.py:217-236): x (bs, n_features, embedding_dim) ->
    0.5 * ((sum over axis 1)^2 - sum over axis 1 of squares), shape (bs, embedding_dim)."""
    x = np.asarray(x, dtype=np.float32)
    assert x.ndim == 3, "inputs should be a 3-D tensor"
    return (0.5 * (np.square(x.sum(axis=1)) - np.square(x).sum(axis=1))).astype(np.float32)


def fm_block(batch, tables, feature_table, continuous, cardinalities, wide_kernel, wide_bias) -> np.ndarray:
    """FMBlock (blocks/interaction.py:256-332) with its default blocks -> (B, 1).
    first order: concat in sorted-name order of one-hot(categorical, depth = int_domain.max + 1; CategoryEncoding,
    schema_utils.categorical_cardinalities) and the continuous columns -> Dense(1, linear).
    pairwise: embeddings stacked with StackFeatures(axis=-1) -> (B, D, F), FMPairwiseInteraction over axis 1, then
    reduce_sum over axis 1 keepdims (:323-328)."""
    feats = prepare_features(batch)
    names = sorted(list(feature_table) + list(continuous))
    B = len(np.asarray(feats[names[0]]).reshape(-1))
    cols = []
    for n in names:
        if n in feature_table:
            ids = np.asarray(feats[n]).reshape(-1).astype(np.int64)
            oh = np.zeros((B, int(cardinalities[n])), dtype=np.float32)
            ok = (ids >= 0) & (ids < oh.shape[1])
            oh[np.nonzero(ok)[0], ids[ok]] = 1.0
            cols.append(oh)
        else:
            cols.append(np.asarray(feats[n], dtype=np.float32).reshape(B, 1))
    first = dense(np.concatenate(cols, axis=1), wide_kernel, wide_bias, "linear")
    emb = {n: embed_feature(tables[t], feats[n], None) for n, t in feature_table.items()}
    stacked = np.stack([emb[k] for k in sorted(emb)], axis=-1)  # (B, D, F): StackFeatures(axis=-1)
    pair = fm_pairwise(stacked).sum(axis=1, keepdims=True)
    return (first + pair).astype(np.float32)


def deepfm_forward(batch, tables, feature_table, continuous, cardinalities, wide_kernel, wide_bias, deep, deep_logit, head) -> np.ndarray:
    """DeepFMModel (models/ranking.py:171-279): sum of the FM tower (fm_block) and the deep tower (sorted-name concat of
    embeddings + continuous -> deep MLP -> MLPBlock([1], linear)), then BinaryOutput's Dense(1, sigmoid) on the 1-wide sum."""
    feats = prepare_features(batch)
    d = {name: embed_feature(tables[tname], feats[name], None) for name, tname in feature_table.items()}
    for n in continuous:
        d[n] = np.asarray(feats[n], dtype=np.float32)
    deep_out = mlp(mlp(concat_features(d), deep), deep_logit)
    z = fm_block(batch, tables, feature_table, continuous, cardinalities, wide_kernel, wide_bias) + deep_out
    return dense(z, head["kernel"], head.get("bias"), head.get("activation", "sigmoid"))


# ------------------------------------------------------------------------------------------------
# a11-a13  two-tower + scorers
# ------------------------------------------------------------------------------------------------
def tower_forward(batch, tables, feature_table, continuous, layers, combiner: str = "mean",
                  l2_normalize: bool = False) -> np.ndarray:
    """One tower of TwoTowerBlock (blocks/retrieval/two_tower.py:98-118): legacy InputBlock
    (inputs/base.py:180-206) -> dict of features -> MLP whose first _Dense concats sorted names."""
    feats = prepare_features(batch)
    d = {name: embed_feature(tables[tname], feats[name], combiner) for name, tname in feature_table.items()}
    for n in continuous:
        d[n] = np.asarray(feats[n], dtype=np.float32)
    out = mlp(d, layers)
    if l2_normalize:  # transforms/regularization.py L2Norm: tf.linalg.l2_normalize(axis=-1), eps 1e-12
        nrm = np.sqrt(np.maximum((out * out).sum(-1, keepdims=True), F32(1e-12)))
        out = (out / nrm).astype(np.float32)
    return out


def rescore_false_negatives(pos_ids, neg_ids, neg_scores, false_negatives_score=MIN_FLOAT):
    """utils/tf_utils.py:126-154."""
    pos_ids = np.asarray(pos_ids).reshape(-1).astype(np.asarray(neg_ids).dtype)
    neg_ids = np.asarray(neg_ids).reshape(-1)
    mask = pos_ids[:, None] == neg_ids[None, :]
    scores = np.where(mask, F32(false_negatives_score), neg_scores).astype(np.float32)
    return scores, ~mask


def retrieval_scores(query: np.ndarray, item: np.ndarray) -> np.ndarray:
    """Inference scorer: blocks/retrieval/base.py:278-281 — (B,1)."""
    return (query.astype(np.float32) * item.astype(np.float32)).sum(-1, keepdims=True).astype(np.float32)


def log_uniform_probs(max_id: int, min_id: int = 0, unique: bool = True, n_sampled: int = 0) -> np.ndarray:
    """outputs/sampling/popularity.py:141-165: p_k = (log(k+2)-log(k+1))/log(R+2), k=0..R,
    R = max_id - min_id; unique -> -expm1(n*log1p(-p)); left-padded with min_id zeros."""
    R = max_id - min_id
    k = np.arange(R + 1, dtype=np.float64)
    p = (np.log(k + 2.0) - np.log(k + 1.0)) / np.log(R + 2.0)
    if unique:
        p = -np.expm1(n_sampled * np.log1p(-p))
    return np.concatenate([np.zeros(min_id), p]).astype(np.float32)


def contrastive_logits(query, pos_item, neg_items, pos_ids=None, neg_ids=None, downscore=True,
                       false_negative_score=MIN_FLOAT, pos_prob=None, neg_prob=None,
                       temperature: float = 1.0) -> Tuple[np.ndarray, np.ndarray]:
    """ItemRetrievalScorer.call_outputs (blocks/retrieval/base.py:339-422) ==
    ContrastiveOutput.outputs (outputs/contrastive.py:303-340): [pos | masked(Q N^T)] fp32,
    targets one-hot on column 0; then LogitsTemperatureScaler (predictions / T)."""
    q = query.astype(np.float32)
    neg = np.matmul(q, neg_items.astype(np.float32).T)
    pos = (q * pos_item.astype(np.float32)).sum(-1, keepdims=True)
    if pos_prob is not None and neg_prob is not None:  # logQ, contrastive.py:317-319
        pos = pos - np.log(np.asarray(pos_prob, np.float32).reshape(-1, 1) + F32(1e-16))
        neg = neg - np.log(np.asarray(neg_prob, np.float32).reshape(1, -1) + F32(1e-16))
    if downscore:
        neg, _ = rescore_false_negatives(pos_ids, neg_ids, neg, false_negative_score)
    out = np.concatenate([pos, neg], axis=-1).astype(np.float32)
    out = (out / F32(temperature)).astype(np.float32)
    targets = np.zeros_like(out)
    targets[:, 0] = 1.0
    return out, targets


def catalog_logits(x: np.ndarray, table: np.ndarray, bias: Optional[np.ndarray] = None) -> np.ndarray:
    """outputs/classification.py:347-357 (EmbeddingTablePrediction) /
    blocks/retrieval/base.py:431-438: x @ E^T (+ bias)."""
    y = np.matmul(x.astype(np.float32), table.astype(np.float32).T)
    if bias is not None:
        y = y + bias.astype(np.float32)
    return y.astype(np.float32)


def softmax_ce_stats(logits: np.ndarray, targets: np.ndarray) -> np.ndarray:
    """Inputs of CategoricalCrossEntropy(from_logits=True) (losses/listwise.py:38-50):
    per row [max, logsumexp, logit[target]] (computed in float64, returned fp32)."""
    l = logits.astype(np.float64)
    m = l.max(-1)
    lse = m + np.log(np.exp(l - m[:, None]).sum(-1))
    t = l[np.arange(len(l)), np.asarray(targets).reshape(-1)]
    return np.stack([m, lse, t], axis=1).astype(np.float32)


def topk(logits: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """tf.math.top_k (outputs/topk.py:221-223, core/index.py:236-237): descending, ties -> lower
    index first."""
    idx = np.argsort(-logits, axis=-1, kind="stable")[:, :k]
    return np.take_along_axis(logits, idx, axis=-1), idx.astype(np.int64)


# ------------------------------------------------------------------------------------------------
# top-k retrieval evaluation
# ------------------------------------------------------------------------------------------------
def topk_index(queries: np.ndarray, values: np.ndarray, ids: Optional[np.ndarray], k: int):
    """TopKIndexBlock.call (core/index.py:232-250) / BruteForce.call (outputs/topk.py:221-223):
    scores = queries @ values^T; top_k; ids gathered.  Returns (top_scores (B,k), top_ids (B,k))."""
    scores = queries.astype(np.float32) @ values.astype(np.float32).T
    s, idx = topk(scores, k)
    return s, (idx if ids is None else np.asarray(ids)[idx])


def topk_targets(positive_ids: np.ndarray, top_ids: np.ndarray) -> np.ndarray:
    """core/index.py:272-276, outputs/topk.py:236-238: one-hot of the positive id among the top ids."""
    return (np.asarray(positive_ids).reshape(-1, 1) == top_ids).astype(np.float32)


def _div_no_nan(a, b):
    return np.where(b != 0, a / np.where(b != 0, b, 1), 0.0).astype(np.float32)


def recall_at(y_true, label_relevant_counts, k):
    """metrics/topk.py:48-66."""
    rel = np.clip(label_relevant_counts, 1, float(k))
    return _div_no_nan(y_true[:, :k].sum(-1), rel)


def precision_at(y_true, label_relevant_counts, k):
    """metrics/topk.py:69-84."""
    return y_true[:, :k].mean(-1).astype(np.float32)


def average_precision_at(y_true, label_relevant_counts, k):
    """metrics/topk.py:87-113."""
    precisions = np.stack([precision_at(y_true, None, j) for j in range(1, k + 1)], axis=-1)
    total = (precisions * y_true[:, :k]).sum(-1)
    return _div_no_nan(total, np.clip(label_relevant_counts, 1, float(k)))


def dcg_at(y_true, k, log_base=2):
    """metrics/topk.py:116-139."""
    disc = 1.0 / (np.log(np.arange(k, dtype=np.float32) + 2) / np.log(np.float32(log_base)))
    return (y_true[:, :k] * disc[None, :]).sum(-1).astype(np.float32)


def ndcg_at(y_true, label_relevant_counts, k, log_base=2):
    """metrics/topk.py:142-166."""
    ideal = (np.arange(k)[None, :] < np.asarray(label_relevant_counts)[:, None]).astype(np.float32)
    return _div_no_nan(dcg_at(y_true, k, log_base), dcg_at(ideal, k, log_base))


def mrr_at(y_true, label_relevant_counts, k):
    """metrics/topk.py:169-187."""
    first = (np.argmax(y_true, axis=-1) + 1).astype(np.float32)
    hit = y_true[:, :k].max(-1)
    return _div_no_nan(np.ones_like(first), first * hit)
