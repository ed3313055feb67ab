"""Two-tower retrieval: TwoTowerBlock, in-batch sampler, ItemRetrievalScorer / ItemRetrievalTask
(v1 API used by mm.TwoTowerModel) and ContrastiveOutput (v2 API).

Reference: merlin/models/tf/blocks/retrieval/{two_tower,base}.py, blocks/sampling/in_batch.py,
prediction_tasks/retrieval.py, outputs/contrastive.py, outputs/sampling/{in_batch,popularity}.py,
utils/tf_utils.py:126-154.  The scorer is one fused kernel: Q.N^T + false-negative mask + logQ +
[pos | neg] layout + temperature, writing the (B, 1+N) logits exactly once.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from .blocks import MLP, dense_engine
from .core import Block, Prediction, TabularData, unique_name
from .inputs import EmbeddingOptions, InputBlock
from .schema import Schema, Tags

MIN_FLOAT = float(np.finfo(np.float16).min) / 100.0  # merlin/models/utils/constants.py:19


class L2Norm(Block):
    """transforms/regularization.py:27-82."""

    def call(self, inputs, **kwargs):
        if isinstance(inputs, dict):
            return {k: ops.l2_normalize(v) for k, v in inputs.items()}
        return ops.l2_normalize(inputs)


class TowerBlock(Block):
    """One tower: legacy InputBlock(schema subset) -> tower MLP (two_tower.py:98-118).  The
    sorted-name concat that the first _Dense applies to the InputBlock's dict is produced directly
    by the fused gather (embeddings land at their concat offsets)."""

    def __init__(self, inputs: InputBlock, mlp: MLP, name: str):
        super().__init__(name)
        self.inputs = inputs
        self.mlp = mlp

    def build(self, device=None):
        self.inputs.build(device)
        _, _, width = self.inputs.layout()
        self.mlp.build_from_width(width, device)
        self.built = True
        return self

    def weights(self):
        out = {f"inputs/{k}": v for k, v in self.inputs.weights().items()}
        out.update({f"mlp/{k}": v for k, v in self.mlp.weights().items()})
        return out

    def call(self, inputs: TabularData, **kwargs) -> torch.Tensor:
        return self.mlp(self.inputs.concat(inputs), **kwargs)


class TwoTowerBlock(Block):
    """blocks/retrieval/two_tower.py:32-118 / DualEncoderBlock (retrieval/base.py:59-129)."""

    def __init__(self, schema: Schema, query_tower: MLP, item_tower: Optional[MLP] = None,
                 query_tower_tag=Tags.USER, item_tower_tag=Tags.ITEM,
                 embedding_options: EmbeddingOptions = EmbeddingOptions(embedding_dims=None, embedding_dim_default=64,
                                                                        infer_embedding_sizes=False,
                                                                        infer_embedding_sizes_multiplier=2.0),
                 post: Optional[Block] = None, **kwargs):
        if schema is None:
            raise ValueError("The schema is required by TwoTower")
        if query_tower is None:
            raise ValueError("The query_tower is required by TwoTower")
        super().__init__(unique_name("two_tower_block"))
        _item_tower = item_tower or query_tower.copy()
        if isinstance(_item_tower, TowerBlock):
            self.item = _item_tower
        else:
            item_schema = schema.select_by_tag(item_tower_tag) if item_tower_tag else schema
            if not item_schema:
                raise ValueError(
                    f"The schema should contain features with the tag `{item_tower_tag}`,"
                    "required by item-tower"
                )
            self.item = TowerBlock(InputBlock(item_schema, embedding_options=embedding_options), _item_tower, "item")
        if isinstance(query_tower, TowerBlock):
            self.query = query_tower
        else:
            query_schema = schema.select_by_tag(query_tower_tag) if query_tower_tag else schema
            if not query_schema:
                raise ValueError(
                    f"The schema should contain features with the tag `{query_schema}`,"
                    "required by query-tower"
                )
            self.query = TowerBlock(InputBlock(query_schema, embedding_options=embedding_options), query_tower, "query")
        if isinstance(post, str):
            if post not in ("l2-norm", "l2_norm"):
                raise ValueError(f"Unknown post block {post!r}")
            post = L2Norm()
        self.post = post
        self.schema = schema

    def build(self, device=None):
        self.query.build(device)
        self.item.build(device)
        self.built = True
        return self

    def weights(self):
        out = {f"query/{k}": v for k, v in self.query.weights().items()}
        out.update({f"item/{k}": v for k, v in self.item.weights().items()})
        return out

    def call(self, inputs: TabularData, **kwargs) -> Dict[str, torch.Tensor]:
        out = {"query": self.query(inputs, **kwargs), "item": self.item(inputs, **kwargs)}
        if self.post is not None:
            out = self.post(out)
        return out


# ------------------------------------------------------------------------------------------------
# samplers
# ------------------------------------------------------------------------------------------------
class InBatchSampler:
    """blocks/sampling/in_batch.py:25-113 / outputs/sampling/in_batch.py:25-100: the negatives are
    the batch's own item embeddings and ids (identity)."""

    def __init__(self, batch_size: Optional[int] = None, **kwargs):
        self.batch_size = batch_size

    def sample(self, item_embeddings: torch.Tensor, item_ids: Optional[torch.Tensor]):
        return item_embeddings, item_ids, None


InBatchSamplerV2 = InBatchSampler


def log_uniform_sampling_probs(max_id: int, min_id: int = 0, max_num_samples: int = 0, unique: bool = True) -> np.ndarray:
    """PopularityBasedSamplerV2 sampling probabilities (outputs/sampling/popularity.py:141-165)."""
    R = max_id - min_id
    k = np.arange(R + 1, dtype=np.float64)
    p = (np.log(k + 2.0) - np.log(k + 1.0)) / np.log(R + 2.0)
    if unique:
        p = -np.expm1(max_num_samples * np.log1p(-p))
    return np.concatenate([np.zeros(min_id), p]).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# scorer (v1) and contrastive output (v2)
# ------------------------------------------------------------------------------------------------
_TARGET_ROWS: Dict[tuple, torch.Tensor] = {}


def _target_row(width: int, device) -> torch.Tensor:
    """[1, 0, 0, ...] (cached per width/device; built without host scalars so it is graph-capturable)."""
    key = (width, str(device))
    row = _TARGET_ROWS.get(key)
    if row is None:
        row = torch.zeros(width, dtype=torch.float32, device=device)
        row[:1].fill_(1.0)
        if len(_TARGET_ROWS) > 16:
            _TARGET_ROWS.clear()
        _TARGET_ROWS[key] = row
    return row


def _score(query, pos_item, neg_items, pos_ids, neg_ids, downscore, false_neg_score, temperature,
           pos_prob=None, neg_prob=None, fused_loss: bool = False) -> Prediction:
    B, N = query.shape[0], neg_items.shape[0]
    if fused_loss:
        # soft-max cross-entropy against the one-hot target on column 0, straight from the GEMM epilogue: the
        # (B, 1+N) logits (1.07 GB at B = 16 384) are never written.  outputs = (B,3) [max, log-sum-exp, positive logit]
        if dense_engine() == "fp32":
            raise NotImplementedError("fused_loss runs on the tensor-core engine")
        stats = ops.inbatch_softmax_ce(query.contiguous(), pos_item.contiguous(), neg_items.contiguous(), pos_ids=pos_ids,
                                       neg_ids=neg_ids, downscore=downscore, false_neg_score=false_neg_score,
                                       pos_prob=pos_prob, neg_prob=neg_prob, temperature=temperature)
        return Prediction(stats, None, negative_candidate_ids=neg_ids, kind="softmax_ce_stats")
    # (B, 1+N) logits as a view into a (B, 4+Nr) buffer starting at physical column 3: the negatives
    # (logical columns 1..N) then start 16-byte aligned in every row, so the GEMM epilogue can use
    # 128-bit stores; Nr = N rounded up to 4 keeps the row stride a multiple of 16 bytes
    Nr = (N + 3) // 4 * 4
    out = torch.empty((B, 4 + Nr), dtype=torch.float32, device=query.device)[:, 3:4 + N]
    ops.inbatch_scores(query.contiguous(), pos_item.contiguous(), neg_items.contiguous(), out, pos_ids=pos_ids,
                       neg_ids=neg_ids, downscore=downscore, false_neg_score=false_neg_score, pos_prob=pos_prob,
                       neg_prob=neg_prob, temperature=temperature, tensor_cores=dense_engine() != "fp32")
    # targets: one-hot on column 0 (retrieval/base.py:413-422) as a broadcast view — the reference
    # materialises a second (B, 1+N) tensor; nothing downstream needs it resident
    return Prediction(out, _target_row(1 + N, query.device).unsqueeze(0).expand(B, 1 + N), negative_candidate_ids=neg_ids)


class ItemRetrievalScorer(Block):
    """blocks/retrieval/base.py:134-502 (in-batch / sampled negatives mode)."""

    def __init__(self, samplers: Sequence = (), sampling_downscore_false_negatives: bool = True,
                 sampling_downscore_false_negatives_value: float = MIN_FLOAT, item_id_feature_name: str = "item_id",
                 item_domain: str = "item_id", query_name: str = "query", item_name: str = "item",
                 cache_query: bool = False, sampled_softmax_mode: bool = False, store_negative_ids: bool = False,
                 **kwargs):
        super().__init__(unique_name("item_retrieval_scorer"))
        self.samplers = list(samplers) if samplers else [InBatchSampler()]
        self.downscore_false_negatives = sampling_downscore_false_negatives
        self.false_negatives_score = sampling_downscore_false_negatives_value
        self.item_id_feature_name = item_id_feature_name
        self.query_name, self.item_name = query_name, item_name
        self.store_negative_ids = store_negative_ids
        if sampled_softmax_mode or cache_query:
            raise NotImplementedError("sampled_softmax_mode / cache_query are outside the in-batch hot path")

    def _check_input_from_two_tower(self, inputs):
        if set(inputs.keys()) != {self.query_name, self.item_name}:
            raise ValueError(
                f"Wrong input-names, expected: {[self.query_name, self.item_name]} "
                f"but got: {inputs.keys()}"
            )

    def call(self, inputs: Dict[str, torch.Tensor], training: bool = False, testing: bool = False, **kwargs):
        """Inference: (B,1) positive scores (retrieval/base.py:277-281)."""
        if training or testing:
            return inputs
        self._check_input_from_two_tower(inputs)
        q, it = inputs[self.query_name], inputs[self.item_name]
        out = torch.empty((q.shape[0], 1), dtype=torch.float32, device=q.device)
        return ops.rowwise_dot(q, it, out)

    def call_outputs(self, predictions: Dict[str, torch.Tensor], features: TabularData, temperature: float = 1.0,
                     fused_loss: bool = False, **kwargs) -> Prediction:
        """Training / testing logits (retrieval/base.py:283-429); `fused_loss=True`: the cross-entropy statistics
        of those logits instead of the logits (see _score)."""
        assert len(self.samplers) > 0, "At least one sampler is required by ItemRetrievalScorer for negative sampling"
        self._check_input_from_two_tower(predictions)
        q, items = predictions[self.query_name], predictions[self.item_name]
        pos_ids = None
        if self.downscore_false_negatives or self.store_negative_ids:
            if self.item_id_feature_name not in features:
                raise ValueError(f"the item id feature {self.item_id_feature_name!r} is required to "
                                 "downscore false negatives")
            pos_ids = features[self.item_id_feature_name].reshape(-1)
        neg_e, neg_i = [], []
        for s in self.samplers:
            e, i, _ = s.sample(items, pos_ids)
            if e.shape[0] > 0:
                neg_e.append(e)
                neg_i.append(i)
        if not neg_e:
            raise Exception(f"No negative items where sampled from samplers {self.samplers}")
        neg = neg_e[0] if len(neg_e) == 1 else torch.cat(neg_e, dim=0)
        nid = None
        if pos_ids is not None:
            nid = neg_i[0] if len(neg_i) == 1 else torch.cat(neg_i, dim=0)
        return _score(q, items, neg, pos_ids, nid, self.downscore_false_negatives, self.false_negatives_score,
                      temperature, fused_loss=fused_loss)


class ItemRetrievalTask(Block):
    """prediction_tasks/retrieval.py:33-191: ItemRetrievalScorer (+ LogitsTemperatureScaler when
    T != 1, applied only in training/testing — transforms/bias.py:44-52)."""

    def __init__(self, schema: Schema, samplers: Sequence = (), target_name: Optional[str] = None,
                 task_name: Optional[str] = None, post_logits=None, logits_temperature: float = 1.0,
                 cache_query: bool = False, store_negative_ids: bool = False, **kwargs):
        super().__init__(task_name or unique_name("item_retrieval_task"))
        if post_logits is not None:
            raise NotImplementedError("post_logits blocks are outside the hot path")
        ids = schema.select_by_tag(Tags.ITEM_ID).column_names
        if not ids:
            raise ValueError("ItemRetrievalTask needs a column tagged ITEM_ID in the schema")
        self.schema = schema
        self.item_id_feature_name = ids[0]
        self.logits_temperature = float(logits_temperature)
        self.scorer = ItemRetrievalScorer(samplers=samplers, item_id_feature_name=self.item_id_feature_name,
                                          cache_query=cache_query, store_negative_ids=store_negative_ids)
        self.target_name = target_name

    def call(self, inputs, features: Optional[TabularData] = None, training: bool = False, testing: bool = False,
             fused_loss: bool = False, **kwargs):
        if training or testing:
            return self.scorer.call_outputs(inputs, features, temperature=self.logits_temperature, fused_loss=fused_loss)
        return self.scorer(inputs)


class ContrastiveOutput(Block):
    """outputs/contrastive.py:47-453 (DotProduct to_call, in-batch / provided negatives).

    call({query_name: (B,D), candidate_name: (B,D)}, candidate_ids, training|testing) -> Prediction
    with logits (B, 1+N); inference -> (B,1) row-wise dot (outputs/base.py:291-322)."""

    def __init__(self, to_call=None, negative_samplers="in-batch", target_name: Optional[str] = None,
                 logits_temperature: float = 1.0, name: Optional[str] = None, downscore_false_negatives: bool = True,
                 false_negative_score: float = MIN_FLOAT, query_name: str = "query", candidate_name: str = "candidate",
                 store_negative_ids: bool = False, logq_sampling_correction: Optional[bool] = False, **kwargs):
        super().__init__(name or unique_name("contrastive_output"))
        if isinstance(negative_samplers, (str, InBatchSampler)):
            negative_samplers = [negative_samplers]
        self.negative_samplers = [InBatchSampler() if s in ("in-batch", "in_batch") else s for s in negative_samplers]
        if not self.negative_samplers:
            raise ValueError("At least one negative sampler is required")
        self.logits_temperature = float(logits_temperature)
        self.downscore_false_negatives = downscore_false_negatives
        self.false_negative_score = false_negative_score
        self.query_name, self.candidate_name = query_name, candidate_name
        self.store_negative_ids = store_negative_ids
        self.logq_sampling_correction = logq_sampling_correction

    def call(self, inputs: Dict[str, torch.Tensor], candidate_ids: Optional[torch.Tensor] = None,
             training: bool = False, testing: bool = False, sampling_probs: Optional[torch.Tensor] = None,
             fused_loss: bool = False, **kwargs):
        q, c = inputs[self.query_name], inputs[self.candidate_name]
        if not (training or testing):
            out = torch.empty((q.shape[0], 1), dtype=torch.float32, device=q.device)
            return ops.rowwise_dot(q, c, out)
        if self.downscore_false_negatives and candidate_ids is None:
            raise ValueError("candidate ids are required to downscore false negatives")
        ids = None if candidate_ids is None else candidate_ids.reshape(-1)
        neg_e, neg_i, neg_p = [], [], []
        for s in self.negative_samplers:
            e, i, p = s.sample(c, ids)
            neg_e.append(e)
            neg_i.append(i)
            neg_p.append(p)
        neg = neg_e[0] if len(neg_e) == 1 else torch.cat(neg_e, dim=0)
        nid = None if ids is None else (neg_i[0] if len(neg_i) == 1 else torch.cat(neg_i, dim=0))
        pos_prob = neg_prob = None
        if self.logq_sampling_correction:
            if sampling_probs is None:
                raise ValueError("logq_sampling_correction needs `sampling_probs` (probability table over item ids)")
            # positive / negative sampling probabilities are looked up by id (contrastive.py:309-319)
            pos_prob = sampling_probs[ids.long()].contiguous()
            neg_prob = sampling_probs[nid.long()].contiguous()
        return _score(q, c, neg, ids, nid, self.downscore_false_negatives, self.false_negative_score,
                      self.logits_temperature, pos_prob, neg_prob, fused_loss=fused_loss)


# ------------------------------------------------------------------------------------------------
# full-catalog scoring (a14)
# ------------------------------------------------------------------------------------------------
class CategoricalOutput(Block):
    """outputs/classification.py:127-216 with the weight-tied `EmbeddingTablePrediction` to_call
    (:311-382): logits = x @ E^T + bias over the whole item table E (N_I, D).

    call(x)                      -> (B, N_I) logits, materialised (what the reference returns; only
                                    sensible for small catalogs)
    softmax_ce_stats(x, targets) -> (B, 3) [row max, log-sum-exp, logit[target]] — everything
                                    CategoricalCrossEntropy(from_logits=True) needs
                                    (losses/listwise.py:38-50): loss = lse - logit[target]
    top_k(x, k)                  -> (scores, ids) in tf.math.top_k order (outputs/topk.py:221-223)
    The last two stream the catalog through one tcgen05 GEMM without ever writing (B, N_I)."""

    def __init__(self, to_call, logits_temperature: float = 1.0, use_bias: bool = True, name: Optional[str] = None, **kwargs):
        from .inputs import EmbeddingTable

        super().__init__(name or unique_name("categorical_output"))
        if not isinstance(to_call, EmbeddingTable):
            raise NotImplementedError("CategoricalOutput(to_call=...) supports a weight-tied EmbeddingTable")
        self.table = to_call
        self.num_classes = to_call.input_dim
        self.logits_temperature = float(logits_temperature)
        self.use_bias = use_bias
        self.bias: Optional[torch.Tensor] = None
        self._e_split: Optional[torch.Tensor] = None
        self._w_split: Optional[torch.Tensor] = None

    def build(self, device=None):
        self.table.build(device)
        if self.use_bias and self.bias is None:
            self.bias = torch.zeros(self.num_classes, dtype=torch.float32, device=self.table.table.device)
        self.built = True
        return self

    def weights(self):
        out = {"embeddings": self.table.embeddings}
        if self.bias is not None:
            out["bias"] = self.bias
        return out

    _TRANSIENT = {"_e_split": None, "_w_split": None}

    def refresh(self) -> None:
        """Drop the cached split-bf16 copies (call after changing the table)."""
        from .core import bump_weights_version

        self._e_split = self._w_split = None
        bump_weights_version()

    _weights_changed = refresh

    def _catalog_split(self) -> torch.Tensor:
        if self._e_split is None:
            self._e_split = ops.split_rows(self.table.embeddings)  # (N_I, 2*Kp), once per catalog
        return self._e_split

    def call(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        self.build(x.device)
        if self._w_split is None:
            self._w_split = ops.split_weights(self.table.embeddings.t().contiguous())
        out = torch.empty((x.shape[0], self.num_classes), dtype=torch.float32, device=x.device)
        ops.dense_tc(ops.split_rows(x), x.shape[1], self._w_split, self.num_classes, self.bias, "linear", out_f32=out)
        if self.logits_temperature != 1.0:
            raise NotImplementedError("materialised logits with a temperature: use softmax_ce_stats / top_k on x / T")
        return out

    def softmax_ce_stats(self, x: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        self.build(x.device)
        stats, _, _ = ops.catalog_score(x, self._catalog_split(), self.num_classes, bias=self.bias, targets=targets, k=0)
        return stats

    def top_k(self, x: torch.Tensor, k: int):
        self.build(x.device)
        _, scores, ids = ops.catalog_score(x, self._catalog_split(), self.num_classes, bias=self.bias, k=k, want_stats=False)
        return scores, ids
