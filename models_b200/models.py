"""Model factories behind the reference constructors: DLRMModel, DCNModel, TwoTowerModel.

Reference: merlin/models/tf/models/ranking.py:23-168, models/retrieval.py:106-203,
models/base.py:1805-1854 (Model.call protocol), outputs/classification.py:72-123 (BinaryOutput),
prediction_tasks/classification.py:59-116 (BinaryClassificationTask).  Only construction and the
forward call are in scope — fit/compile/optimizers/metrics are not.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import ops
from .blocks import (FM, MLP, CrossBlock, CrossBlockSeq, DLRM, DLRMBlock, FMBlock, MLPBlock, _Dense, dense_engine,
                     run_dense_chain)
from .core import Block, Prediction, TabularData, batch_size_of, default_device, to_device, unique_name
from .inputs import EmbeddingOptions, EmbeddingsBlock, InputBlockV2
from .retrieval import ItemRetrievalTask, TwoTowerBlock
from .schema import Schema, Tags


class BinaryOutput(Block):
    """outputs/classification.py:72-123: Dense(1, activation="sigmoid") on the body output."""

    def __init__(self, target: Optional[Union[str, object]] = None, name: Optional[str] = None, **kwargs):
        tname = getattr(target, "name", target)
        super().__init__(name or (f"{tname}/binary_output" if tname else unique_name("binary_output")))
        self.target = tname
        self.to_call = _Dense(1, activation="sigmoid", name=f"{self.name}/dense")

    def build(self, width: Optional[int] = None, device=None):
        self.to_call.build(width, device)
        self.built = True
        return self

    def weights(self):
        return {f"dense/{k}": v for k, v in self.to_call.weights().items()}

    def call(self, inputs: torch.Tensor, **kwargs) -> torch.Tensor:
        return self.to_call(inputs)


class BinaryClassificationTask(BinaryOutput):
    """prediction_tasks/classification.py:59-116 (v1): Dense(1, linear) followed by an fp32 sigmoid —
    the same function as BinaryOutput, fused in the layer epilogue here."""

    def __init__(self, target: Optional[str] = None, task_name: Optional[str] = None, **kwargs):
        super().__init__(target, name=task_name or (f"{target}/binary_classification_task" if target else None))


def parse_prediction_blocks(schema: Schema, prediction_blocks=None) -> Block:
    """models/utils.py:12-31 / outputs/block.py:79-128: default = BinaryOutput for the (single)
    binary-classification target of the schema."""
    if prediction_blocks is None:
        targets = schema.select_by_tag(Tags.BINARY_CLASSIFICATION)
        if not len(targets):
            targets = schema.select_by_tag(Tags.TARGET)
        if not len(targets):
            raise ValueError("The schema has no target column: pass `prediction_tasks` explicitly")
        if len(targets) > 1:
            binary = [c.name for c in targets if c.has_tag(Tags.BINARY_CLASSIFICATION)]
            if len(binary) != 1:
                raise NotImplementedError("multi-task outputs are outside the hot path; pass one BinaryOutput")
            return BinaryOutput(binary[0])
        return BinaryOutput(targets.first.name)
    if isinstance(prediction_blocks, (list, tuple)):
        if len(prediction_blocks) != 1:
            raise NotImplementedError("multi-task outputs are outside the hot path")
        prediction_blocks = prediction_blocks[0]
    if not isinstance(prediction_blocks, Block):
        raise ValueError(f"Unsupported prediction task {prediction_blocks!r}")
    return prediction_blocks


def expected_input_columns(schema: Schema) -> List[str]:
    """models/base.py:1730-1749: non-target columns; list columns as `__values` + `__offsets`."""
    cols = []
    for c in schema.excluding_by_tag(Tags.TARGET):
        if c.is_list and c.is_ragged:
            cols += [c.name + "__values", c.name + "__offsets"]
        else:
            cols.append(c.name)
    return cols


class Model(Block):
    """models/base.py:1621-2245: model(inputs, targets=None, training=False, testing=False) with `inputs` a dict keyed by
    schema column names; `compile(optimizer)` / `fit` / `train_step` for the DLRM path (models_b200/train.py)."""

    def __init__(self, body: Block, prediction: Block, schema: Schema):
        super().__init__(unique_name("model"))
        self.body = body
        self.prediction = prediction
        self.schema = schema
        self._pinned: Dict[str, torch.Tensor] = {}

    _TRANSIENT = {"_pinned": {}, "_trainer": None}

    @property
    def blocks(self) -> List[Block]:
        return [self.body, self.prediction]

    # -- checkpoint boundary (models_b200/io.py; reference: models/base.py:1687-1728) -----------
    def save(self, export_path, include_optimizer: bool = True, save_traces: bool = True) -> None:
        """Variables (Keras layouts, one .npy each), block structure and `.merlin` schema metadata.
        `include_optimizer` / `save_traces` are accepted for signature parity (forward path only)."""
        from . import io as _io

        _io.save_model(self, export_path)

    @classmethod
    def load(cls, export_path, device=None) -> "Model":
        from . import io as _io

        return _io.load_model(export_path, device)

    def load_weights(self, source, name_map=None, strict: bool = True):
        """Assign variables from an export directory or a {name: array} mapping (e.g. a Keras checkpoint
        exported as `{v.name: v.numpy()}`); see io.load_weights."""
        from . import io as _io

        return _io.load_weights(self, source, name_map=name_map, strict=strict)

    def state_dict(self) -> Dict[str, np.ndarray]:
        from . import io as _io

        return _io.state_dict(self)

    def output_schema(self) -> Schema:
        """One float column per prediction task (what `get_output_schema` records for the reference)."""
        from .schema import ColumnSchema

        target = getattr(self.prediction, "target", None) or getattr(self.prediction, "target_name", None)
        name = f"{target}/{self.prediction.name}" if target else self.prediction.name
        return Schema([ColumnSchema(name, dtype="float32")])

    def weights(self):
        out = {f"body/{k}": v for k, v in self.body.weights().items()}
        out.update({f"prediction/{k}": v for k, v in self.prediction.weights().items()})
        return out

    def _check_inputs(self, inputs: TabularData) -> None:
        if not isinstance(inputs, dict):
            raise ValueError(f"Model inputs must be a dict of features, got {type(inputs).__name__}")
        missing = [c for c in self.input_columns() if c not in inputs]
        if missing:
            raise ValueError(f"Missing input features: {missing}")

    def input_columns(self) -> List[str]:
        return expected_input_columns(self.schema)

    def id_bytes(self) -> Dict[str, int]:
        """Narrowest id width (1, 2 or 3 bytes) each scalar categorical input column can travel at from the
        host: its table has <= 2^8 / 2^16 / 2^24 rows.  `HostBatch.like(batch, names, id_bytes=...)` packs
        the pinned batch accordingly (the loader hand-off is PCIe-bound: a Criteo sample shrinks from 156 to
        104 bytes); the fused lookup kernel reads packed ids natively, other paths widen them on the device."""
        out: Dict[str, int] = {}
        for emb in self.embedding_blocks():
            for f, table in emb.feature_to_table.items():
                col = self.schema.get(f)
                if col is None or col.is_list:
                    continue
                rows = table.input_dim
                if rows <= (1 << 8):
                    out[f] = 1
                elif rows <= (1 << 16):
                    out[f] = 2
                elif rows <= (1 << 24):
                    out[f] = 3
        return out

    def call(self, inputs: TabularData, targets=None, training: bool = False, testing: bool = False, **kwargs):
        self._check_inputs(inputs)
        x = self.body(inputs, training=training, testing=testing)
        return self.prediction(x, features=inputs, targets=targets, training=training, testing=testing)

    # -- training (models_b200/train.py; reference: models/base.py:1121-1231) ------------------
    def _compile_training(self, optimizer, loss=None) -> None:
        from .train import get_optimizer

        if loss not in (None, "binary_crossentropy"):
            raise NotImplementedError(f"loss {loss!r}: only the BinaryOutput default (binary cross-entropy) is implemented")
        self.optimizer = get_optimizer(optimizer)
        self._trainer = None

    def trainer(self, batch_size: int, group=None):
        """The static-buffer training engine for batches of (up to) `batch_size` samples (train.DLRMTrainer)."""
        from .train import DLRMTrainer

        if getattr(self, "optimizer", None) is None:
            raise RuntimeError("compile() the model with an optimizer before training it")
        tr = getattr(self, "_trainer", None)
        if tr is not None:
            if group is not None and tr.group is not group:
                raise ValueError("this model already trains with another process group")
            if tr.B < batch_size:
                raise NotImplementedError("the training batch size grew after the first step: create the engine for the largest "
                                          "batch first (model.trainer(batch_size) before the first train_step)")
            return tr
        tr = self._trainer = DLRMTrainer(self, self.optimizer, batch_size, group=group)
        return tr

    def train_step(self, data) -> Dict[str, torch.Tensor]:
        """One optimizer step on `data` = (inputs, targets[, sample_weight]); returns the reference's step metrics
        {"loss", "loss_batch", "regularization_loss"} as device scalars (models/base.py:1121-1177).  The loss scalars are
        views of the engine's loss buffer: valid until the next step (clone to keep)."""
        if getattr(self, "optimizer", None) is None:
            raise RuntimeError("compile() the model with an optimizer before training it")
        if not isinstance(data, (tuple, list)) or len(data) < 2:
            raise ValueError("train_step expects (inputs, targets) or (inputs, targets, sample_weight)")
        x, y = data[0], data[1]
        sw = data[2] if len(data) > 2 else None
        if isinstance(y, dict):
            if len(y) != 1:
                raise NotImplementedError("multi-task targets are not implemented in the training step")
            y = next(iter(y.values()))
        if y is None:
            raise ValueError("train_step needs targets")
        self._check_inputs(x)
        tr = self.trainer(batch_size_of(x))
        loss = tr.step(x, y, sw)
        return {"loss": loss[0], "loss_batch": loss[0], "regularization_loss": torch.zeros((), device=loss.device)}

    def fit(self, x=None, y=None, batch_size: Optional[int] = None, epochs: int = 1, steps_per_epoch: Optional[int] = None,
            verbose: int = 0, **kwargs):
        """Keras `fit` over a models_b200.Loader (or any iterable of (inputs, targets)): returns a History-like object
        whose `.history["loss"]` holds the mean batch loss of every epoch."""
        if x is None:
            raise ValueError("fit needs a loader / iterable of (inputs, targets) batches")
        if y is not None:
            raise NotImplementedError("fit(x, y): pass a Loader or an iterable of (inputs, targets) batches")
        bs = batch_size or getattr(x, "batch_size", None)
        history = {"loss": []}
        for _ in range(int(epochs)):
            total, n = None, 0
            for inputs, targets in x:
                if getattr(self, "_trainer", None) is None and bs:
                    self._check_inputs(inputs)
                    self.trainer(int(bs))
                m = self.train_step((inputs, targets))
                total = m["loss_batch"].clone() if total is None else total + m["loss_batch"]
                n += 1
                if steps_per_epoch and n >= steps_per_epoch:
                    break
            if n == 0:
                raise ValueError("fit: the loader produced no batches")
            history["loss"].append(float(total.item()) / n)
            self._trainer.check_indices()
        from .train import History

        self.history = History(history)
        return self.history

    # -- CUDA-graph runtime (models_b200/graph.py) ---------------------------------------------
    def embedding_blocks(self) -> List[EmbeddingsBlock]:
        """Every EmbeddingsBlock of the model (they share one out-of-range index counter)."""
        found: List[EmbeddingsBlock] = []

        def walk(o, depth=0):
            if isinstance(o, EmbeddingsBlock):
                if o not in found:
                    found.append(o)
                return
            if depth > 6 or not isinstance(o, Block):
                return
            for v in vars(o).values():
                if isinstance(v, Block):
                    walk(v, depth + 1)
                elif isinstance(v, (list, tuple)):
                    for e in v:
                        walk(e, depth + 1)

        walk(self)
        return found

    def index_error_counter(self, device) -> Optional[torch.Tensor]:
        blocks = [b for b in self.embedding_blocks() if b.check_indices]
        if not blocks:
            return None
        c = blocks[0].counter(device)
        for b in blocks[1:]:
            b.oob_counter = c
        return c

    def defer_index_check(self, flag: bool) -> None:
        for b in self.embedding_blocks():
            b.defer_check = bool(flag)

    def compile(self, example: Union[Dict[str, np.ndarray], "HostBatch", str, None] = None, *, optimizer=None, loss=None,
                metrics=None, run_eagerly=None, **call_kwargs):
        """Two uses, told apart by the argument:

        * `compile(optimizer="adam")` / `compile("adagrad")` / `compile(optimizer=mm.Adagrad(0.01))` — Keras `compile`
          (models/base.py: the reference's models are compiled before `fit`): picks the optimizer of the training step
          (models_b200/train.py).  The loss is the prediction task's default (binary cross-entropy for BinaryOutput);
          `metrics` / `run_eagerly` are accepted for signature parity.
        * `compile(example_batch, **call_kwargs)` — capture this model's forward for `example`'s batch layout into a CUDA
          graph; the result maps a packed pinned HostBatch to pinned host predictions with one H2D, one graph launch and
          one D2H (models_b200/graph.py)."""
        from .graph import CompiledForward, HostBatch
        from .train import Optimizer

        if optimizer is not None or isinstance(example, (str, Optimizer)) or example is None:
            if example is not None and optimizer is not None:
                raise ValueError("compile(): pass either an example batch (graph capture) or an optimizer (training)")
            return self._compile_training(optimizer if optimizer is not None else (example or "adam"), loss)
        if not isinstance(example, HostBatch):
            example = HostBatch.like(example, self.input_columns())
        return CompiledForward(self, example, **call_kwargs)

    def pipeline(self, example: Union[Dict[str, np.ndarray], "HostBatch"], depth: int = 2, **call_kwargs) -> "PipelinedForward":
        """`depth` graph instances on separate streams so that the H2D copy of batch i+1 overlaps the
        forward of batch i (models_b200/graph.py)."""
        from .graph import HostBatch, PipelinedForward

        if not isinstance(example, HostBatch):
            example = HostBatch.like(example, self.input_columns())
        return PipelinedForward(self, example, depth=depth, **call_kwargs)

    # -- host-buffer entry point (the e2e path of bench.py) -----------------------------------
    def forward_host(self, batch: Dict[str, np.ndarray], stream: Optional[torch.cuda.Stream] = None, **kwargs):
        """Host numpy batch -> pinned staging -> H2D -> forward -> D2H of the predictions."""
        dev = default_device()
        dev_inputs = {}
        for k in self.input_columns():
            src = torch.from_numpy(np.ascontiguousarray(batch[k]))
            pin = self._pinned.get(k)
            if pin is None or pin.shape != src.shape or pin.dtype != src.dtype:
                pin = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
                self._pinned[k] = pin
            pin.copy_(src)
            dev_inputs[k] = pin.to(dev, non_blocking=True)
        out = self.call(dev_inputs, **kwargs)
        pred = out.outputs if isinstance(out, Prediction) else out
        return pred.cpu()


class RankingModel(Model):
    """DLRM / DCN: body -> (B, h) -> BinaryOutput (B,1)."""

    def build(self, device=None):
        self.body.build(device)
        self.prediction.build(self.body_width(), device)
        self.built = True
        return self

    def _all_onehot(self, inputs: TabularData) -> bool:
        from .core import get_feature

        emb = self.body.embeddings
        return all(emb.feature_to_table[f].lookup_kind(get_feature(inputs, f)) == "onehot" for f in emb.feature_names)

    def body_width(self) -> int:
        if isinstance(self.body, DLRM):
            if self.body.top_block is not None:
                return self.body.top_block.dense_layers[-1].units
            return self.body.output_width_before_top()
        return self.body.output_width()

    def call(self, inputs: TabularData, targets=None, training: bool = False, testing: bool = False, **kwargs):
        self._check_inputs(inputs)
        if not self.built:
            self.build(next(iter(inputs.values())).device)
        if isinstance(self.body, DLRM) and self.body.top_block is not None:
            # top MLP + output layer as ONE dense chain (no fp32 round trip between them)
            # top MLP (+ its normalizations, folded) + the output Dense as one chain
            layers, tail = self.body.top_block.chain([self.prediction.to_call])
            assert tail is None
            if dense_engine() != "fp32" and self.body.can_emit_split():
                # production path: bottom vector and table rows in the interaction kernel's operand format when the
                # mirrors are on (no bf16 split inside the hot loop); split-bf16 row straight into the top tower
                op = self.body.use_operand_rows() and self._all_onehot(inputs)
                bottom = self.body.bottom_forward(inputs, operand_out=op)
                a = self.body.interaction_forward(inputs, bottom, as_split=True, operand_rows=op)
                return run_dense_chain(None, layers, a_split=a, K=self.body.output_width_before_top())
            bottom = self.body.bottom_forward(inputs)
            x = self.body.interaction_forward(inputs, bottom)
            return run_dense_chain(x, layers)
        if isinstance(self.body, DeepFMBody):
            return self.body.forward(inputs, out_layer=self.prediction.to_call)
        if isinstance(self.body, DCNBody) and self.body.stacked:
            x = self.body.cross(self.body.input_block(inputs))
            layers, tail = self.body.deep.chain([self.prediction.to_call])
            assert tail is None
            return run_dense_chain(x, layers)
        x = self.body(inputs, training=training)
        return self.prediction(x)


def DLRMModel(schema: Schema, *, embeddings: Optional[EmbeddingsBlock] = None, embedding_dim: Optional[int] = None,
              embedding_options: Optional[EmbeddingOptions] = None, bottom_block: Optional[MLP] = None,
              top_block: Optional[MLP] = None, prediction_tasks=None) -> RankingModel:
    """models/ranking.py:23-92."""
    prediction = parse_prediction_blocks(schema, prediction_tasks)
    body = DLRMBlock(schema, embedding_dim=embedding_dim, embedding_options=embedding_options, embeddings=embeddings,
                     bottom_block=bottom_block, top_block=top_block)
    return RankingModel(body, prediction, schema)


class DCNBody(Block):
    """input_block.connect(CrossBlock(depth), deep_block) (stacked) or connect_branch(..., "concat")."""

    def __init__(self, input_block: InputBlockV2, cross: CrossBlockSeq, deep: MLP, stacked: bool):
        super().__init__(unique_name("dcn_body"))
        self.input_block, self.cross, self.deep, self.stacked = input_block, cross, deep, stacked

    def build(self, device=None):
        self.input_block.build(device)
        _, _, d = self.input_block.layout()
        for l in self.cross.cross_layers:
            l.build(d, device)
        self.deep.build_from_width(d, device)
        self.built = True
        return self

    def output_width(self) -> int:
        _, _, d = self.input_block.layout()
        last = self.deep.dense_layers[-1].units
        return last if self.stacked else d + last

    def weights(self):
        out = {f"input/{k}": v for k, v in self.input_block.weights().items()}
        for l in self.cross.cross_layers:
            out.update({f"{l.name}/{k}": v for k, v in l.weights().items()})
        out.update({f"deep/{k}": v for k, v in self.deep.weights().items()})
        return out

    def call(self, inputs: TabularData, **kwargs):
        x0 = self.input_block(inputs)
        c = self.cross(x0)
        if self.stacked:
            return self.deep(c)
        d = self.deep(x0)
        out = torch.empty((x0.shape[0], c.shape[1] + d.shape[1]), dtype=torch.float32, device=x0.device)
        return ops.concat_columns([c, d] if self.branch_order() == ("cross", "deep") else [d, c], out)

    def branch_order(self):
        """Column order of the non-stacked concat.  The reference's `connect_branch(CrossBlock(depth), deep_block,
        aggregation="concat")` builds a ParallelBlock keyed by the two layers' auto names and ConcatFeatures sorts the
        keys (core/combinators.py, core/aggregation.py:54-66): both are `sequential_block[_N]`, the deep block is created
        first, so the order is normally [deep | cross] — by STRING comparison of the names (`..._10` < `..._9`).  The same
        rule is applied to this package's Keras-style names; set `body.concat_order = ("cross", "deep")` (or the reverse)
        to pin the layout of an imported checkpoint's output kernel explicitly."""
        forced = getattr(self, "concat_order", None)
        if forced is not None:
            if tuple(forced) not in (("cross", "deep"), ("deep", "cross")):
                raise ValueError("concat_order must be ('cross', 'deep') or ('deep', 'cross')")
            return tuple(forced)
        return ("cross", "deep") if self.cross.name < self.deep.name else ("deep", "cross")


class DeepFMBody(Block):
    """ParallelBlock({"fm": FMBlock, "deep": input_block -> deep_block -> deep_logit_block}, "element-wise-sum")
    (models/ranking.py:250-274): (B, 1).  The deep tower is the usual concat + dense chain; the FM pairwise term, the wide
    part, the sum with the deep logit and (from RankingModel) the output layer are ONE kernel (ops.deepfm_head)."""

    def __init__(self, input_block: InputBlockV2, fm: FM, deep: MLP, deep_logit: MLP):
        super().__init__(unique_name("deepfm_body"))
        self.input_block, self.fm, self.deep, self.deep_logit = input_block, fm, deep, deep_logit
        if deep.has_normalization or deep_logit.has_normalization:
            raise NotImplementedError("DeepFMModel: normalization inside deep_block / deep_logit_block is not implemented")
        if deep_logit.dense_layers[-1].units != 1:
            raise ValueError("The last dimension of deep_logit_block needs to be 1")

    def build(self, device=None):
        self.input_block.build(device)
        _, _, d = self.input_block.layout()
        self.deep.build_from_width(d, device)
        self.deep_logit.build_from_width(self.deep.dense_layers[-1].units, device)
        self.fm.build(device)
        self.built = True
        return self

    def output_width(self) -> int:
        return 1

    def weights(self):
        out = {f"input/{k}": v for k, v in self.input_block.weights().items()}
        out.update({f"fm/{k}": v for k, v in self.fm.weights().items()})
        out.update({f"deep/{k}": v for k, v in self.deep.weights().items()})
        out.update({f"deep_logit/{k}": v for k, v in self.deep_logit.weights().items()})
        return out

    def forward(self, inputs: TabularData, out_layer: Optional[_Dense] = None) -> torch.Tensor:
        if not self.built:
            self.build(next(iter(inputs.values())).device)
        x0 = self.input_block(inputs)
        deep = run_dense_chain(x0, self.deep.dense_layers + self.deep_logit.dense_layers)
        return self.fm.head(inputs, addend=deep, out_layer=out_layer)

    def call(self, inputs: TabularData, **kwargs) -> torch.Tensor:
        return self.forward(inputs)


def DeepFMModel(schema: Schema, embedding_dim: Optional[int] = None, deep_block: Optional[MLP] = None,
                input_block: Optional[InputBlockV2] = None, wide_input_block=None, wide_logit_block=None,
                deep_logit_block: Optional[MLP] = None, prediction_tasks=None, **kwargs) -> RankingModel:
    """models/ranking.py:171-279: sigmoid(Dense(1)(FM(x) + deep(x))) with FM = wide + pairwise (blocks.FM), deep =
    MLP over the concatenated embeddings and continuous features followed by MLPBlock([1], linear).  Defaults as the
    reference: deep_block = MLPBlock([64]); one embedding dimension for every categorical feature (`embedding_dim`)."""
    if input_block is None:
        if embedding_dim is None:
            raise ValueError("DeepFMModel needs `embedding_dim` (the FM term stacks the embeddings: one dimension for all tables)")
        from .inputs import Embeddings

        cat = schema.select_by_tag(Tags.CATEGORICAL).excluding_by_tag(Tags.TARGET)
        input_block = InputBlockV2(schema, categorical=Embeddings(cat, dim=embedding_dim), **kwargs)
    fm = FMBlock(schema, fm_input_block=input_block, wide_input_block=wide_input_block, wide_logit_block=wide_logit_block)
    deep_block = deep_block if deep_block is not None else MLPBlock([64])
    deep_logit_block = deep_logit_block if deep_logit_block is not None else MLPBlock([1], activation="linear", use_bias=True)
    prediction = parse_prediction_blocks(schema, prediction_tasks)
    return RankingModel(DeepFMBody(input_block, fm, deep_block, deep_logit_block), prediction, schema)


def DCNModel(schema: Schema, depth: int, deep_block: Optional[MLP] = None, stacked: bool = True,
             input_block: Optional[InputBlockV2] = None, prediction_tasks=None, **kwargs) -> RankingModel:
    """models/ranking.py:95-168 (default deep_block = MLPBlock([512, 256]))."""
    deep_block = deep_block if deep_block is not None else MLPBlock([512, 256])
    input_block = input_block or InputBlockV2(schema, **kwargs)
    prediction = parse_prediction_blocks(schema, prediction_tasks)
    body = DCNBody(input_block, CrossBlock(depth), deep_block, stacked)
    return RankingModel(body, prediction, schema)


def _brute_force(k: int):
    from .topk import BruteForce

    return BruteForce(k=k)


class RetrievalModel(Model):
    """models/base.py:2259-2489, forward only."""

    _TRANSIENT = {"pre_eval_topk": None}

    def build(self, device=None):
        self.body.build(device)
        self.built = True
        return self

    def input_columns(self) -> List[str]:
        tb: TwoTowerBlock = self.body
        used = Schema(list(tb.query.inputs.schema) + [c for c in tb.item.inputs.schema
                                                      if c.name not in tb.query.inputs.schema])
        return expected_input_columns(used)

    def call(self, inputs: TabularData, targets=None, training: bool = False, testing: bool = False, **kwargs):
        self._check_inputs(inputs)
        if not self.built:
            self.build(next(iter(inputs.values())).device)
        emb = self.body(inputs, training=False)
        # fused_loss=True (training/testing): Prediction.outputs = (B,3) [max, log-sum-exp, positive logit] of the in-batch
        # logits — loss = outputs[:,1] - outputs[:,2] — instead of the (B, 1+B) logits themselves
        return self.prediction(emb, features=inputs, training=training, testing=testing, **kwargs)

    # -- top-k retrieval / evaluation (models/base.py:2266-2489) ---------------------
    @property
    def retrieval_block(self) -> TwoTowerBlock:
        return self.body

    def query_encoder(self) -> Block:
        """Query tower (+ the model's `post`, e.g. L2 normalisation) as a feature-dict -> (B, D) block."""
        from .topk import TowerEncoder

        return TowerEncoder(self.body.query, self.body.post)

    def candidate_encoder(self) -> Block:
        from .topk import TowerEncoder

        return TowerEncoder(self.body.item, self.body.post)

    def _item_id_column(self) -> str:
        tagged = self.schema.select_by_tag(Tags.ITEM_ID)
        if not tagged:
            raise ValueError("the schema has no column tagged ITEM_ID")
        return tagged.first.name

    def query_embeddings(self, data: Dict[str, np.ndarray], batch_size: int = 65536, query_id: Optional[str] = None):
        """models/base.py:2354-2385: (ids, embeddings) of the query tower over `data`."""
        from .topk import encode_rows

        query_id = query_id or (self.schema.select_by_tag(Tags.USER_ID).first.name if self.schema.select_by_tag(Tags.USER_ID) else None)
        return encode_rows(self.query_encoder(), data, query_id, batch_size)

    def item_embeddings(self, data: Dict[str, np.ndarray], batch_size: int = 65536, item_id: Optional[str] = None):
        """models/base.py:2387-2418: (ids, embeddings) of the item tower over `data`."""
        from .topk import encode_rows

        return encode_rows(self.candidate_encoder(), data, item_id or self._item_id_column(), batch_size)

    def to_top_k_encoder(self, candidates, candidate_id: Optional[str] = None, k: int = 10, batch_size: int = 65536,
                         **kwargs):
        """models/base.py `to_top_k_encoder`: query tower -> brute-force top-k over `candidates` — raw item
        features (dict, encoded by the item tower) or precomputed (ids, embeddings) / a DataFrame indexed by id."""
        from .topk import TopKEncoder

        enc = TopKEncoder(self.query_encoder(), candidates=None if isinstance(candidates, dict) else candidates,
                          candidate_encoder=self.candidate_encoder(), k=k, target=self._item_id_column(),
                          topk_layer=kwargs.pop("topk_layer", None) or _brute_force(k), **kwargs)
        if isinstance(candidates, dict):
            enc.index_candidates(candidates, candidate_id or self._item_id_column(), batch_size)
        return enc

    def evaluate(self, x, item_corpus=None, metrics=None, batch_size: int = 65536, return_dict: bool = True, **kwargs):
        """models/base.py:2266-2351.  `x`: one feature-dict batch or an iterable of them (host arrays or device
        tensors).  With `item_corpus` (a TopKIndexBlock, or the item features of the corpus as a dict: deduplicated
        by item id and encoded by the item tower) every query is ranked against the whole corpus by the fused
        score + top-k kernel; without it the batch's own items are the candidates (in-batch evaluation)."""
        from .topk import (NDCGAt, RecallAt, TopKIndexBlock, evaluate_topk, unique_rows_by_features)

        metrics = list(metrics) if metrics else [RecallAt(10), NDCGAt(10)]
        kmax = max(m.k for m in metrics)
        item_id = self._item_id_column()
        if item_corpus is not None:
            if isinstance(item_corpus, TopKIndexBlock):
                index = item_corpus
                if index._k < kmax:
                    raise ValueError(f"the index returns {index._k} candidates, the metrics need {kmax}")
            elif isinstance(item_corpus, dict):
                corpus = unique_rows_by_features(item_corpus, item_id)
                if not self.built:
                    self.build(default_device())
                index = TopKIndexBlock.from_block(self.candidate_encoder(), corpus, k=kmax, id_column=item_id,
                                                  batch_size=batch_size)
            else:
                raise ValueError(f"`item_corpus` must be either a `TopKIndexBlock` or a dict of item features. Got {type(item_corpus)}")
            self.pre_eval_topk = index
            q = self.query_encoder()

            def predict(b):
                if not self.built:
                    self.build(next(iter(b.values())).device)
                return index.call_outputs(b[item_id].reshape(-1), q(b))
        else:
            def predict(b):
                out = self(b, testing=True)
                k = min(kmax, out.outputs.shape[1])
                scores, order = torch.topk(out.outputs, k, dim=1)
                return Prediction(scores, torch.gather(out.targets, 1, order))
        return evaluate_topk(predict, x, metrics)


class RetrievalModelV2(Model):
    """models/base.py RetrievalModelV2 (forward): query Encoder, candidate Encoder, one output layer.
    Inference: output({"query": q, "candidate": c}) -> (B,1); training/testing: [positive | negatives] logits from the
    output's samplers (or their soft-max CE statistics with fused_loss=True)."""

    def __init__(self, query, candidate, output, schema: Optional[Schema] = None, candidate_id_tag=Tags.ITEM_ID):
        from .retrieval import Encoder

        if schema is None:
            cols = list(query.schema) + [c for c in candidate.schema if c.name not in query.schema]
            schema = Schema(cols)
        super().__init__(query, output, schema)
        self.query_encoder_block, self.candidate_encoder_block = query, candidate
        ids = candidate.schema.select_by_tag(candidate_id_tag).column_names
        if not ids:
            raise ValueError(f"the candidate tower has no column tagged {candidate_id_tag}")
        self.candidate_id_name = ids[0]

    @property
    def blocks(self) -> List[Block]:
        return [self.query_encoder_block, self.candidate_encoder_block, self.prediction]

    def weights(self):
        out = {f"query/{k}": v for k, v in self.query_encoder_block.weights().items()}
        out.update({f"candidate/{k}": v for k, v in self.candidate_encoder_block.weights().items()})
        return out

    def build(self, device=None):
        self.query_encoder_block.build(device)
        self.candidate_encoder_block.build(device)
        self.built = True
        return self

    def input_columns(self) -> List[str]:
        return expected_input_columns(self.schema)

    def query_encoder(self) -> Block:
        return self.query_encoder_block

    def candidate_encoder(self) -> Block:
        return self.candidate_encoder_block

    def call(self, inputs: TabularData, targets=None, training: bool = False, testing: bool = False, **kwargs):
        self._check_inputs(inputs)
        if not self.built:
            self.build(next(iter(inputs.values())).device)
        out = self.prediction
        enc = {out.query_name: self.query_encoder_block(inputs), out.candidate_name: self.candidate_encoder_block(inputs)}
        return out(enc, candidate_ids=inputs[self.candidate_id_name], training=training, testing=testing, **kwargs)


def TwoTowerModelV2(query_tower, candidate_tower, candidate_id_tag=Tags.ITEM_ID, outputs=None, logits_temperature: float = 1.0,
                    negative_samplers=None, schema: Optional[Schema] = None, **kwargs) -> RetrievalModelV2:
    """models/retrieval.py:409-486: two Encoder towers + ContrastiveOutput(DotProduct, in-batch negatives by default)."""
    from .retrieval import ContrastiveOutput, Encoder

    assert isinstance(query_tower, Encoder), ValueError("The query tower should be an instance of `Encoder` class")
    assert isinstance(candidate_tower, Encoder), ValueError("The query tower should be an instance of `Encoder` class")
    if not outputs:
        if not negative_samplers:
            negative_samplers = ["in-batch"]
        outputs = ContrastiveOutput(to_call=None, negative_samplers=negative_samplers, logits_temperature=logits_temperature,
                                    **kwargs)
    if isinstance(outputs, (list, tuple)):
        if len(outputs) != 1:
            raise NotImplementedError("multi-task outputs are outside the hot path")
        outputs = outputs[0]
    return RetrievalModelV2(query_tower, candidate_tower, outputs, schema=schema, candidate_id_tag=candidate_id_tag)


def TwoTowerModel(schema: Schema, query_tower: MLP, item_tower: Optional[MLP] = None, query_tower_tag=Tags.USER,
                  item_tower_tag=Tags.ITEM,
                  embedding_options: EmbeddingOptions = EmbeddingOptions(embedding_dims=None, embedding_dim_default=64,
                                                                         infer_embedding_sizes=False,
                                                                         infer_embedding_sizes_multiplier=2.0),
                  post: Optional[Block] = None, prediction_tasks=None, logits_temperature: float = 1.0,
                  samplers: Sequence = (), **kwargs) -> RetrievalModel:
    """models/retrieval.py:106-203."""
    if not prediction_tasks:
        prediction_tasks = ItemRetrievalTask(schema, logits_temperature=logits_temperature, samplers=list(samplers))
    if isinstance(prediction_tasks, (list, tuple)):
        prediction_tasks = prediction_tasks[0]
    two_tower = TwoTowerBlock(schema=schema, query_tower=query_tower, item_tower=item_tower,
                              query_tower_tag=query_tower_tag, item_tower_tag=item_tower_tag,
                              embedding_options=embedding_options, post=post)
    return RetrievalModel(two_tower, prediction_tasks, schema)
