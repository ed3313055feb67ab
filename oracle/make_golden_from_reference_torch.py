"""Generate golden vectors by EXECUTING the reference's own PyTorch backend
(/root/reference/merlin/models/torch) in the build container — TEST INFRASTRUCTURE.

    python oracle/make_golden_from_reference_torch.py      # writes tests/golden/ref_torch_*.npz

The reference cannot be imported as a package here: merlin-core (merlin.schema, merlin.dispatch,
merlin.dtypes, ...), merlin-dataloader, torchmetrics and pytorch_lightning are not installed and
there is no network.  This script therefore registers *stand-in modules* for those third-party
packages (a functools.singledispatch-backed LazyDispatcher, this repo's Schema shim as
merlin.schema, inert placeholders for the rest), maps the `merlin.models.torch` package path to
the reference tree WITHOUT running its __init__ (which imports the data loader), and then imports
and runs the reference's module files unmodified, from where they lie:

  torch/transforms/agg.py   Concat, Stack            (sorted-name aggregation)
  torch/blocks/dlrm.py      DLRMInteraction, InteractionBlock
  torch/blocks/cross.py     CrossBlock, LazyMirrorLinear
  torch/blocks/mlp.py       MLPBlock
  torch/inputs/embedding.py EmbeddingTable.forward_bag path via F.embedding_bag semantics
  torch/blocks/dlrm.py      DLRMBlock END TO END (schema -> per-feature nn.Embedding tables -> bottom MLP over the
                            concatenated continuous features -> Stack -> DLRMInteraction -> [bottom | interactions]
                            -> top MLP), built from this repo's Schema shim standing in for merlin.schema
  torch/models/ranking.py   DLRMModel and DCNModel END TO END including BinaryOutput (Linear(1) + sigmoid): the
                            model-level outputs {target: (B, 1)} (pytorch_lightning.LightningModule is replaced by
                            an empty torch.nn.Module subclass)
  torch/models/ranking.py + torch/outputs/classification.py:44   ONE TRAINING STEP's gradients of that DLRMModel: the
                            backend's default BinaryOutput loss (nn.BCELoss on the sigmoid outputs) and torch.autograd
                            through the reference's modules -> loss, d/d(every Linear kernel and bias), d/d(tables)
  torch/outputs/classification.py  EmbeddingTablePrediction (weight-tied catalog logits x @ E^T + bias) with the
                            backend's default loss nn.CrossEntropyLoss evaluated on them
  utils/schema_utils.py     infer_embedding_dim / get_embedding_size_from_cardinality (backend-independent host code,
                            the default dims of InputBlockV2 / DCNModel: sum = 1024 on the bundled Criteo schema)
  torch/outputs/contrastive.py  ContrastiveOutput.contrastive_outputs ([positive | negatives] logits, one-hot
                            targets), rescore_false_negatives (accidental hits -> MIN_FLOAT)
  torch/outputs/sampling/in_batch.py   InBatchNegativeSampler
  torch/outputs/sampling/popularity.py LogUniformSampler.get_log_uniform_distr / get_unique_sampling_distr

Nothing is copied from the reference; only its outputs on seeded inputs are stored, together with
the inputs and weights, so tests/golden/replay.py can re-evaluate the oracle on the GPU box where
/root/reference does not exist.
"""
from __future__ import annotations

import importlib
import sys
import types
from functools import singledispatch
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden"


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, n):
        return _Inert()


class LazyDispatcher:
    """Stand-in for merlin.dispatch.lazy.LazyDispatcher (merlin-core): singledispatch + no-op
    lazy registration."""

    def __init__(self, func_or_name):
        if callable(func_or_name):
            self.__name__ = func_or_name.__name__
            self.dispatcher = singledispatch(func_or_name)
        else:
            self.__name__ = str(func_or_name)

            def _default(*a, **k):
                raise NotImplementedError(self.__name__)

            self.dispatcher = singledispatch(_default)

    def register(self, cls, func=None):
        return self.dispatcher.register(cls, func=func)

    def register_lazy(self, toplevel):
        return lambda f: f

    def dispatch(self, cls):
        return self.dispatcher.dispatch(cls)

    def __call__(self, arg, *a, **k):
        return self.dispatcher.dispatch(type(arg))(arg, *a, **k)


def _ns(name, path=None, **attrs):
    m = types.ModuleType(name)
    if path:
        m.__path__ = [str(path)]
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stand_ins():
    sys.path.insert(0, str(ROOT))
    import models_b200.schema as S

    _ns("merlin", REF / "merlin")
    _ns("merlin.schema", None, Schema=S.Schema, Tags=S.Tags, ColumnSchema=S.ColumnSchema, TagSet=set, TagsType=object)
    _ns("merlin.schema.tags", None, TagsType=object, Tags=S.Tags)
    _ns("merlin.schema.io", None)
    _ns("merlin.schema.io.tensorflow_metadata", None, TensorflowMetadata=_Inert)
    _ns("merlin.dtypes", None, float32=_Inert(), float64=_Inert(), int32=_Inert(), int64=_Inert())
    _ns("merlin.io", None, Dataset=_Inert)
    _ns("merlin.table", None, TensorTable=_Inert)
    _ns("merlin.dataloader", None)
    _ns("merlin.dataloader.torch", None, Loader=_Inert)
    _ns("merlin.core", None)
    _ns("merlin.core.dispatch", None, DataFrameType=object)
    _ns("merlin.dispatch", None)
    _ns("merlin.dispatch.lazy", None, LazyDispatcher=LazyDispatcher)
    tm = _ns("torchmetrics", None, Metric=_Inert, AUROC=_Inert, Accuracy=_Inert, Precision=_Inert, Recall=_Inert,
             MeanSquaredError=_Inert, MetricCollection=_Inert)
    tm.__getattr__ = lambda name: _Inert  # any other metric class (RetrievalHitRate, ...) is an inert placeholder
    import torch

    class LightningModule(torch.nn.Module):  # stand-in base of models/base.py Model(LightningModule, Block)
        pass

    _ns("pytorch_lightning", None, LightningModule=LightningModule, Trainer=_Inert, LightningDataModule=object)
    _ns("merlin.models", REF / "merlin" / "models")
    _ns("merlin.models.torch", REF / "merlin" / "models" / "torch")
    for sub in ("blocks", "inputs", "utils", "transforms", "outputs", "models"):
        _ns(f"merlin.models.torch.{sub}", REF / "merlin" / "models" / "torch" / sub)
    _ns("merlin.models.utils", REF / "merlin" / "models" / "utils")


def main():
    if not REF.exists():
        raise SystemExit("/root/reference is not present: golden vectors can only be regenerated in the build container")
    install_stand_ins()
    import torch

    agg = importlib.import_module("merlin.models.torch.transforms.agg")
    dlrm = importlib.import_module("merlin.models.torch.blocks.dlrm")
    cross = importlib.import_module("merlin.models.torch.blocks.cross")
    mlpm = importlib.import_module("merlin.models.torch.blocks.mlp")
    OUT.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(20260924)
    torch.manual_seed(0)

    # ---- 1. Stack + DLRMInteraction + InteractionBlock (shortcut concat) -----------------------
    B, D = 33, 16
    names = ["C1", "C10", "C2", "C21", "C3", "continuous"]  # 'continuous' = bottom-MLP output key
    feats = {n: rng.standard_normal((B, D)).astype(np.float32) for n in names}
    tin = {k: torch.from_numpy(v) for k, v in feats.items()}
    stacked = agg.Stack(dim=1)(tin)
    inter = dlrm.DLRMInteraction()(stacked)
    block = dlrm.InteractionBlock(dlrm.DLRMInteraction())
    block_out = block(tin)
    np.savez(OUT / "ref_torch_dlrm_interaction.npz", kind="dlrm_interaction", names=np.array(names),
             **{f"in_{k}": v for k, v in feats.items()}, stacked=stacked.numpy(), interactions=inter.numpy(),
             block_out=block_out.numpy())

    # ---- 2. Concat (sorted names, (B,) -> (B,1), float cast) -----------------------------------
    cfe = {"I1": rng.random(B).astype(np.float32), "I10": rng.random(B).astype(np.float32),
           "I2": rng.random((B, 1)).astype(np.float32), "emb_b": rng.standard_normal((B, 5)).astype(np.float32),
           "Z": rng.integers(0, 5, (B, 2)).astype(np.int64)}
    cat = agg.Concat()({k: torch.from_numpy(v) for k, v in cfe.items()})
    np.savez(OUT / "ref_torch_concat.npz", kind="concat", names=np.array(sorted(cfe)),
             **{f"in_{k}": v for k, v in cfe.items()}, out=cat.numpy())

    # ---- 3. CrossBlock (DCN-v2) ----------------------------------------------------------------
    d, depth = 24, 3
    x = rng.standard_normal((B, d)).astype(np.float32)
    cb = cross.CrossBlock.with_depth(depth)
    y = cb(torch.from_numpy(x))
    ws = []
    for m in cb.values:
        lin = m if isinstance(m, torch.nn.Linear) else [c for c in m.modules() if isinstance(c, torch.nn.Linear)][0]
        ws.append((lin.weight.detach().numpy().T.copy(), lin.bias.detach().numpy().copy()))
    np.savez(OUT / "ref_torch_cross.npz", kind="cross", x=x, out=y.detach().numpy(),
             **{f"kernel_{i}": w for i, (w, _) in enumerate(ws)}, **{f"bias_{i}": b for i, (_, b) in enumerate(ws)})

    # ---- 4. MLPBlock ---------------------------------------------------------------------------
    xin = rng.standard_normal((B, 13)).astype(np.float32)
    mlp = mlpm.MLPBlock([32, 16])
    ym = mlp(torch.from_numpy(xin))
    lins = [m for m in mlp.modules() if isinstance(m, torch.nn.Linear)]
    np.savez(OUT / "ref_torch_mlp.npz", kind="mlp", x=xin, out=ym.detach().numpy(),
             **{f"kernel_{i}": l.weight.detach().numpy().T.copy() for i, l in enumerate(lins)},
             **{f"bias_{i}": l.bias.detach().numpy().copy() for i, l in enumerate(lins)})

    # ---- 5. embedding bag (the op torch/inputs/embedding.py:264-293 dispatches to) --------------
    emb = importlib.import_module("merlin.models.torch.inputs.embedding")
    table = rng.standard_normal((19, 8)).astype(np.float32)
    lens = rng.integers(1, 5, B)
    offsets = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    values = rng.integers(0, 19, offsets[-1]).astype(np.int64)
    outs = {}
    for mode in ("mean", "sum"):
        outs[mode] = torch.nn.functional.embedding_bag(torch.from_numpy(values), torch.from_numpy(table),
                                                       torch.from_numpy(offsets[:-1]), mode=mode).numpy()
    src = Path(emb.__file__).read_text()
    assert "embedding_bag" in src, "reference torch EmbeddingTable no longer uses F.embedding_bag"
    np.savez(OUT / "ref_torch_embedding_bag.npz", kind="embedding_bag", table=table, values=values, offsets=offsets,
             out_mean=outs["mean"], out_sum=outs["sum"])
    # ---- 6. contrastive logits: [positive | negatives] layout, false-negative rescoring, targets ----
    import types

    con = importlib.import_module("merlin.models.torch.outputs.contrastive")
    Bc, Dc, Nn = 29, 12, 17
    q = rng.standard_normal((Bc, Dc)).astype(np.float32)
    pos = rng.standard_normal((Bc, Dc)).astype(np.float32)
    ids = rng.integers(0, 9, Bc).astype(np.int64)            # duplicates -> accidental hits beyond the diagonal
    sampler = importlib.import_module("merlin.models.torch.outputs.sampling.in_batch").InBatchNegativeSampler()
    neg_t, neg_id_t = sampler(torch.from_numpy(pos), torch.from_numpy(ids))   # in-batch: the batch's own items
    # utils/constants.py:19 `np.finfo(np.float16).min / 100.0` is -655.04 under the NumPy 1.x the reference pins
    # (value-based casting -> float64); under this container's NumPy 2 (NEP 50) the same expression stays float16 and
    # rounds to -655.0.  The functions take the score as an argument: pass the reference's intended value.
    MINF = float(np.finfo(np.float16).min) / 100.0
    fake = types.SimpleNamespace(downscore_false_negatives=True, false_negative_score=MINF)
    out_ds = con.ContrastiveOutput.contrastive_outputs(fake, torch.from_numpy(q), torch.from_numpy(pos), neg_t,
                                                       positive_id=torch.from_numpy(ids), negative_id=neg_id_t)
    target = fake.target.numpy().copy()
    fake2 = types.SimpleNamespace(downscore_false_negatives=False, false_negative_score=MINF)
    out_plain = con.ContrastiveOutput.contrastive_outputs(fake2, torch.from_numpy(q), torch.from_numpy(pos), neg_t)
    # a separate negative set (sampled softmax shape): N != B
    neg2 = rng.standard_normal((Nn, Dc)).astype(np.float32)
    neg2_ids = rng.integers(0, 9, Nn).astype(np.int64)
    scores2 = (q @ neg2.T).astype(np.float32)
    resc, valid = con.rescore_false_negatives(torch.from_numpy(ids), torch.from_numpy(neg2_ids), torch.from_numpy(scores2),
                                              MINF)
    bias_mod = importlib.import_module("merlin.models.torch.transforms.bias")
    TEMP = 0.37
    out_scaled = bias_mod.LogitsTemperatureScaler(TEMP)(out_ds)   # applied after the rescoring: MIN_FLOAT / T on hits
    np.savez(OUT / "ref_torch_contrastive.npz", kind="contrastive", query=q, positive=pos, ids=ids,
             temperature=np.float32(TEMP), out_scaled=out_scaled.numpy(),
             negative=neg_t.numpy(), negative_ids=neg_id_t.numpy(), out_downscored=out_ds.numpy(), out_plain=out_plain.numpy(),
             target=target, min_float=np.float32(MINF), min_float_as_imported=np.float32(con.MIN_FLOAT), neg2=neg2, neg2_ids=neg2_ids, scores2=scores2,
             rescored2=resc.numpy(), valid2=valid.numpy())

    # ---- 7. log-uniform (popularity) sampling probabilities --------------------------------------
    pop = importlib.import_module("merlin.models.torch.outputs.sampling.popularity")
    fs = types.SimpleNamespace()
    cases = [(20, 0, 7), (50, 3, 11), (1000, 1, 100)]
    blobs = {}
    for i, (max_id, min_id, n_sample) in enumerate(cases):
        d = pop.LogUniformSampler.get_log_uniform_distr(fs, max_id, min_id)
        u = pop.LogUniformSampler.get_unique_sampling_distr(fs, d.clone(), n_sample)
        blobs[f"probs_{i}"] = d.numpy()
        blobs[f"unique_{i}"] = u.numpy()
    np.savez(OUT / "ref_torch_log_uniform.npz", kind="log_uniform", cases=np.array(cases, dtype=np.int64), **blobs)
    # ---- 8. the whole DLRMBlock of the reference's torch backend ----------------------------------
    import models_b200.schema as S

    cats = [("C1", 30), ("C10", 7), ("C2", 101), ("C21", 12), ("C3", 55)]      # Criteo-style names: 'C10' < 'C2'
    conts = ["I1", "I10", "I2", "I3"]
    cols = [S.ColumnSchema(n, tags=("categorical",), dtype="int64", properties={"domain": {"min": 0, "max": mx, "name": n}})
            for n, mx in cats]
    cols += [S.ColumnSchema(n, tags=("continuous",), dtype="float32") for n in conts]
    torch.manual_seed(7)
    dim, Bm = 16, 37
    blk = dlrm.DLRMBlock(S.Schema(cols), dim, bottom_block=mlpm.MLPBlock([32, dim]), top_block=mlpm.MLPBlock([24, 8]))
    batch = {n: rng.integers(0, mx + 1, Bm).astype(np.int64) for n, mx in cats}
    batch.update({n: rng.random(Bm).astype(np.float32) for n in conts})
    out = blk({k: torch.from_numpy(v) for k, v in batch.items()})
    blobs = {}
    lin_bottom, lin_top = [], []
    for name, m in blk.named_modules():
        if isinstance(m, torch.nn.Embedding):
            feat = [n for n, _ in cats if f".{n}." in f".{name}."][0]
            blobs[f"table_{feat}"] = m.weight.detach().numpy().copy()
        elif isinstance(m, torch.nn.Linear):
            (lin_bottom if ".continuous." in f".{name}." else lin_top).append(m)
    assert len(lin_bottom) == 2 and len(lin_top) == 2 and len([k for k in blobs if k.startswith("table_")]) == len(cats)
    for tag, lins in (("bottom", lin_bottom), ("top", lin_top)):
        for i, l in enumerate(lins):
            blobs[f"{tag}_kernel_{i}"] = l.weight.detach().numpy().T.copy()   # Keras layout (in, out)
            blobs[f"{tag}_bias_{i}"] = l.bias.detach().numpy().copy()
            blobs[f"{tag}_act_{i}"] = np.array("relu")
    np.savez(OUT / "ref_torch_dlrm_block.npz", kind="dlrm_block", cat_names=np.array([n for n, _ in cats]),
             cat_max=np.array([mx for _, mx in cats], dtype=np.int64), cont_names=np.array(conts), dim=np.int64(dim),
             out=out.detach().numpy(), **{f"batch_{k}": v for k, v in batch.items()}, **blobs)
    # ---- 9. DLRMModel / DCNModel of the torch backend, BinaryOutput included ----------------------
    ranking = importlib.import_module("merlin.models.torch.models.ranking")
    target = S.ColumnSchema("click", tags=("target", "binary_classification"), dtype="int64")
    mschema = S.Schema(cols + [target])
    feed = {k: torch.from_numpy(v) for k, v in batch.items()}

    def linears(module):
        return [(n, m) for n, m in module.named_modules() if isinstance(m, torch.nn.Linear)]

    def tables_of(module):
        out = {}
        for name, m in module.named_modules():
            if isinstance(m, torch.nn.Embedding):
                feat = [n for n, _ in cats if f".{n}." in f".{name}."][0]
                out[f"table_{feat}"] = m.weight.detach().numpy().copy()
        return out

    def pack(tag, lins, act):
        d = {}
        for i, l in enumerate(lins):
            d[f"{tag}_kernel_{i}"] = l.weight.detach().numpy().T.copy()
            d[f"{tag}_bias_{i}"] = l.bias.detach().numpy().copy()
            d[f"{tag}_act_{i}"] = np.array(act)
        return d

    torch.manual_seed(11)
    dm = ranking.DLRMModel(mschema, dim=dim, bottom_block=mlpm.MLPBlock([32, dim]), top_block=mlpm.MLPBlock([24, 8]))
    dout = dm(feed)["click"]
    L = linears(dm)
    assert len(L) == 5
    np.savez(OUT / "ref_torch_dlrm_model.npz", kind="dlrm_model", cat_names=np.array([n for n, _ in cats]),
             cat_max=np.array([mx for _, mx in cats], dtype=np.int64), cont_names=np.array(conts), dim=np.int64(dim),
             out=dout.detach().numpy(), **{f"batch_{k}": v for k, v in batch.items()}, **tables_of(dm),
             **pack("bottom", [m for n, m in L if ".continuous." in f".{n}."], "relu"),
             **pack("top", [m for n, m in L if ".continuous." not in f".{n}."][:2], "relu"),
             **pack("head", [L[-1][1]], "sigmoid"))

    # ---- 9a. one TRAINING step's gradients of the same DLRMModel: the backend's default BinaryOutput loss (nn.BCELoss on
    # the sigmoid outputs, torch/outputs/classification.py:44) and torch.autograd through the reference's own modules.
    # Its own rng: the draws of the sections below must not move.
    y_train = np.random.default_rng(914).integers(0, 2, Bm).astype(np.float32)
    loss_mods = [m for m in dm.modules() if isinstance(m, torch.nn.BCELoss)]
    assert len(loss_mods) == 1, loss_mods
    dm.zero_grad()
    train_loss = loss_mods[0](dout, torch.from_numpy(y_train).reshape(-1, 1))
    train_loss.backward()

    def pack_grads(tag, lins):
        d = {}
        for i, l in enumerate(lins):
            d[f"grad_{tag}_kernel_{i}"] = l.weight.grad.detach().numpy().T.copy()
            d[f"grad_{tag}_bias_{i}"] = l.bias.grad.detach().numpy().copy()
        return d

    tgrads = {}
    for name, m in dm.named_modules():
        if isinstance(m, torch.nn.Embedding):
            feat = [n for n, _ in cats if f".{n}." in f".{name}."][0]
            tgrads[f"grad_table_{feat}"] = m.weight.grad.detach().numpy().copy()
    np.savez(OUT / "ref_torch_dlrm_train.npz", kind="dlrm_train", cat_names=np.array([n for n, _ in cats]),
             cat_max=np.array([mx for _, mx in cats], dtype=np.int64), cont_names=np.array(conts), dim=np.int64(dim),
             out=dout.detach().numpy(), targets=y_train, loss=np.float32(train_loss.item()),
             **{f"batch_{k}": v for k, v in batch.items()}, **tables_of(dm),
             **pack("bottom", [m for n, m in L if ".continuous." in f".{n}."], "relu"),
             **pack("top", [m for n, m in L if ".continuous." not in f".{n}."][:2], "relu"),
             **pack("head", [L[-1][1]], "sigmoid"),
             **pack_grads("bottom", [m for n, m in L if ".continuous." in f".{n}."]),
             **pack_grads("top", [m for n, m in L if ".continuous." not in f".{n}."][:2]),
             **pack_grads("head", [L[-1][1]]), **tgrads)

    torch.manual_seed(12)
    cm = ranking.DCNModel(mschema, depth=3, deep_block=mlpm.MLPBlock([32, 16]))
    cout = cm(feed)["click"]
    L = linears(cm)
    tabs = tables_of(cm)
    d_in = sum(t.shape[1] for t in tabs.values()) + len(conts)
    cross_l = [m for n, m in L if m.in_features == d_in and m.out_features == d_in]
    rest = [m for n, m in L if not (m.in_features == d_in and m.out_features == d_in)]
    assert len(cross_l) == 3 and len(rest) == 3
    np.savez(OUT / "ref_torch_dcn_model.npz", kind="dcn_model", cat_names=np.array([n for n, _ in cats]),
             cat_max=np.array([mx for _, mx in cats], dtype=np.int64), cont_names=np.array(conts),
             emb_dims=np.array([tabs[f"table_{n}"].shape[1] for n, _ in cats], dtype=np.int64), out=cout.detach().numpy(),
             **{f"batch_{k}": v for k, v in batch.items()}, **tabs, **pack("cross", cross_l, "linear"),
             **pack("deep", rest[:2], "relu"), **pack("head", [rest[2]], "sigmoid"))
    # ---- 9b. DCNModel over a schema with a ragged multi-hot feature (name__values / name__offsets) ----
    cats_mh = [("C1", 30), ("C10", 7), ("genres", 18)]
    cols_mh = [S.ColumnSchema("C1", tags=("categorical",), dtype="int64", properties={"domain": {"min": 0, "max": 30, "name": "C1"}}),
               S.ColumnSchema("C10", tags=("categorical",), dtype="int64", properties={"domain": {"min": 0, "max": 7, "name": "C10"}}),
               S.ColumnSchema("genres", tags=("categorical",), dtype="int64", is_list=True, is_ragged=True,
                              properties={"domain": {"min": 0, "max": 18, "name": "genres"}, "value_count": {"min": 1, "max": 4}}),
               S.ColumnSchema("I1", tags=("continuous",), dtype="float32"), S.ColumnSchema("I2", tags=("continuous",), dtype="float32"),
               target]
    lens = rng.integers(1, 5, Bm)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    batch_mh = {"C1": rng.integers(0, 31, Bm).astype(np.int64), "C10": rng.integers(0, 8, Bm).astype(np.int64),
                "genres__values": rng.integers(0, 19, int(offs[-1])).astype(np.int64), "genres__offsets": offs,
                "I1": rng.random(Bm).astype(np.float32), "I2": rng.random(Bm).astype(np.float32)}
    torch.manual_seed(14)
    mh = ranking.DCNModel(S.Schema(cols_mh), depth=2, deep_block=mlpm.MLPBlock([16, 8]))
    mh_out = mh({k: torch.from_numpy(v) for k, v in batch_mh.items()})["click"]
    combiners = {getattr(m, "seq_combiner") for _, m in mh.named_modules() if isinstance(getattr(m, "seq_combiner", None), str)}
    assert combiners == {"mean"}, combiners   # the default bag combiner of the backend
    tabs = {}
    for name, m in mh.named_modules():
        if isinstance(m, torch.nn.Embedding):
            feat = [n for n, _ in cats_mh if f".{n}." in f".{name}."][0]
            tabs[f"table_{feat}"] = m.weight.detach().numpy().copy()
    L = linears(mh)
    d_in = sum(t.shape[1] for t in tabs.values()) + 2
    cross_l = [m for n, m in L if m.in_features == d_in and m.out_features == d_in]
    rest = [m for n, m in L if not (m.in_features == d_in and m.out_features == d_in)]
    assert len(cross_l) == 2 and len(rest) == 3
    np.savez(OUT / "ref_torch_dcn_multihot.npz", kind="dcn_model", cat_names=np.array([n for n, _ in cats_mh]),
             cat_max=np.array([mx for _, mx in cats_mh], dtype=np.int64), cont_names=np.array(["I1", "I2"]),
             list_names=np.array(["genres"]), emb_dims=np.array([tabs[f"table_{n}"].shape[1] for n, _ in cats_mh], dtype=np.int64),
             out=mh_out.detach().numpy(), **{f"batch_{k}": v for k, v in batch_mh.items()}, **tabs,
             **pack("cross", cross_l, "linear"), **pack("deep", rest[:2], "relu"), **pack("head", [rest[2]], "sigmoid"))

    # ---- 10. weight-tied catalog logits (CategoricalOutput / EmbeddingTablePrediction) -------------
    clsm = importlib.import_module("merlin.models.torch.outputs.classification")
    item = S.ColumnSchema("item_id", tags=("categorical", "item_id"), dtype="int64",
                          properties={"domain": {"min": 0, "max": 299, "name": "item_id"}})
    torch.manual_seed(13)
    etab = emb.EmbeddingTable(16, S.Schema([item]))
    pred = clsm.EmbeddingTablePrediction(etab)
    with torch.no_grad():
        pred.bias.copy_(torch.from_numpy((rng.standard_normal(pred.num_classes) * 0.1).astype(np.float32)))
    xq = rng.standard_normal((41, 16)).astype(np.float32)
    tgt = rng.integers(0, 300, 41).astype(np.int64)
    logits = pred(torch.from_numpy(xq))
    ce = torch.nn.CrossEntropyLoss(reduction="none")(logits, torch.from_numpy(tgt))   # the backend's default loss
    tk = torch.topk(logits, 10, dim=1)
    np.savez(OUT / "ref_torch_catalog.npz", kind="catalog", table=pred.embeddings().detach().numpy().copy(),
             bias=pred.bias.detach().numpy().copy(), x=xq, targets=tgt, logits=logits.detach().numpy(),
             cross_entropy=ce.detach().numpy(), topk_scores=tk.values.detach().numpy(), topk_ids=tk.indices.numpy())
    # ---- 11. MLPBlock with every activation the hot path offers (torch modules = Keras definitions:
    #          GELU erf-based, SELU/ELU with the standard constants) -------------------------------
    xa = (rng.standard_normal((Bm, 11)) * 2.5).astype(np.float32)
    act_blobs = {}
    for aname, amod in (("sigmoid", torch.nn.Sigmoid), ("tanh", torch.nn.Tanh), ("selu", torch.nn.SELU), ("elu", torch.nn.ELU),
                        ("gelu", torch.nn.GELU), ("relu", torch.nn.ReLU)):
        torch.manual_seed(21)
        blk_a = mlpm.MLPBlock([24, 9], activation=amod)
        ya = blk_a(torch.from_numpy(xa))
        lins_a = [m for m in blk_a.modules() if isinstance(m, torch.nn.Linear)]
        act_blobs[f"out_{aname}"] = ya.detach().numpy()
        for i, l in enumerate(lins_a):  # same seed -> same weights for every activation; store once per name anyway
            act_blobs[f"{aname}_kernel_{i}"] = l.weight.detach().numpy().T.copy()
            act_blobs[f"{aname}_bias_{i}"] = l.bias.detach().numpy().copy()
    np.savez(OUT / "ref_torch_mlp_activations.npz", kind="mlp_acts", x=xa,
             names=np.array(["sigmoid", "tanh", "selu", "elu", "gelu", "relu"]), **act_blobs)

    # ---- 12. inferred embedding dims (merlin/models/utils/schema_utils.py:169-207) ------------------
    su = importlib.import_module("merlin.models.utils.schema_utils")
    from models_b200 import datasets as _ds

    names = sorted(_ds.CRITEO_MAX)
    maxes = np.array([_ds.CRITEO_MAX[n] for n in names] + [0, 1, 2, 15, 16, 255, 256, 4095, 10**6, 4 * 10**7], dtype=np.int64)
    colsd = [S.ColumnSchema(f"f{i}", tags=("categorical",), dtype="int64", properties={"domain": {"min": 0, "max": int(m), "name": f"f{i}"}})
             for i, m in enumerate(maxes)]
    dims_default = np.array([su.infer_embedding_dim(c) for c in colsd], dtype=np.int64)                      # x2, multiple of 8
    dims_plain = np.array([su.infer_embedding_dim(c, multiplier=3.0, ensure_multiple_of_8=False) for c in colsd], dtype=np.int64)
    np.savez(OUT / "ref_torch_embedding_dims.npz", kind="embedding_dims", max_id=maxes, dims_default=dims_default,
             dims_mult3_plain=dims_plain, n_criteo=np.int64(len(names)))
    # ---- 13. two towers: TabularInputBlock (continuous + embedded categoricals, sorted-name concat) -> MLPBlock,
    #          on the ML-1M column names of SURVEY §8(a11): query = userId + TE_* floats, item = movieId + genres
    #          (ragged multi-hot, mean combiner) + TE_movieId_rating; then the retrieval scorer pieces of the same
    #          backend on the tower outputs (row-wise dot at inference, contrastive [positive | in-batch negatives]).
    tab = importlib.import_module("merlin.models.torch.inputs.tabular")
    embm = importlib.import_module("merlin.models.torch.inputs.embedding")
    Bt, dt = 41, 16

    def catc(name, mx, tags=(), is_list=False):
        props = {"domain": {"min": 0, "max": mx, "name": name}}
        if is_list:
            props["value_count"] = {"min": 1, "max": 4}
        return S.ColumnSchema(name, tags=("categorical",) + tuple(tags), dtype="int64", is_list=is_list, is_ragged=is_list,
                              properties=props)

    q_cols = [catc("userId", 6040, ("user", "user_id"))] + [
        S.ColumnSchema(n, tags=("continuous", "user"), dtype="float32")
        for n in ("TE_age_rating", "TE_gender_rating", "TE_occupation_rating", "TE_userId_rating", "TE_zipcode_rating")]
    i_cols = [catc("movieId", 3684, ("item", "item_id")), catc("genres", 18, ("item",), is_list=True),
              S.ColumnSchema("TE_movieId_rating", tags=("continuous", "item"), dtype="float32")]

    def tower_init(block):
        block.add_route(S.Tags.CONTINUOUS, required=False)
        block.add_route(S.Tags.CATEGORICAL, embm.EmbeddingTables(dt, seq_combiner="mean"))

    lens = rng.integers(1, 5, Bt)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tbatch = {"userId": rng.integers(1, 6041, Bt).astype(np.int64), "movieId": rng.integers(1, 3685, Bt).astype(np.int64),
              "genres__values": rng.integers(1, 19, int(offs[-1])).astype(np.int64), "genres__offsets": offs,
              "TE_movieId_rating": rng.random(Bt).astype(np.float32)}
    tbatch["movieId"][5] = tbatch["movieId"][3]  # duplicate item -> an accidental hit off the diagonal
    for c in q_cols[1:]:
        tbatch[c.name] = rng.random(Bt).astype(np.float32)
    blobs = {}
    outs = {}
    for tag, cols_t, seed in (("query", q_cols, 31), ("item", i_cols, 32)):
        torch.manual_seed(seed)
        sch = S.Schema(cols_t)
        inp = tab.TabularInputBlock(sch, init=tower_init, agg="concat")
        mlp_t = mlpm.MLPBlock([24, 12])
        names_t = [c.name for c in cols_t]
        feed_t = {k: torch.from_numpy(v) for k, v in tbatch.items() if any(k == n or k.startswith(n + "__") for n in names_t)}
        concat = inp(feed_t)
        out_t = mlp_t(concat)
        outs[tag] = out_t.detach()
        blobs[f"{tag}_concat"] = concat.detach().numpy()
        blobs[f"{tag}_out"] = out_t.detach().numpy()
        for name, m in inp.named_modules():
            if isinstance(m, torch.nn.Embedding):
                feat = [c.name for c in cols_t if f".{c.name}." in f".{name}."][0]
                blobs[f"{tag}_table_{feat}"] = m.weight.detach().numpy().copy()
        for i, l in enumerate([m for m in mlp_t.modules() if isinstance(m, torch.nn.Linear)]):
            blobs[f"{tag}_kernel_{i}"] = l.weight.detach().numpy().T.copy()
            blobs[f"{tag}_bias_{i}"] = l.bias.detach().numpy().copy()
    qo, io = outs["query"], outs["item"]
    inference_scores = (qo * io).sum(-1, keepdim=True)  # DotProduct of the towers at inference (outputs/contrastive.py)
    item_ids = torch.from_numpy(tbatch["movieId"])
    neg_e, neg_i = sampler(io, item_ids)
    fake3 = types.SimpleNamespace(downscore_false_negatives=True, false_negative_score=MINF)
    logits_t = con.ContrastiveOutput.contrastive_outputs(fake3, qo, io, neg_e, positive_id=item_ids, negative_id=neg_i)
    np.savez(OUT / "ref_torch_two_tower.npz", kind="two_tower", dim=np.int64(dt), query_cols=np.array([c.name for c in q_cols]),
             item_cols=np.array([c.name for c in i_cols]), inference_scores=inference_scores.numpy(),
             train_logits=logits_t.numpy(), train_targets=fake3.target.numpy(), min_float=np.float32(MINF),
             **{f"batch_{k}": v for k, v in tbatch.items()}, **blobs)
    print("wrote", sorted(p.name for p in OUT.glob("ref_torch_*.npz")))


if __name__ == "__main__":
    main()
