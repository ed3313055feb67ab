"""Write the oracle's own golden vectors (oracle_*.npz): small seeded model-level cases of the
BASELINE configs.  They pin the oracle against drift and give the GPU tests fixed, committed
inputs + weights + expected logits that do not depend on any RNG implementation.

    python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from models_b200 import datasets  # noqa: E402
from models_b200.schema import Tags  # noqa: E402
from oracle import oracle  # noqa: E402

OUT = Path(__file__).resolve().parent


def mk_layers(rng, dims, width, act="relu", last=None):
    out = []
    for i, d in enumerate(dims):
        lim = np.sqrt(6.0 / (width + d))
        a = last if (last is not None and i == len(dims) - 1) else act
        out.append({"kernel": rng.uniform(-lim, lim, (width, d)).astype(np.float32),
                    "bias": (rng.standard_normal(d) * 0.1).astype(np.float32), "activation": a})
        width = d
    return out


def pack_layers(name, layers):
    d = {}
    for i, l in enumerate(layers):
        d[f"{name}_kernel_{i}"] = l["kernel"]
        if l.get("bias") is not None:
            d[f"{name}_bias_{i}"] = l["bias"]
        d[f"{name}_act_{i}"] = np.array(l.get("activation") or "linear")
    return d


def main():
    rng = np.random.default_rng(777)
    # ---- DLRM, Criteo shape with capped cardinalities, D = 16 ------------------------------------
    schema = datasets.criteo_schema({k: min(v, 200) for k, v in datasets.CRITEO_MAX.items()})
    cat = schema.select_by_tag(Tags.CATEGORICAL)
    cont = schema.select_by_tag(Tags.CONTINUOUS).column_names
    D, B = 16, 48
    tables = {c.name: rng.uniform(-0.05, 0.05, (c.int_domain.max + 1, D)).astype(np.float32) for c in cat}
    f2t = {c.name: c.name for c in cat}
    batch, _ = datasets.split_targets(schema, datasets.generate_batch(schema, B, seed=11, index_law="uniform"))
    bottom, top = mk_layers(rng, [32, D], 13), mk_layers(rng, [32, 16, 8], D + 27 * 26 // 2)
    head = mk_layers(rng, [1], 8, last="sigmoid")
    exp = oracle.dlrm_forward(batch, tables, f2t, cont, bottom, top, head[0])
    np.savez(OUT / "oracle_dlrm_criteo_small.npz", kind="oracle_model",
             spec=json.dumps({"model": "dlrm", "feature_table": f2t, "continuous": cont}),
             expected=exp, **{f"batch_{k}": v for k, v in batch.items()}, **{f"table_{k}": v for k, v in tables.items()},
             **pack_layers("bottom", bottom), **pack_layers("top", top), **pack_layers("head", head))

    # ---- DCN-v2, inferred dims ---------------------------------------------------------------------
    dims = {c.name: oracle.infer_embedding_dim(c.int_domain.max + 1) for c in cat}
    tables2 = {c.name: rng.uniform(-0.05, 0.05, (c.int_domain.max + 1, dims[c.name])).astype(np.float32) for c in cat}
    d = sum(dims.values()) + 13
    cross = [{"kernel": (rng.standard_normal((d, d)) * 0.02).astype(np.float32),
              "bias": (rng.standard_normal(d) * 0.01).astype(np.float32), "activation": "linear"} for _ in range(3)]
    deep, head2 = mk_layers(rng, [32, 16], d), mk_layers(rng, [1], 16, last="sigmoid")
    exp2 = oracle.dcn_forward(batch, tables2, f2t, cont, cross, deep, head2[0])
    np.savez(OUT / "oracle_dcn_criteo_small.npz", kind="oracle_model",
             spec=json.dumps({"model": "dcn", "feature_table": f2t, "continuous": cont}), expected=exp2,
             **{f"batch_{k}": v for k, v in batch.items()}, **{f"table_{k}": v for k, v in tables2.items()},
             **pack_layers("cross", cross), **pack_layers("deep", deep), **pack_layers("head", head2))

    # ---- Two-tower, MovieLens-1M schema, B = 32 (config 1 shape) -----------------------------------
    ml = datasets.movielens_1m_schema()
    mb, _ = datasets.split_targets(ml, datasets.generate_batch(ml, 32, seed=12))
    mb["movieId"][5] = mb["movieId"][9]  # force one accidental hit off the diagonal
    q_cat, i_cat = ["userId"], ["movieId", "genres"]
    q_con = ["TE_age_rating", "TE_gender_rating", "TE_occupation_rating", "TE_zipcode_rating", "TE_userId_rating"]
    i_con = ["TE_movieId_rating"]
    tb = {n: rng.normal(0, 0.05, (ml[n].int_domain.max + 1, 16)).astype(np.float32) for n in q_cat + i_cat}
    ql, il = mk_layers(rng, [32, 16], 16 + 5), mk_layers(rng, [32, 16], 32 + 1)
    spec = {"model": "two_tower", "feature_table": {}, "query_feature_table": {n: n for n in q_cat},
            "item_feature_table": {n: n for n in i_cat}, "query_continuous": q_con, "item_continuous": i_con,
            "item_id": "movieId", "temperature": 0.5}
    q = oracle.tower_forward(mb, tb, spec["query_feature_table"], q_con, ql)
    it = oracle.tower_forward(mb, tb, spec["item_feature_table"], i_con, il)
    exp3, _ = oracle.contrastive_logits(q, it, it, mb["movieId"], mb["movieId"], True, oracle.MIN_FLOAT, temperature=0.5)
    np.savez(OUT / "oracle_two_tower_ml1m_small.npz", kind="oracle_model", spec=json.dumps(spec), expected=exp3,
             **{f"batch_{k}": v for k, v in mb.items()}, **{f"table_{k}": v for k, v in tb.items()},
             **pack_layers("query", ql), **pack_layers("item", il))
    print("wrote", sorted(p.name for p in OUT.glob("oracle_*.npz")))


if __name__ == "__main__":
    main()
