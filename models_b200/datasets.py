"""Workload shapes and synthetic batches for the BASELINE.json configs.

The shapes restate facts of the reference's bundled schemas
(merlin/datasets/advertising/criteo/transformed/schema.pbtxt,
merlin/datasets/entertainment/movielens/1m/schema.pbtxt): column names, tags and
`int_domain.max`.  `Schema.from_proto_text` reads the reference files themselves when they are
available; these built-ins exist because /root/reference is absent on the GPU box.

The generator follows merlin/datasets/synthetic.py: id columns are
clip(int(lognormal(3, 1)), 1, max) (:199-203, :218-222, :244-248), other integer columns
randint(min, max) (:275-277), floats uniform(0, 1) (:283-285); ragged list columns draw a
length per row (:356-361).  A "uniform" law (randint(0, card)) is offered for the gather
benchmark's worst case, and "zipf" for a skewed realistic one.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .schema import ColumnSchema, Schema, Tags

# int_domain.max of C1..C26 in the bundled transformed-Criteo schema (table rows = max + 1)
CRITEO_MAX = {
    "C1": 9999999, "C2": 29427, "C3": 15127, "C4": 7295, "C5": 19901, "C6": 3, "C7": 6465,
    "C8": 1310, "C9": 61, "C10": 9999999, "C11": 622921, "C12": 219556, "C13": 10, "C14": 2209,
    "C15": 9779, "C16": 71, "C17": 4, "C18": 963, "C19": 14, "C20": 9999999, "C21": 4384510,
    "C22": 9999999, "C23": 290588, "C24": 10829, "C25": 95, "C26": 34,
}

# "Criteo-TB shape" = MLPerf DLRM-DCNv2 capped cardinalities (not in
# the reference, an external shape used for the row-sharded 8-GPU configuration)
CRITEO_TB_ROWS = [
    40000000, 39060, 17295, 7424, 20265, 3, 7122, 1543, 63, 40000000, 3067956, 405282, 10, 2209,
    11938, 155, 4, 976, 14, 40000000, 40000000, 40000000, 590152, 12973, 108, 36,
]


def _cat(name, max_, tags=(), is_list=False, is_ragged=False, domain_name=None, value_count=None):
    props = {"domain": {"min": 0, "max": int(max_), "name": domain_name or name}}
    if value_count:
        props["value_count"] = value_count
    return ColumnSchema(name, tags=(Tags.CATEGORICAL,) + tuple(tags), dtype="int64", is_list=is_list,
                        is_ragged=is_ragged, properties=props)


def _cont(name, tags=()):
    return ColumnSchema(name, tags=(Tags.CONTINUOUS,) + tuple(tags), dtype="float32")


def criteo_schema(cardinality_max: Optional[Dict[str, int]] = None) -> Schema:
    """26 categorical (C1..C26) + 13 continuous (I1..I13) + binary `label`; file order of the
    reference schema (C21 first: it carries the item_id tag)."""
    mx = dict(CRITEO_MAX if cardinality_max is None else cardinality_max)
    order = ["C21"] + [f"C{i}" for i in range(1, 27) if i != 21]
    cols = [_cat(n, mx[n], tags=(Tags.ITEM_ID,) if n == "C21" else ()) for n in order]
    cols += [_cont(f"I{i}") for i in range(1, 14)]
    cols.append(ColumnSchema("label", tags=(Tags.BINARY_CLASSIFICATION, Tags.TARGET), dtype="int64"))
    return Schema(cols)


def criteo_tb_schema() -> Schema:
    return criteo_schema({f"C{i + 1}": r - 1 for i, r in enumerate(CRITEO_TB_ROWS)})


def movielens_1m_schema() -> Schema:
    cols = [
        _cat("userId", 6040, tags=(Tags.USER, Tags.USER_ID)),
        _cat("movieId", 3684, tags=(Tags.ITEM, Tags.ITEM_ID)),
        _cat("title", 3684),
        _cat("genres", 18, tags=(Tags.ITEM,), is_list=True, is_ragged=True, value_count={"min": 1, "max": None}),
        _cat("gender", 2),
        _cat("age", 7),
        _cat("occupation", 21),
        _cat("zipcode", 3439),
        _cont("TE_age_rating", tags=(Tags.USER,)),
        _cont("TE_gender_rating", tags=(Tags.USER,)),
        _cont("TE_occupation_rating", tags=(Tags.USER,)),
        _cont("TE_zipcode_rating", tags=(Tags.USER,)),
        _cont("TE_movieId_rating", tags=(Tags.ITEM,)),
        _cont("TE_userId_rating", tags=(Tags.USER,)),
        ColumnSchema("rating_binary", tags=(Tags.BINARY_CLASSIFICATION, Tags.TARGET), dtype="int64"),
        ColumnSchema("rating", tags=(Tags.TARGET, Tags.REGRESSION), dtype="float32"),
    ]
    return Schema(cols)


def retrieval_10m_schema(n_items: int = 10_000_000, n_users: int = 1_000_000) -> Schema:
    """Config 3: item-id table 10 M x 64, user-id table 1 M x 64, two small
    categorical features per tower."""
    return Schema([
        _cat("user_id", n_users - 1, tags=(Tags.USER, Tags.USER_ID)),
        _cat("user_age", 9, tags=(Tags.USER,)),
        _cat("user_geo", 2999, tags=(Tags.USER,)),
        _cat("item_id", n_items - 1, tags=(Tags.ITEM, Tags.ITEM_ID)),
        _cat("item_category", 499, tags=(Tags.ITEM,)),
        _cat("item_brand", 19999, tags=(Tags.ITEM,)),
    ])


KNOWN = {
    "criteo": criteo_schema,
    "criteo-tb": criteo_tb_schema,
    "movielens-1m": movielens_1m_schema,
    "retrieval-10m": retrieval_10m_schema,
}


def get_schema(name: str) -> Schema:
    if name not in KNOWN:
        raise ValueError(f"Unknown dataset {name!r}; known: {sorted(KNOWN)}")
    return KNOWN[name]()


def generate_batch(schema: Schema, num_rows: int, seed: int = 1234, index_law: str = "reference",
                   index_dtype=np.int32, min_list_len: int = 1, max_list_len: int = 4,
                   zipf_a: float = 1.05) -> Dict[str, np.ndarray]:
    """Synthetic feature dict in the reference's batch format: scalar columns (B,), ragged list
    columns as `name__values` (nnz,) + `name__offsets` (B+1,) int32
    (merlin/models/tf/transforms/features.py:190-210).  Targets are included under their names."""
    if index_law not in ("reference", "uniform", "zipf"):
        raise ValueError("index_law must be 'reference', 'uniform' or 'zipf'")
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for col in schema:
        dom = col.int_domain
        if col.has_tag(Tags.BINARY_CLASSIFICATION):
            out[col.name] = rng.integers(0, 2, num_rows).astype(col.dtype if col.dtype != "str" else "int64")
            continue
        if dom is not None and dom.max is not None:
            hi = int(dom.max)

            def draw(n, is_id=col.has_tag(Tags.ID)):
                if index_law == "uniform":
                    return rng.integers(0, hi + 1, n)
                if index_law == "zipf":
                    return np.minimum(rng.zipf(zipf_a, n) - 1, hi)
                if is_id:
                    return np.clip(rng.lognormal(3.0, 1.0, n).astype(np.int64), 1, hi)
                return rng.integers(int(dom.min or 0), max(hi, int(dom.min or 0) + 1), n)

            if col.is_list:
                lens = rng.integers(min_list_len, max_list_len + 1, num_rows)
                offs = np.zeros(num_rows + 1, dtype=np.int32)
                np.cumsum(lens, out=offs[1:])
                out[col.name + "__values"] = draw(int(offs[-1])).astype(index_dtype)
                out[col.name + "__offsets"] = offs
            else:
                out[col.name] = draw(num_rows).astype(index_dtype)
        elif col.dtype.startswith("float"):
            out[col.name] = rng.uniform(0.0, 1.0, num_rows).astype(np.float32)
        else:
            out[col.name] = rng.integers(0, 2, num_rows).astype(np.int64)
    return out


def split_targets(schema: Schema, batch: Dict[str, np.ndarray]):
    """(features, targets) — what merlin's Loader yields."""
    tnames = set(schema.select_by_tag(Tags.TARGET).column_names)
    feats = {k: v for k, v in batch.items() if k not in tnames}
    targs = {k: v for k, v in batch.items() if k in tnames}
    return feats, targs
