"""Minimal `merlin.schema` stand-in: Tags, ColumnSchema, Schema + readers for the reference's
bundled `schema.pbtxt` (text protobuf) and `schema.json` (tensorflow-metadata JSON) files.

merlin-core is not installable here, and the hot path only needs what the reference's model
factories read from a schema: column names, tags, int domains (cardinality = max + 1,
merlin/models/tf/inputs/embedding.py:92-93; shared tables by domain name :670-672), dtypes and
list-ness.  Method names and semantics follow merlin.schema so user code reads the same.
"""
from __future__ import annotations

import json
import re
from dataclasses import dataclass, field, replace
from enum import Enum
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Union


class Tags(str, Enum):
    CATEGORICAL = "categorical"
    CONTINUOUS = "continuous"
    LIST = "list"
    SEQUENCE = "sequence"
    TEXT = "text"
    TIME = "time"
    ID = "id"
    USER = "user"
    USER_ID = "user_id"
    ITEM = "item"
    ITEM_ID = "item_id"
    SESSION = "session"
    SESSION_ID = "session_id"
    CONTEXT = "context"
    TARGET = "target"
    BINARY_CLASSIFICATION = "binary_classification"
    BINARY = "binary"
    MULTI_CLASS_CLASSIFICATION = "multi_class_classification"
    MULTI_CLASS = "multi_class"
    REGRESSION = "regression"
    EMBEDDING = "embedding"

    def __str__(self) -> str:  # so f"{Tags.ITEM}" prints like merlin's Tags
        return f"Tags.{self.name}"


# merlin.schema.tags: compound tags imply their parts (USER_ID -> USER + ID, ...)
_COMPOUND = {
    Tags.USER_ID: (Tags.USER, Tags.ID),
    Tags.ITEM_ID: (Tags.ITEM, Tags.ID),
    Tags.SESSION_ID: (Tags.SESSION, Tags.ID),
    Tags.BINARY_CLASSIFICATION: (Tags.BINARY,),
    Tags.MULTI_CLASS_CLASSIFICATION: (Tags.MULTI_CLASS,),
}

TagLike = Union[str, Tags]


def _norm_tag(t: TagLike) -> str:
    return t.value if isinstance(t, Tags) else str(t).lower()


def _norm_tags(tags: Union[TagLike, Iterable[TagLike], None]) -> List[str]:
    if tags is None:
        return []
    if isinstance(tags, (str, Tags)):
        tags = [tags]
    out: List[str] = []
    for t in tags:
        s = _norm_tag(t)
        if s not in out:
            out.append(s)
        for k, parts in _COMPOUND.items():
            if s == k.value:
                for p in parts:
                    if p.value not in out:
                        out.append(p.value)
    return out


def _tagset(tags) -> set:
    if isinstance(tags, (str, Tags)):
        tags = [tags]
    return {_norm_tag(t) for t in tags}


@dataclass(frozen=True)
class Domain:
    min: Union[int, float] = 0
    max: Optional[Union[int, float]] = None
    name: Optional[str] = None


@dataclass(frozen=True)
class ColumnSchema:
    name: str
    tags: tuple = ()
    dtype: str = "float32"  # numpy dtype name
    is_list: bool = False
    is_ragged: bool = False
    properties: Dict = field(default_factory=dict)

    def __post_init__(self):
        object.__setattr__(self, "tags", tuple(_norm_tags(self.tags)))
        if self.is_ragged and not self.is_list:
            raise ValueError(f"column {self.name!r}: is_ragged requires is_list")

    # -- domains -------------------------------------------------------------------------
    def _domain(self) -> Optional[Domain]:
        d = self.properties.get("domain")
        if d is None:
            return None
        return Domain(d.get("min", 0), d.get("max"), d.get("name"))

    @property
    def int_domain(self) -> Optional[Domain]:
        return self._domain() if self.dtype.startswith(("int", "uint")) else None

    @property
    def float_domain(self) -> Optional[Domain]:
        return self._domain() if self.dtype.startswith("float") else None

    @property
    def value_count(self) -> Optional[Domain]:
        v = self.properties.get("value_count")
        return None if v is None else Domain(v.get("min", 0), v.get("max"))

    # -- copies ---------------------------------------------------------------------------
    def with_name(self, name: str) -> "ColumnSchema":
        return replace(self, name=name)

    def with_tags(self, tags) -> "ColumnSchema":
        return replace(self, tags=tuple(list(self.tags) + _norm_tags(tags)))

    def with_properties(self, properties: Dict) -> "ColumnSchema":
        return replace(self, properties={**self.properties, **properties})

    def with_dtype(self, dtype: str) -> "ColumnSchema":
        return replace(self, dtype=str(dtype))

    def has_tag(self, tag: TagLike) -> bool:
        return _norm_tag(tag) in self.tags


class Schema:
    """Ordered collection of ColumnSchema (insertion order preserved, names unique)."""

    def __init__(self, column_schemas: Union[None, Sequence[Union[ColumnSchema, str]], Dict[str, ColumnSchema]] = None):
        cols: Dict[str, ColumnSchema] = {}
        if isinstance(column_schemas, dict):
            column_schemas = list(column_schemas.values())
        for c in column_schemas or []:
            if isinstance(c, str):
                c = ColumnSchema(c)
            if not isinstance(c, ColumnSchema):
                raise TypeError(f"expected ColumnSchema or str, got {type(c).__name__}")
            if c.name in cols:
                raise ValueError(f"duplicate column {c.name!r} in schema")
            cols[c.name] = c
        self.column_schemas: Dict[str, ColumnSchema] = cols

    # -- selection (merlin.schema.Schema API) ---------------------------------------------
    def select_by_tag(self, tags) -> "Schema":
        want = _tagset(tags)
        return Schema([c for c in self if want & set(c.tags)])

    def excluding_by_tag(self, tags) -> "Schema":
        drop = _tagset(tags)
        return Schema([c for c in self if not (drop & set(c.tags))])

    remove_by_tag = excluding_by_tag

    def select_by_name(self, names) -> "Schema":
        names = [names] if isinstance(names, str) else list(names)
        return Schema([self.column_schemas[n] for n in names if n in self.column_schemas])

    def excluding_by_name(self, names) -> "Schema":
        names = {names} if isinstance(names, str) else set(names)
        return Schema([c for c in self if c.name not in names])

    without = excluding_by_name
    remove_col = excluding_by_name

    @property
    def column_names(self) -> List[str]:
        return list(self.column_schemas.keys())

    @property
    def first(self) -> ColumnSchema:
        return next(iter(self.column_schemas.values()))

    def get(self, name: str, default=None):
        return self.column_schemas.get(name, default)

    def __iter__(self) -> Iterator[ColumnSchema]:
        return iter(self.column_schemas.values())

    def __len__(self) -> int:
        return len(self.column_schemas)

    def __bool__(self) -> bool:
        return len(self.column_schemas) > 0

    def __contains__(self, name) -> bool:
        return name in self.column_schemas

    def __getitem__(self, key):
        if isinstance(key, (list, tuple)):
            return self.select_by_name(key)
        return self.column_schemas[key]

    def __setitem__(self, key: str, col: ColumnSchema):
        self.column_schemas[key] = col

    def __add__(self, other: "Schema") -> "Schema":
        cols = dict(self.column_schemas)
        for c in other:
            cols[c.name] = c
        return Schema(list(cols.values()))

    def __eq__(self, other) -> bool:
        return isinstance(other, Schema) and list(self) == list(other)

    def __repr__(self) -> str:
        return "Schema([" + ", ".join(c.name for c in self) + "])"

    # -- readers ----------------------------------------------------------------------------
    @classmethod
    def from_proto_text(cls, path_or_text: str) -> "Schema":
        text = path_or_text
        if "\n" not in path_or_text and "{" not in path_or_text:
            with open(path_or_text, "r") as f:
                text = f.read()
        msg = _parse_text_proto(text)
        return cls([_feature_to_column(f) for f in msg.get("feature", [])])

    @classmethod
    def from_json(cls, path_or_text: str) -> "Schema":
        text = path_or_text
        if not path_or_text.lstrip().startswith("{"):
            with open(path_or_text, "r") as f:
                text = f.read()
        msg = json.loads(text)
        return cls([_feature_to_column(_json_feature(f)) for f in msg.get("feature", [])])

    @classmethod
    def load(cls, path: str) -> "Schema":
        return cls.from_json(str(path)) if str(path).endswith(".json") else cls.from_proto_text(str(path))

    # -- writer (tensorflow-metadata JSON, merlin/models/utils/schema_utils.py schema_to_tensorflow_metadata_json) --
    def to_json(self, path=None) -> str:
        feats = []
        for c in self:
            is_int = c.dtype.startswith(("int", "uint"))
            f = {"name": c.name, "type": "INT" if is_int else ("BYTES" if c.dtype == "str" else "FLOAT")}
            vc = c.properties.get("value_count")
            if c.is_list:
                v = {}
                if vc and vc.get("min") is not None:
                    v["min"] = str(int(vc["min"]))
                if vc and vc.get("max") is not None:
                    v["max"] = str(int(vc["max"]))
                f["valueCount"] = v
            d = c.properties.get("domain")
            if d is not None:
                if is_int:
                    dom = {"min": str(int(d.get("min", 0) or 0))}
                    if d.get("max") is not None:
                        dom["max"] = str(int(d["max"]))
                    dom["isCategorical"] = "categorical" in c.tags
                else:
                    dom = {"min": float(d.get("min", 0) or 0)}
                    if d.get("max") is not None:
                        dom["max"] = float(d["max"])
                if d.get("name") is not None:
                    dom["name"] = d["name"]
                f["intDomain" if is_int else "floatDomain"] = dom
            if c.tags:
                f["annotation"] = {"tag": list(c.tags)}
            feats.append(f)
        text = json.dumps({"feature": feats}, indent=2)
        if path is not None:
            with open(path, "w") as fh:
                fh.write(text)
        return text


# ---------------------------------------------------------------------------------------------
# text-protobuf reader (enough for tensorflow_metadata Schema files)
# ---------------------------------------------------------------------------------------------
_TOKEN = re.compile(r'\s*(?:(#[^\n]*)|("(?:\\.|[^"\\])*")|([{}:<>])|([^\s{}:<>"]+))')


def _tokens(text: str):
    pos = 0
    n = len(text)
    while pos < n:
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                return
            raise ValueError(f"schema.pbtxt: cannot tokenise at offset {pos}")
        pos = m.end()
        if m.group(1):
            continue
        if m.group(2):
            yield ("str", m.group(2)[1:-1])
        elif m.group(3):
            yield ("sym", m.group(3))
        else:
            yield ("atom", m.group(4))


def _parse_text_proto(text: str) -> Dict[str, list]:
    toks = list(_tokens(text))
    i = 0

    def parse_message(closing: Optional[str]):
        nonlocal i
        msg: Dict[str, list] = {}
        while i < len(toks):
            kind, val = toks[i]
            if kind == "sym" and val == closing:
                i += 1
                return msg
            if kind != "atom":
                raise ValueError(f"schema.pbtxt: expected a field name, got {val!r}")
            name = val
            i += 1
            if i < len(toks) and toks[i] == ("sym", ":"):
                i += 1
            kind, val = toks[i]
            if kind == "sym" and val in "{<":
                i += 1
                value = parse_message("}" if val == "{" else ">")
            else:
                i += 1
                value = val if kind == "str" else _atom(val)
            msg.setdefault(name, []).append(value)
        if closing is not None:
            raise ValueError("schema.pbtxt: unbalanced braces")
        return msg

    return parse_message(None)


def _atom(s: str):
    if s in ("true", "True"):
        return True
    if s in ("false", "False"):
        return False
    try:
        return int(s)
    except ValueError:
        try:
            return float(s)
        except ValueError:
            return s  # enum name


def _one(msg: Dict[str, list], key: str, default=None):
    v = msg.get(key)
    return default if not v else v[0]


def _json_feature(f: dict) -> Dict[str, list]:
    """tensorflow-metadata JSON (camelCase, ints as strings) -> the text-proto dict shape."""

    def conv(d, keymap):
        out = {}
        for k, v in d.items():
            k2 = keymap.get(k, k)
            if isinstance(v, dict):
                v = conv(v, keymap)
            elif isinstance(v, str) and re.fullmatch(r"-?\d+", v):
                v = int(v)
            out[k2] = v if isinstance(v, list) else [v]
        return out

    keymap = {"intDomain": "int_domain", "floatDomain": "float_domain", "valueCount": "value_count",
              "isCategorical": "is_categorical", "extraMetadata": "extra_metadata"}
    return conv(f, keymap)


def _feature_to_column(f: Dict[str, list]) -> ColumnSchema:
    name = _one(f, "name")
    ftype = str(_one(f, "type", "FLOAT")).upper()
    props: Dict = {}
    dtype = {"INT": "int64", "FLOAT": "float32", "BYTES": "str"}.get(ftype, "float32")
    dom = _one(f, "int_domain") if ftype == "INT" else _one(f, "float_domain")
    if isinstance(dom, dict):
        d = {"min": _one(dom, "min", 0), "max": _one(dom, "max")}
        if _one(dom, "name") is not None:
            d["name"] = _one(dom, "name")
        props["domain"] = d
    vc = _one(f, "value_count")
    is_list = isinstance(vc, dict)
    is_ragged = False
    if is_list:
        vmin, vmax = _one(vc, "min", 0), _one(vc, "max")
        props["value_count"] = {"min": vmin, "max": vmax}
        is_ragged = vmax is None or vmin != vmax
    tags: List[str] = []
    ann = _one(f, "annotation")
    if isinstance(ann, dict):
        tags = [str(t) for t in ann.get("tag", [])]
    return ColumnSchema(name=name, tags=tuple(tags), dtype=dtype, is_list=is_list, is_ragged=is_ragged,
                        properties=props)
