"""Graph schema + service options of the streaming service (D1).

File formats are the reference's: the JSON schema with ``attr_defs / vertex_defs / edge_defs /
edge_relation_defs`` (dynamic_graph_service/conf/u2i/schema.u2i.json, src/common/schema.{h,cc}) and
YAML option files (src/common/options.{h,cc}, conf/ut/*.yml) - only the options that still mean
something without Kafka/RocksDB are interpreted, the rest is kept verbatim in ``Options.raw``."""
from __future__ import annotations

import json
from typing import Dict, List, Optional

_LIST_TYPES = {"INT32_LIST", "INT64_LIST", "FLOAT32_LIST", "FLOAT64_LIST"}


class AttrDef(object):
    def __init__(self, type_id: int, name: str, value_type: str):
        self.type, self.name, self.value_type = int(type_id), name, value_type

    @property
    def is_list(self):
        return self.value_type in _LIST_TYPES

    @property
    def is_float(self):
        return self.value_type.startswith("FLOAT")


class Schema(object):
    def __init__(self, d: dict):
        self.raw = d
        self.attrs: Dict[int, AttrDef] = {a["type"]: AttrDef(a["type"], a["name"], a["value_type"]) for a in d.get("attr_defs", [])}
        self.attr_by_name = {a.name: a for a in self.attrs.values()}
        self.vertex_name = {v["vtype"]: v["name"] for v in d.get("vertex_defs", [])}
        self.vertex_id = {n: t for t, n in self.vertex_name.items()}
        self.vertex_attrs = {v["name"]: [self.attrs[t] for t in v.get("attr_types", [])] for v in d.get("vertex_defs", [])}
        self.edge_name = {e["etype"]: e["name"] for e in d.get("edge_defs", [])}
        self.edge_id = {n: t for t, n in self.edge_name.items()}
        self.edge_attrs = {e["name"]: [self.attrs[t] for t in e.get("attr_types", [])] for e in d.get("edge_defs", [])}
        self.relations = {}
        for r in d.get("edge_relation_defs", []):
            self.relations[self.edge_name[r["etype"]]] = (self.vertex_name[r["src_vtype"]], self.vertex_name[r["dst_vtype"]])

    @staticmethod
    def from_json(path: str) -> "Schema":
        with open(path) as f:
            return Schema(json.load(f))

    def to_service_schema(self, capacity: int = 1024, feat_dims: Optional[Dict[str, int]] = None) -> dict:
        """-> the dict ``DynamicGraphService`` takes; vertex tables start at ``capacity`` rows and grow."""
        feat_dims = feat_dims or {}
        return {"vertices": {n: {"count": capacity, "feat_dim": feat_dims.get(n, 0)} for n in self.vertex_id},
                "edges": {e: {"src": s, "dst": d} for e, (s, d) in self.relations.items()}}


class Options(object):
    """YAML option file (``key: value`` / nested maps) -> attribute access with defaults."""

    DEFAULTS = {"worker-type": "Sampling", "record-polling": {"process-concurrency": 10, "retry-interval-ms": 1000},
                "sample-store": {"ttl-hours": 1200, "in-memory-mode": True},
                "coordinator-client": {"heartbeat-interval-in-sec": 5}, "http-port": 0,
                "checkpoint": {"path": ".", "keep": 3}}

    def __init__(self, raw: Optional[dict] = None):
        self.raw = dict(self.DEFAULTS)
        for k, v in (raw or {}).items():
            if isinstance(v, dict) and isinstance(self.raw.get(k), dict):
                self.raw[k] = {**self.raw[k], **v}
            else:
                self.raw[k] = v

    @staticmethod
    def from_yaml(path: str) -> "Options":
        import yaml
        with open(path) as f:
            return Options(yaml.safe_load(f) or {})

    def get(self, dotted: str, default=None):
        cur = self.raw
        for p in dotted.split("."):
            if not isinstance(cur, dict) or p not in cur:
                return default
            cur = cur[p]
        return cur
