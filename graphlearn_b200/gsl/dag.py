"""Dag = one GSL query (graphlearn/python/gsl/dag.py:25-130): a registry of aliased
nodes plus the root; ``set_ready`` freezes it into an executable plan."""
from __future__ import annotations

import threading

_LOCK = threading.Lock()
_COUNT = [0]


def get_dag_name():
    with _LOCK:
        _COUNT[0] += 1
        return "dag_%d" % _COUNT[0]


class Dag(object):
    def __init__(self, graph):
        self.graph = graph
        self.name = get_dag_name()
        self.root = None
        self._alias_to_node = {}
        self._aliases = []
        self._ready = False
        self._value_func = None
        self._next_id = 0

    def next_id(self):
        self._next_id += 1
        return self._next_id

    def add_node(self, alias, node, temp=False):
        if alias in self._alias_to_node and self._alias_to_node[alias] is not node:
            raise ValueError("alias {!r} is already used in this query.".format(alias))
        self._alias_to_node[alias] = node
        if alias not in self._aliases:
            self._aliases.append(alias)

    def get_node(self, alias):
        node = self._alias_to_node.get(alias)
        if node is None:
            raise ValueError("alias {!r} not found in the query.".format(alias))
        return node

    def list_alias(self):
        return list(self._aliases)

    @property
    def node_types(self):
        return sorted({n.type for n in self._alias_to_node.values() if isinstance(n.type, str)})

    def is_ready(self):
        return self._ready

    def set_ready(self, func=lambda x: x):
        if self.root is None:
            raise ValueError("query has no root (start it with g.V() or g.E()).")
        self._ready = True
        self._value_func = func

    @property
    def value_func(self):
        return self._value_func

    def __str__(self):
        return "Dag(%s, aliases=%s)" % (self.name, self._aliases)

    # ------------------------------------------------------------------ DagDef: the query as plain data (server mode)
    def to_def(self) -> dict:
        """Serialisable description of the query (the reference's DagDef proto, graphlearn/proto/dag.proto:7-24)."""
        from .dag_node import DagNode
        nodes, seen = [], set()

        def ref(v):
            return {"__node__": v.nid} if isinstance(v, DagNode) else v

        def walk(n):
            if n.nid in seen:
                return
            seen.add(n.nid)
            nodes.append({"nid": n.nid, "cls": type(n).__name__, "op": n.op_name, "params": {k: ref(v) for k, v in n.params.items()},
                          "alias": n.get_alias(), "type": n._type, "base_type": n._base_type, "shape": n._shape, "sparse": n._sparse,
                          "upstream": None if n.upstream is None else n.upstream.nid,
                          "filter": None if n._filter is None else n._filter.nid})
            for d in n.downstreams:
                walk(d)
        walk(self.root)
        return {"root": self.root.nid, "nodes": nodes, "ready": self._ready}

    @staticmethod
    def from_def(graph, d: dict) -> "Dag":
        from . import dag_node as DN
        dag = Dag(graph)
        by_id = {}
        for nd in sorted(d["nodes"], key=lambda x: x["nid"]):          # construction order: upstream first
            cls = getattr(DN, nd["cls"])
            n = object.__new__(cls)
            n._dag, n._graph = dag, graph
            n._op_name, n._params = nd["op"], dict(nd["params"])
            n._upstream = by_id.get(nd["upstream"])
            n._downstreams = []
            n._alias, n._type, n._base_type = nd["alias"], nd["type"], nd["base_type"]
            n._shape = tuple(nd["shape"]) if nd["shape"] is not None else None
            n._sparse, n._filter = nd["sparse"], None
            n._nid = nd["nid"]
            dag._next_id = max(dag._next_id, nd["nid"])
            if n._upstream is not None:
                n._upstream._downstreams.append(n)
            by_id[nd["nid"]] = n
            if n._alias:
                dag.add_node(n._alias, n)
        for nd in d["nodes"]:
            n = by_id[nd["nid"]]
            if nd["filter"] is not None:
                n._filter = by_id[nd["filter"]]
            for k, v in list(n._params.items()):
                if isinstance(v, dict) and "__node__" in v:
                    n._params[k] = by_id[v["__node__"]]
        dag.root = by_id[d["root"]]
        if d.get("ready"):
            dag.set_ready()
        return dag
