"""Query plans (D6): SOURCE / VERTEX_SAMPLER / EDGE_SAMPLER nodes linked into a tree
(dynamic_graph_service/fbs/plan_node.fbs:2-6, query_plan.fbs; JSON form conf/*/install_query.*.json,
produced by the Java GSL client's plan builder).  ``QueryPlan("user").out("click", 10).out("sim", 5)``
is the in-Python shorthand for a chain."""
from __future__ import annotations

from typing import Dict, List, Optional


class PlanNode(object):
    def __init__(self, nid: int, kind: str, vtype: Optional[str] = None, etype: Optional[str] = None, fanout: int = 0,
                 versions: int = 1, parent: Optional[int] = None):
        self.id, self.kind, self.vtype, self.etype = nid, kind, vtype, etype
        self.fanout, self.versions, self.parent = int(fanout), int(versions), parent
        self.children: List[int] = []

    def __repr__(self):
        return "PlanNode(%d %s v=%s e=%s k=%d <- %s)" % (self.id, self.kind, self.vtype, self.etype, self.fanout, self.parent)


class QueryPlan(object):
    def __init__(self, source_type: str):
        self.source_type = source_type
        self.nodes: Dict[int, PlanNode] = {0: PlanNode(0, "SOURCE", vtype=source_type)}
        self._tail = 0

    # ---- chain shorthand (kept from the first version of the service)
    def out(self, edge_type: str, k: int, strategy: str = "topk_by_timestamp"):
        assert strategy == "topk_by_timestamp", "the streaming sampler keeps the k most recent edges"
        nid = max(self.nodes) + 1
        self.add(PlanNode(nid, "EDGE_SAMPLER", etype=edge_type, fanout=k, parent=self._tail))
        self._tail = nid
        return self

    def add(self, node: PlanNode):
        self.nodes[node.id] = node
        if node.parent is not None:
            self.nodes[node.parent].children.append(node.id)
        return node

    @property
    def hops(self):
        """(edge_type, k) of every EDGE_SAMPLER in topological (id) order."""
        return [(n.etype, n.fanout) for _, n in sorted(self.nodes.items()) if n.kind == "EDGE_SAMPLER"]

    def edge_nodes(self):
        return [n for _, n in sorted(self.nodes.items()) if n.kind == "EDGE_SAMPLER"]

    def topo_order(self):
        order, stack = [], [0]
        while stack:
            n = stack.pop()
            order.append(n)
            stack.extend(reversed(self.nodes[n].children))
        return order

    @staticmethod
    def from_json(d: dict, schema) -> "QueryPlan":
        """d: the ``query_plan`` object (or the whole install-query request) of the reference."""
        if "query_plan" in d:
            d = d["query_plan"]
        raw = {n["id"]: n for n in d["plan_nodes"]}
        par = {}
        for n in d["plan_nodes"]:
            for l in n.get("links", []):
                par.setdefault(l["node"], n["id"])
        params = lambda n: {p["key"]: p["value"] for p in n.get("params", [])}  # noqa: E731
        src = [n for n in d["plan_nodes"] if n["kind"] == "SOURCE"][0]
        src_vtype = params(src).get("vtype")
        if src_vtype is None:
            # some install-query files (conf/dblp) leave the SOURCE node without parameters: its vertex type is the one its
            # first downstream sampler starts from
            for l in src.get("links", []):
                src_vtype = params(raw[l["node"]]).get("vtype")
                if src_vtype is not None:
                    break
        if src_vtype is None:
            raise ValueError("install query: cannot determine the vertex type of the SOURCE node")
        plan = QueryPlan(schema.vertex_name[src_vtype])
        remap = {src["id"]: 0}
        for nid in sorted(raw):
            n = raw[nid]
            if n["kind"] == "SOURCE":
                continue
            p = params(n)
            new_id = max(plan.nodes) + 1
            remap[nid] = new_id
            parent = remap.get(par.get(nid, src["id"]), 0)
            if n["kind"] == "EDGE_SAMPLER":
                plan.add(PlanNode(new_id, "EDGE_SAMPLER", vtype=schema.vertex_name.get(p.get("vtype")),
                                  etype=schema.edge_name[p["etype"]], fanout=p.get("fanout", 1), parent=parent))
            else:
                plan.add(PlanNode(new_id, "VERTEX_SAMPLER", vtype=schema.vertex_name[p["vtype"]],
                                  versions=p.get("versions", 1), parent=parent))
        plan.json_ids = remap
        return plan
