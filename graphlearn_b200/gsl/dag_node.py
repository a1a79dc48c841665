"""GSL query nodes: the Gremlin-like traversal builder.

API parity with graphlearn/python/gsl/dag_node.py:164-305,458-778:
``batch / shuffle / alias / sample / by / filter / where / each / values`` and the
traversals ``outV inV outE inE outNeg inNeg Neg random_walk SubGraph``.

Unlike the reference, building a query does not serialise a DAG proto for a
server: the nodes form a *static sampling plan* that ``gsl.executor`` walks on
the device every ``Dataset.next()``.  Every aliased traversal node implicitly
carries its attribute lookup and degree fetch (the reference adds LookupNodes
/ GetDegree child nodes, dag_node.py:71-79,558-564) - here they are lazy
device gathers on the resulting ``Nodes`` / ``Edges`` objects.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

from .. import errors

NODE, EDGE_SRC, EDGE_DST = 0, 1, 2

NEIGHBOR_STRATEGIES = ("random", "random_without_replacement", "topk", "in_degree", "edge_weight", "full")
NEGATIVE_STRATEGIES = ("random", "in_degree", "node_weight")
TRAVERSE_STRATEGIES = ("by_order", "random", "shuffle")


class DagNode(object):
    def __init__(self, dag, op_name="", params=None, upstream: Optional["DagNode"] = None):
        self._dag = dag
        self._graph = dag.graph
        self._op_name = op_name
        self._params: Dict[str, object] = dict(params or {})
        self._upstream = upstream
        self._downstreams: List["DagNode"] = []
        self._alias: Optional[str] = None
        self._type = None          # node type (vertex nodes) or edge type (edge nodes)
        self._base_type = None     # unmasked node type: the id space used by edges
        self._shape = None
        self._sparse = False
        self._filter: Optional["DagNode"] = None
        self._nid = dag.next_id()
        if upstream is not None:
            upstream._downstreams.append(self)

    # ---- introspection
    nid = property(lambda self: self._nid)
    op_name = property(lambda self: self._op_name)
    type = property(lambda self: self._type)
    shape = property(lambda self: self._shape)
    sparse = property(lambda self: self._sparse)
    params = property(lambda self: self._params)
    upstream = property(lambda self: self._upstream)
    downstreams = property(lambda self: self._downstreams)
    pos_downstreams = property(lambda self: [d for d in self._downstreams if not d._params.get("negative")])
    neg_downstreams = property(lambda self: [d for d in self._downstreams if d._params.get("negative")])

    @property
    def decoder(self):
        return self._graph.get_node_decoder(self._type)

    def get_alias(self):
        return self._alias

    def set_output_type(self, t, base_type=None):
        self._type = t
        self._base_type = base_type if base_type is not None else t

    # ---- modifiers
    def alias(self, alias):
        if not isinstance(alias, str) or not alias:
            raise ValueError("alias must be a non-empty string.")
        self._alias = alias
        self._dag.add_node(alias, self)
        return self

    def batch(self, batch_size):
        if not isinstance(batch_size, int) or batch_size <= 0:
            raise ValueError("batch_size must be a positive integer.")
        self._params["batch_size"] = batch_size
        self._params.setdefault("strategy", "by_order")
        self._shape = (batch_size,)
        return self

    def shuffle(self, traverse=False):
        self._params["strategy"] = "shuffle" if traverse else "random"
        return self

    def sample(self, count):
        if not isinstance(count, int):
            raise ValueError("sample count must be an integer.")
        self._params["neighbor_count"] = count
        up = self._upstream._shape if self._upstream is not None else None
        n = 1
        for s in (up or ()):
            n *= s
        self._shape = (n, count) if up else (None, count)
        return self

    def by(self, strategy):
        neg = bool(self._params.get("negative"))
        allowed = NEGATIVE_STRATEGIES if neg else NEIGHBOR_STRATEGIES
        if not neg:
            from ..ops.sampling import registered_samplers
            allowed = tuple(allowed) + tuple(registered_samplers())            # gl.register_sampler
        if strategy not in allowed:
            raise ValueError("strategy must be one of {}, got {!r}".format(allowed, strategy))
        self._params["strategy"] = strategy
        if strategy == "full":
            self._sparse = True
        return self

    def filter(self, target):
        """Exclude sampled neighbours equal to the (per-row) ids of `target`
        (an upstream DagNode or its alias) - e.g. drop the positive dst when sampling."""
        if isinstance(target, str):
            target = self._dag.get_node(target)
        if not isinstance(target, DagNode):
            raise ValueError("filter target must be a DagNode or an alias.")
        self._filter = target
        return self

    def where(self, target, condition=None):
        """Conditional negative sampling (dag_node.py:233-292): negatives share the selected
        attribute columns with `target`."""
        if isinstance(target, str):
            target = self._dag.get_node(target)
        if not self._params.get("negative"):
            raise ValueError("where() is only valid after outNeg/inNeg/Neg.")
        cond = dict(condition or {})
        self._params["conditional"] = True
        self._params["dst_node"] = target
        self._params["condition"] = {
            "batch_share": bool(cond.get("batch_share", False)),
            "unique": bool(cond.get("unique", False)),
            "int_cols": list(cond.get("int_cols", [])), "int_props": list(cond.get("int_props", [])),
            "float_cols": list(cond.get("float_cols", [])), "float_props": list(cond.get("float_props", [])),
            "str_cols": list(cond.get("str_cols", [])), "str_props": list(cond.get("str_props", [])),
        }
        return self

    def each(self, func: Callable[["DagNode"], object]):
        func(self)
        return self

    def remove_property(self):
        self._params["no_property"] = True
        return self

    def values(self, func=lambda x: x):
        self._dag.set_ready(func)
        return self._dag


class TraverseVertexDagNode(DagNode):
    def _hop(self, cls, op, edge_type, params):
        return cls(self._dag, op_name=op, params=params, upstream=self)

    def _edge_types(self, edge_type, reverse):
        topo = self._graph.get_topology()
        if not topo.is_exist(edge_type):
            raise ValueError("edge type %r not in graph" % (edge_type,))
        return (topo.get_dst_type(edge_type), topo.get_src_type(edge_type)) if reverse else \
            (topo.get_src_type(edge_type), topo.get_dst_type(edge_type))

    def outV(self, edge_type=None):
        frm, to = self._edge_types(edge_type, False)
        n = TraverseVertexDagNode(self._dag, "Sampler", {"edge_type": edge_type, "direction": "out",
                                                         "strategy": "random"}, upstream=self)
        n.set_output_type(to)
        return n

    def inV(self, edge_type=None):
        frm, to = self._edge_types(edge_type, True)
        n = TraverseVertexDagNode(self._dag, "Sampler", {"edge_type": edge_type, "direction": "in",
                                                         "strategy": "random"}, upstream=self)
        n.set_output_type(to)
        return n

    def outE(self, edge_type):
        self._edge_types(edge_type, False)
        n = TraverseEdgeDagNode(self._dag, "Sampler", {"edge_type": edge_type, "direction": "out",
                                                       "strategy": "random", "emit": "edges"}, upstream=self)
        n._type = edge_type
        return n

    def inE(self, edge_type):
        self._edge_types(edge_type, True)
        n = TraverseEdgeDagNode(self._dag, "Sampler", {"edge_type": edge_type, "direction": "in",
                                                       "strategy": "random", "emit": "edges"}, upstream=self)
        n._type = edge_type
        return n

    def outNeg(self, edge_type):
        frm, to = self._edge_types(edge_type, False)
        n = TraverseNegVertexDagNode(self._dag, "NegativeSampler", {"edge_type": edge_type, "direction": "out",
                                                                    "strategy": "random", "negative": True},
                                     upstream=self)
        n.set_output_type(to)
        return n

    def inNeg(self, edge_type):
        frm, to = self._edge_types(edge_type, True)
        n = TraverseNegVertexDagNode(self._dag, "NegativeSampler", {"edge_type": edge_type, "direction": "in",
                                                                    "strategy": "random", "negative": True},
                                     upstream=self)
        n.set_output_type(to)
        return n

    def Neg(self, node_type):
        n = TraverseNegVertexDagNode(self._dag, "NegativeSampler", {"node_type": node_type, "strategy": "node_weight",
                                                                    "negative": True}, upstream=self)
        n.set_output_type(node_type)
        return n

    def random_walk(self, edge_type, walk_len=1, p=1.0, q=1.0):
        frm, to = self._edge_types(edge_type, False)
        n = TraverseVertexDagNode(self._dag, "RandomWalk", {"edge_type": edge_type, "walk_len": int(walk_len),
                                                            "p": float(p), "q": float(q)}, upstream=self)
        n.set_output_type(to)
        up = self._shape or (None,)
        n._shape = (up[0], int(walk_len))
        return n

    def SubGraph(self, edge_type, num_nbrs=None, need_dist=False):
        """Induce a subgraph among (optionally expanded) batch nodes (dag_node.py:532-556)."""
        n = SubGraphDagNode(self._dag, params={"nbr_type": edge_type, "num_nbrs": list(num_nbrs or []),
                                               "need_dist": need_dist, "from_upstream": True}, upstream=self)
        n._type = edge_type
        return n


class TraverseNegVertexDagNode(TraverseVertexDagNode):
    pass


class TraverseEdgeDagNode(DagNode):
    """Edges produced by outE/inE; ``inV()`` / ``outV()`` expose the end points."""

    def _endpoint(self, which):
        n = FakeNode(self._dag, self, which)
        return n

    def inV(self):
        return self._endpoint("dst")

    def outV(self):
        return self._endpoint("src")


class TraverseSourceEdgeDagNode(TraverseEdgeDagNode):
    """Root of E(): iterates edges of a type."""

    def outV(self):
        return self._endpoint("src")

    def inV(self, edge_type=None):
        return self._endpoint("dst")

    def SubGraph(self, edge_type=None, num_nbrs=None, need_dist=False):
        n = SubGraphDagNode(self._dag, params={"nbr_type": edge_type or self._params["edge_type"],
                                               "num_nbrs": list(num_nbrs or []), "need_dist": need_dist,
                                               "from_upstream": True, "from_edges": True}, upstream=self)
        n._type = edge_type or self._params["edge_type"]
        return n


class FakeNode(TraverseVertexDagNode):
    """Endpoint view of an edge node (reference: dag_node.py:761-790): no op of its own."""

    def __init__(self, dag, edge_node: DagNode, which: str):
        super().__init__(dag, op_name="EdgeEndpoint", params={"which": which}, upstream=edge_node)
        topo = dag.graph.get_topology()
        et = edge_node._params.get("edge_type")
        rev = edge_node._params.get("direction") == "in" or edge_node._params.get("reverse")
        st, dt = topo.get_src_type(et), topo.get_dst_type(et)
        if rev:
            st, dt = dt, st
        self.set_output_type(st if which == "src" else dt)
        self._shape = edge_node._shape


class SubGraphDagNode(DagNode):
    def __init__(self, dag, op_name="SubGraphSampler", params=None, upstream=None):
        super().__init__(dag, op_name, params, upstream)
        self._type = (params or {}).get("nbr_type")
