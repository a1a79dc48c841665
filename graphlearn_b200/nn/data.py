"""Model-side data containers (graphlearn/python/nn/data.py, nn/tf/data/{egograph,
batchgraph,hetero_batchgraph}.py): ``Data`` (one node/edge set), ``EgoGraph`` (dense fixed
fan-out hops), ``BatchGraph`` / ``HeteroBatchGraph`` (concatenated subgraphs with node offsets).
Everything holds device tensors."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch


class Data(object):
    def __init__(self, ids=None, ints=None, floats=None, strings=None, labels=None, weights=None, **kwargs):
        # the reference's constructor names (nn/data.py Data(ids, int_attrs, float_attrs, string_attrs, ...)) are accepted too
        ints = kwargs.pop("int_attrs", ints)
        floats = kwargs.pop("float_attrs", floats)
        strings = kwargs.pop("string_attrs", strings)
        self.ids, self.ints, self.floats, self.strings = ids, ints, floats, strings
        self.labels, self.weights = labels, weights
        for k, v in kwargs.items():
            setattr(self, k, v)

    @staticmethod
    def from_values(v, flatten=True) -> "Data":
        """From a ``Nodes`` / ``Edges`` result object (device tensors, no host copies)."""
        def t(name):
            x = v.tensor(name)
            if x is None or not isinstance(x, torch.Tensor):
                return x
            if flatten:
                return x.reshape(-1, x.shape[-1]) if name.endswith("_attrs") else x.reshape(-1)
            return x
        ids = v.tensor("ids") if "ids" in v._t else v.tensor("dst_ids")
        return Data(ids.reshape(-1) if flatten else ids, t("int_attrs"), t("float_attrs"), v._t.get("string_attrs"),
                    t("labels"), t("weights"))

    # the reference's attribute names (nn/data.py Data: int_attrs / float_attrs / string_attrs)
    int_attrs = property(lambda self: self.ints)
    float_attrs = property(lambda self: self.floats)
    string_attrs = property(lambda self: self.strings)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self

    def apply(self, func):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, func(v))
        return self


class EgoGraph(object):
    """src + K hops of fixed fan-out neighbours (egograph.py:41-123)."""

    def __init__(self, src: Data, nbr_nodes: Sequence[Data], node_schema=None, nbr_nums: Sequence[int] = (),
                 nbr_edges: Optional[Sequence[Data]] = None, edge_schema=None):
        self._src, self._nbr_nodes, self._nbr_edges = src, list(nbr_nodes), list(nbr_edges or [])
        self._node_schema, self._edge_schema = node_schema, edge_schema
        self._nbr_nums = list(nbr_nums)

    src = property(lambda self: self._src)
    nbr_nodes = property(lambda self: self._nbr_nodes)
    nbr_edges = property(lambda self: self._nbr_edges)
    nbr_nums = property(lambda self: self._nbr_nums)
    node_schema = property(lambda self: self._node_schema)
    edge_schema = property(lambda self: self._edge_schema)

    def hop_node(self, i) -> Data:
        return self._nbr_nodes[i]

    def hop_edge(self, i) -> Data:
        return self._nbr_edges[i]

    def hops(self) -> List[Data]:
        return [self._src] + self._nbr_nodes

    def transform(self, encoders) -> List[torch.Tensor]:
        """Encode every hop with its FeatureEncoder (one per hop or one shared) -> dense inputs."""
        encs = encoders if isinstance(encoders, (list, tuple)) else [encoders] * (len(self._nbr_nodes) + 1)
        return [e(d.floats, d.ints, d.strings) for e, d in zip(encs, self.hops())]


class TemporalGraph(object):
    """EgoGraph with time: centre nodes + their timestamps, and per hop the neighbour nodes, the timestamps of the
    connecting edges and (optionally) the edge features (graphlearn/python/nn/tf/data/temporalgraph.py:29-140).
    ``transform(time_encoder)`` replaces raw times by encoded time SPANS (centre time - edge time), the input the
    temporal layers (``nn.EgoTGATConv``, ``models.TGN``) consume."""

    def __init__(self, src: Data, src_t: torch.Tensor, nbr_nodes: Sequence[Data], nbr_t: Sequence[torch.Tensor],
                 nbr_edges: Optional[Sequence[Data]] = None, nbr_nums: Sequence[int] = (), node_schema=None,
                 edge_schema=None, time_dim: int = 16):
        self.src, self.src_t = src, src_t
        self.nbr_nodes, self.nbr_t, self.nbr_edges = list(nbr_nodes), list(nbr_t), list(nbr_edges or [])
        self.nbr_nums, self.node_schema, self.edge_schema, self.time_dim = list(nbr_nums), node_schema, edge_schema, time_dim

    def hop_node(self, i) -> Data:
        return self.nbr_nodes[i]

    def hop_edge(self, i) -> Data:
        return self.nbr_edges[i]

    def hop_t(self, i) -> torch.Tensor:
        return self.nbr_t[i]

    def time_spans(self):
        """per hop: (time of the hop's parent element) - (edge time), flattened like the hop."""
        spans, parent_t = [], self.src_t.reshape(-1)
        for t, k in zip(self.nbr_t, self.nbr_nums):
            tt = t.reshape(-1)
            spans.append(parent_t.repeat_interleave(k) - tt)
            parent_t = tt
        return spans

    def transform(self, time_encoder):
        enc = [time_encoder(s.clamp(min=0)) for s in self.time_spans()]
        return TemporalGraph(self.src, time_encoder(torch.zeros_like(self.src_t.reshape(-1))), self.nbr_nodes, enc,
                             self.nbr_edges, self.nbr_nums, None, None, self.time_dim)


class BatchGraph(object):
    """Several subgraphs stacked into one big disconnected graph (batchgraph.py): node features
    concatenated, edge_index shifted by the per-graph node offset."""

    def __init__(self, edge_index, nodes: Data, node_schema=None, graph_node_offsets=None, edges: Optional[Data] = None,
                 graph_edge_offsets=None, additional_keys=(), **kwargs):
        self.edge_index, self.nodes, self.edges = edge_index, nodes, edges
        self.node_schema = node_schema
        self.graph_node_offsets, self.graph_edge_offsets = graph_node_offsets, graph_edge_offsets
        self.additional_keys = list(additional_keys)
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return int(self.nodes.size(0)) if isinstance(self.nodes, torch.Tensor) else int(self.nodes.ids.numel())

    @property
    def num_edges(self):
        return int(self.edge_index.size(1))

    @property
    def num_graphs(self):
        return int(self.graph_node_offsets.numel()) - 1

    @property
    def graph_assign(self):
        n = self.graph_node_offsets[1:] - self.graph_node_offsets[:-1]
        return torch.repeat_interleave(torch.arange(n.numel(), device=n.device), n)

    def transform(self, encoder=None) -> "BatchGraph":
        """BatchGraph whose ``nodes`` is the dense feature matrix the sparse convs consume (batchgraph.py ``transform``):
        ``encoder`` is a ``FeatureHandler`` / ``FeatureEncoder`` (or any callable on ``Data``); None = the float attributes."""
        if isinstance(self.nodes, torch.Tensor):
            return self
        if encoder is None:
            x = self.nodes.floats.float()
        elif hasattr(encoder, "_fspec"):                                   # FeatureHandler: takes the Data object
            x = encoder(self.nodes)
        else:
            try:
                x = encoder(self.nodes.floats, self.nodes.ints, self.nodes.strings)
            except TypeError:
                x = encoder(self.nodes)
        extra = {k: getattr(self, k) for k in self.additional_keys if hasattr(self, k)}
        out = BatchGraph(self.edge_index, x, self.node_schema, self.graph_node_offsets, self.edges, self.graph_edge_offsets,
                         self.additional_keys, **extra)
        out.raw_nodes = self.nodes
        return out

    def to_graphs(self):
        """split the batch back into per-graph ``(edge_index [2, m_i] with local node numbering, node ``Data`` slice)`` pairs
        (batchgraph.py ``to_graphs``)"""
        off = self.graph_node_offsets.tolist()
        out = []
        row = self.edge_index[0]
        for i in range(len(off) - 1):
            lo, hi = off[i], off[i + 1]
            m = (row >= lo) & (row < hi)
            nodes = self.nodes[lo:hi] if isinstance(self.nodes, torch.Tensor) else Data(
                *[None if v is None else v[lo:hi] for v in (self.nodes.ids, self.nodes.ints, self.nodes.floats, self.nodes.strings,
                                                            self.nodes.labels, self.nodes.weights)])
            out.append((self.edge_index[:, m] - lo, nodes))
        return out

    @staticmethod
    def from_graphs(graphs, additional_keys=()) -> "BatchGraph":
        """graphs: list of ``SubGraph`` results."""
        dev = graphs[0].edge_index_t.device
        offs, eis, ids, floats, ints, labels = [0], [], [], [], [], []
        for g in graphs:
            n = g.num_nodes
            eis.append(g.edge_index_t + offs[-1])
            ids.append(g.nodes.ids_t.reshape(-1))
            f = g.nodes.tensor("float_attrs")
            if isinstance(f, torch.Tensor):
                floats.append(f.reshape(n, -1))
            i = g.nodes.tensor("int_attrs")
            if isinstance(i, torch.Tensor):
                ints.append(i.reshape(n, -1))
            l = g.nodes.tensor("labels")
            if isinstance(l, torch.Tensor):
                labels.append(l.reshape(-1))
            offs.append(offs[-1] + n)
        nodes = Data(torch.cat(ids), torch.cat(ints) if ints else None, torch.cat(floats) if floats else None, None,
                     torch.cat(labels) if labels else None)
        extra = {}
        for key in additional_keys:
            vals = [torch.as_tensor(getattr(g, key)).to(dev) for g in graphs if hasattr(g, key)]
            if vals:
                extra[key] = torch.cat(vals)
        return BatchGraph(torch.cat(eis, 1), nodes, graph_node_offsets=torch.tensor(offs, device=dev),
                          additional_keys=additional_keys, **extra)


class HeteroBatchGraph(object):
    """dict-of-types variant: edge_index_dict[(src_t, edge_t, dst_t)], nodes_dict[type]."""

    def __init__(self, edge_index_dict: Dict, nodes_dict: Dict[str, Data], graph_node_offsets_dict=None, edges_dict=None,
                 graph_edge_offsets_dict=None, node_schema_dict=None, edge_schema_dict=None):
        self.edge_index_dict, self.nodes_dict = edge_index_dict, nodes_dict
        self.graph_node_offsets_dict = graph_node_offsets_dict or {}
        self.edges_dict, self.graph_edge_offsets_dict = edges_dict or {}, graph_edge_offsets_dict or {}
        self.node_schema_dict, self.edge_schema_dict = node_schema_dict, edge_schema_dict

    @property
    def node_types(self):
        return list(self.nodes_dict)

    @property
    def edge_types(self):
        return list(self.edge_index_dict)

    def num_edges(self, key):
        return int(self.edge_index_dict[key].size(1))

    @property
    def num_graphs(self):
        for off in self.graph_node_offsets_dict.values():
            return int(off.numel()) - 1
        return 0

    def transform(self, encoders=None) -> "HeteroBatchGraph":
        """HeteroBatchGraph whose ``nodes_dict`` holds dense feature matrices: ``encoders`` = {node type: FeatureHandler /
        FeatureEncoder / callable on Data}; types without an encoder keep their float attributes (hetero_batchgraph.py:113-152)"""
        encoders = encoders or {}
        out = {}
        for t, d in self.nodes_dict.items():
            if isinstance(d, torch.Tensor):
                out[t] = d
                continue
            enc = encoders.get(t)
            if enc is None:
                out[t] = d.floats.float()
            elif hasattr(enc, "_fspec"):
                out[t] = enc(d)
            else:
                try:
                    out[t] = enc(d.floats, d.ints, d.strings)
                except TypeError:
                    out[t] = enc(d)
        return HeteroBatchGraph(self.edge_index_dict, out, self.graph_node_offsets_dict, self.edges_dict, self.graph_edge_offsets_dict,
                                self.node_schema_dict, self.edge_schema_dict)

    @staticmethod
    def from_graphs(graphs) -> "HeteroBatchGraph":
        """stack ``nn.HeteroSubGraph`` objects: per node type the ``Data`` rows are concatenated, every ``edge_index_dict[(h, r, t)]``
        is shifted by the per-graph offsets of its two end types (row -> t, col -> h, the HeteroSubGraph convention)"""
        types = list(graphs[0].nodes_dict)
        offs = {t: [0] for t in types}
        for g in graphs:
            for t in types:
                offs[t].append(offs[t][-1] + int(g.nodes_dict[t].ids.numel()))
        nodes = {}
        for t in types:
            ds = [g.nodes_dict[t] for g in graphs]
            cat = lambda name: (torch.cat([getattr(d, name) for d in ds]) if all(isinstance(getattr(d, name), torch.Tensor) for d in ds) else None)  # noqa: E731
            nodes[t] = Data(cat("ids"), cat("ints"), cat("floats"), None, cat("labels"), cat("weights"))
        ei = {}
        for key in graphs[0].edge_index_dict:
            h, _, tt = key
            parts = []
            for i, g in enumerate(graphs):
                e = g.edge_index_dict[key]
                shift = torch.tensor([[offs[tt][i]], [offs[h][i]]], device=e.device, dtype=e.dtype)
                parts.append(e + shift)
            ei[key] = torch.cat(parts, 1)
        dev = next(iter(ei.values())).device if ei else torch.device("cpu")
        return HeteroBatchGraph(ei, nodes, {t: torch.tensor(o, device=dev) for t, o in offs.items()})

    def num_nodes(self, t):
        return int(self.nodes_dict[t].ids.numel())
