"""Heterogeneous subgraph layers and the sub-graph induction hooks of the reference's model layer:

  HeteroConv          per edge type conv + aggregation of results that land on the same node type
                      (graphlearn/python/nn/tf/layers/hetero_conv.py:33-76)
  BipartiteSAGEConv   the ``SubConv`` contract for h != t: node_vec = [x_src, x_dst] (sub_conv.py:24-39)
  LinkPredictor       MLP on pair embeddings -> logits (nn/tf/model/link_predictor.py:31-70)
  HeteroSubGraph      dict-of-types induced subgraph container (nn/hetero_subgraph.py)
  SubGraphInducer / SubGraphProcessor   user hooks turning GSL results into (pos, neg) subgraphs / post-processing
                      one SubGraph (nn/tf/data/subgraph_inducer.py, subgraph_processor.py)

Convention (same as nn/sparse_conv.py): ``edge_index[0]`` = destination (row that receives), ``edge_index[1]`` =
source; for an edge type (h, r, t) the rows index nodes of type t and the columns nodes of type h."""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops.sparse import spmm


class BipartiteSAGEConv(nn.Module):
    def __init__(self, in_src: int, in_dst: int, out_dim: int, agg_type: str = "mean", bias: bool = True):
        super().__init__()
        self.lin_self = nn.Linear(in_dst, out_dim, bias=bias)
        self.lin_nbr = nn.Linear(in_src, out_dim, bias=False)
        self.agg_type = agg_type

    def forward(self, edge_index, node_vec):
        x_src, x_dst = (node_vec, node_vec) if isinstance(node_vec, torch.Tensor) else node_vec
        row, col = edge_index[0], edge_index[1]
        n = x_dst.size(0)
        agg = spmm(x_src.float(), row, col, None, n)
        if self.agg_type == "mean":
            deg = torch.zeros(n, device=row.device).scatter_add(0, row, torch.ones(row.numel(), device=row.device))
            agg = agg / deg.clamp(min=1)[:, None]
        return self.lin_self(x_dst.float()) + self.lin_nbr(agg)


class _Homo(nn.Module):
    """adapts a homogeneous conv(x, edge_index) of nn/sparse_conv.py to the SubConv call order."""

    def __init__(self, conv):
        super().__init__()
        self.conv = conv

    def forward(self, edge_index, node_vec):
        return self.conv(node_vec, edge_index)


class HeteroConv(nn.Module):
    def __init__(self, conv_dict: Dict[Tuple[str, str, str], nn.Module], agg_type: str = "mean"):
        super().__init__()
        assert agg_type in ("sum", "mean", "min", "max")
        self.keys = list(conv_dict)
        self.convs = nn.ModuleList([c if hasattr(c, "lin_self") and isinstance(c, BipartiteSAGEConv) else
                                    (c if _takes_edge_first(c) else _Homo(c)) for c in conv_dict.values()])
        self.agg_type = agg_type

    def forward(self, edge_index_dict, node_vec_dict):
        out = defaultdict(list)
        for key, conv in zip(self.keys, self.convs):
            if key not in edge_index_dict:
                continue
            h, _, t = key
            vec = node_vec_dict[h] if h == t else [node_vec_dict[h], node_vec_dict[t]]
            out[t].append(conv(edge_index_dict[key], vec))
        res = {}
        for t, vs in out.items():
            if len(vs) == 1:
                res[t] = vs[0]
            else:
                st = torch.stack(vs)
                res[t] = {"sum": st.sum(0), "mean": st.mean(0), "min": st.min(0).values, "max": st.max(0).values}[self.agg_type]
        return res


def _takes_edge_first(c) -> bool:
    import inspect
    try:
        params = list(inspect.signature(c.forward).parameters)
    except (TypeError, ValueError):
        return False
    return bool(params) and params[0] == "edge_index"


class LinkPredictor(nn.Module):
    def __init__(self, input_dim: int, num_layers: int = 2, dropout: float = 0.0):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(input_dim, 1 if i == num_layers - 1 else input_dim) for i in range(num_layers)])
        self.dropout = dropout

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for lin in self.layers[:-1]:
            x = F.relu(lin(x))
            if self.dropout and self.training:
                x = F.dropout(x, self.dropout)
        return self.layers[-1](x).squeeze(-1)


class HeteroSubGraph(object):
    """``edge_index_dict[(h, r, t)]`` (rows index t, cols index h) + ``nodes_dict[type]`` (``nn.Data``)."""

    def __init__(self, edge_index_dict: Dict[Tuple[str, str, str], torch.Tensor], nodes_dict: Dict[str, object],
                 edges_dict: Optional[Dict] = None, **kwargs):
        self.edge_index_dict, self.nodes_dict, self.edges_dict = edge_index_dict, nodes_dict, edges_dict or {}
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def node_types(self):
        return list(self.nodes_dict)

    @property
    def edge_types(self):
        return list(self.edge_index_dict)

    def num_nodes(self, t):
        return int(self.nodes_dict[t].ids.numel())

    def num_edges(self, key):
        return int(self.edge_index_dict[key].size(1))

    @property
    def keys(self):
        """names of the attributes that are set (hetero_subgraph.py ``keys``)"""
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith("_")]


class SubGraphInducer(object):
    """Subclass and implement ``induce_func(data_dict) -> (pos_subgraphs, neg_subgraphs | None)``."""

    def __init__(self, use_neg=False, edge_types=None, use_edges=False, node_types=None, addl_types_and_shapes=None):
        self.use_neg, self.edge_types, self.use_edges = use_neg, edge_types, use_edges
        self.node_types, self.addl_types_and_shapes = node_types, addl_types_and_shapes

    def induce_func(self, data_dict):
        raise NotImplementedError


class SubGraphProcessor(object):
    """Subclass and implement ``process_func(subgraph) -> subgraph`` (e.g. SEAL labelling of a sampled SubGraph)."""

    def __init__(self, edge_types=None, use_edges=False, node_types=None, addl_types_and_shapes=None):
        self.edge_types, self.use_edges = edge_types, use_edges
        self.node_types, self.addl_types_and_shapes = node_types, addl_types_and_shapes

    def process_func(self, subgraph):
        return subgraph
