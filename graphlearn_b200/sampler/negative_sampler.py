"""Negative samplers (graphlearn/python/sampler/negative_sampler.py:64-229)."""
from __future__ import annotations

import numpy as np
import torch

from ..data import values as V_
from ..ops import negative as NEG
from ..ops import rng as rng_ops


def _t(x, dev):
    return torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x).to(dev).reshape(-1).to(torch.int64)


class NegativeSampler(object):
    """object_type is an edge type (strategies random | in_degree) or a node type (node_weight)."""

    def __init__(self, graph, object_type, expand_factor, strategy="random"):
        self._g, self._type, self._k, self._strategy = graph, object_type, int(expand_factor), strategy
        self._rng = rng_ops.DeviceRng(graph.runtime, 211)
        self._is_edge = graph.get_topology().is_exist(object_type)
        if strategy == "node_weight" and self._is_edge:
            raise ValueError("node_weight negative sampling takes a node type")

    def get(self, ids):
        g = self._g
        ids_t = _t(ids, g.device)
        gen = self._rng.torch_generator(1)
        if self._is_edge:
            csr = g.store.edges[self._type]
            src_v = g.to_vids(csr.src_type, ids_t)
            neg = NEG.edge_negative(g.store, self._type, src_v, self._k, self._strategy, gen, rng=self._rng, salt=7)
            dst_t = csr.dst_type
        else:
            src_v = g.to_vids(self._type, ids_t)
            neg = NEG.node_weight_negative(g.store, self._type, src_v, self._k, gen)
            dst_t = self._type
        self._rng.advance(1)
        return V_.Nodes(g.to_ids(dst_t, neg), dst_t, shape=(int(ids_t.numel()), self._k), graph=g, vids=neg)


class ConditionalNegativeSampler(object):
    def __init__(self, graph, object_type, expand_factor, strategy="random", batch_share=False, unique=False,
                 int_cols=None, int_props=None, float_cols=None, float_props=None, str_cols=None, str_props=None):
        self._g, self._type, self._k, self._strategy = graph, object_type, int(expand_factor), strategy
        self._cond = {"batch_share": batch_share, "unique": unique, "int_cols": list(int_cols or []),
                      "int_props": list(int_props or []), "float_cols": list(float_cols or []),
                      "float_props": list(float_props or []), "str_cols": list(str_cols or []),
                      "str_props": list(str_props or [])}
        self._rng = rng_ops.DeviceRng(graph.runtime, 223)

    def get(self, src_ids, dst_ids):
        """``object_type`` is an edge type (negatives for (src, dst) pairs of that type) or - with ``strategy="node_weight"`` - a
        NODE type: candidates are that type's nodes, ``src_ids`` only fix the batch shape (conditional_negative_sampler.cc)."""
        g = self._g
        if self._type in g.store.edges:
            csr = g.store.edges[self._type]
            src_type, dst_type = csr.src_type, csr.dst_type
            s = g.to_vids(src_type, _t(src_ids, g.device))
        else:
            dst_type = self._type
            s = _t(src_ids, g.device).reshape(-1)                       # not looked up: ids of whatever type the caller pairs with
        d = g.to_vids(dst_type, _t(dst_ids, g.device))
        neg = NEG.conditional_negative(g.store, self._type, s, d, self._k, self._strategy, self._cond,
                                       self._rng.torch_generator(2))
        self._rng.advance(1)
        return V_.Nodes(g.to_ids(dst_type, neg), dst_type, shape=(int(s.numel()), self._k), graph=g, vids=neg)


from .neighbor_sampler import _fixed_strategy  # noqa: E402

RandomNegativeSampler = _fixed_strategy(NegativeSampler, "random", "RandomNegativeSampler")
InDegreeNegativeSampler = _fixed_strategy(NegativeSampler, "in_degree", "InDegreeNegativeSampler")
NodeWeightNegativeSampler = _fixed_strategy(NegativeSampler, "node_weight", "NodeWeightNegativeSampler")
