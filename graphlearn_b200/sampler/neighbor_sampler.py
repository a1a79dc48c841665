"""Multi-hop neighbour samplers (graphlearn/python/sampler/neighbor_sampler.py:78-213).

``NeighborSampler.get(ids)`` walks a meta-path of edge types with fixed fan-outs
and returns ``Layers`` (1-based ``layer_nodes(i)`` / ``layer_edges(i)``);
``FullNeighborSampler`` returns sparse layers.  Every hop is one launch of the
peer-memory sampling kernel; attributes stay lazy (fetched on first access)."""
from __future__ import annotations

import numpy as np
import torch

from ..data import values as V_
from ..ops import rng as rng_ops
from ..ops import sampling as S


class NeighborSampler(object):
    def __init__(self, graph, meta_path, expand_factor, strategy="random"):
        self._g = graph
        self._meta_path = [meta_path] if isinstance(meta_path, str) else list(meta_path)
        ef = [expand_factor] if isinstance(expand_factor, int) else list(expand_factor)
        if len(ef) != len(self._meta_path):
            raise ValueError("The meta_path must have the same number of elements as num_at_each_hop")
        self._expand = ef
        self._strategy = strategy
        self._rng = rng_ops.DeviceRng(graph.runtime, 101)
        topo = graph.get_topology()
        for e in self._meta_path:
            if not topo.is_exist(e):
                raise ValueError("edge type %r not in graph" % (e,))

    def _src_vids(self, etype, ids):
        csr = self._g.store.edges[etype]
        return self._g.to_vids(csr.src_type, ids), csr

    def get(self, ids):
        g = self._g
        ids_t = torch.as_tensor(np.asarray(ids) if not isinstance(ids, torch.Tensor) else ids).to(g.device).reshape(-1)
        layers = V_.Layers()
        cur_ids = ids_t
        cur_v = None
        for hop, (etype, k) in enumerate(zip(self._meta_path, self._expand)):
            csr = g.store.edges[etype]
            if self._strategy == "in_degree":
                g.store.ensure_indegree_weights(etype)
            src_v = cur_v if cur_v is not None else g.to_vids(csr.src_type, cur_ids)
            nbr, eid = S.sample_neighbors(csr, src_v.reshape(-1), k, self._strategy, rng=self._rng, salt=hop + 1)
            B = int(src_v.numel())
            nbr_ids = g.to_ids(csr.dst_type, nbr)
            nodes = V_.Nodes(nbr_ids, csr.dst_type, shape=(B, k), graph=g, vids=nbr)
            edges = V_.Edges(cur_ids.reshape(-1, 1).expand(B, k), csr.src_type, nbr_ids, csr.dst_type, etype, eid,
                             shape=(B, k), graph=g, src_vids=src_v.reshape(-1, 1).expand(B, k))
            layers.append_layer(V_.Layer(nodes, edges, shape=(B, k)))
            cur_ids, cur_v = nbr_ids.reshape(-1), nbr.reshape(-1)
        self._rng.advance(1)
        return layers


class FullNeighborSampler(NeighborSampler):
    def __init__(self, graph, meta_path, expand_factor=0, strategy="full"):
        mp = [meta_path] if isinstance(meta_path, str) else list(meta_path)
        ef = [expand_factor] * len(mp) if isinstance(expand_factor, int) else list(expand_factor)
        super().__init__(graph, mp, ef, "random")
        self._strategy = "full"

    def get(self, ids):
        g = self._g
        ids_t = torch.as_tensor(np.asarray(ids) if not isinstance(ids, torch.Tensor) else ids).to(g.device).reshape(-1)
        layers = V_.Layers()
        cur_ids, cur_v = ids_t, None
        for etype, k in zip(self._meta_path, self._expand):
            csr = g.store.edges[etype]
            src_v = cur_v if cur_v is not None else g.to_vids(csr.src_type, cur_ids)
            vals, eids, offs = S.sample_full(csr, src_v.reshape(-1), cap=max(int(k), 0))
            counts = offs[1:] - offs[:-1]
            B = int(src_v.numel())
            maxd = int(counts.max().item()) if B else 0
            nbr_ids = g.to_ids(csr.dst_type, vals)
            nodes = V_.SparseNodes(nbr_ids, counts, (B, maxd), csr.dst_type, graph=g, vids=vals)
            src_rep = torch.repeat_interleave(cur_ids, counts)
            edges = V_.SparseEdges(src_rep, csr.src_type, nbr_ids, csr.dst_type, etype, counts, (B, maxd), edge_ids=eids,
                                   graph=g)
            layers.append_layer(V_.Layer(nodes, edges, shape=(B, maxd)))
            cur_ids, cur_v = nbr_ids, vals
        return layers


def _fixed_strategy(base, strategy, doc):
    """Named strategy subclasses of the reference (neighbor_sampler.py:143-163 etc.): same constructor minus
    the ``strategy`` argument."""
    def __init__(self, graph, *args, **kwargs):
        kwargs.pop("strategy", None)
        base.__init__(self, graph, *args, strategy=strategy, **kwargs)
    return type(doc, (base,), {"__init__": __init__, "__doc__": "%s fixed to strategy %r." % (base.__name__, strategy)})


RandomNeighborSampler = _fixed_strategy(NeighborSampler, "random", "RandomNeighborSampler")
RandomWithoutReplacementNeighborSampler = _fixed_strategy(NeighborSampler, "random_without_replacement",
                                                          "RandomWithoutReplacementNeighborSampler")
EdgeWeightNeighborSampler = _fixed_strategy(NeighborSampler, "edge_weight", "EdgeWeightNeighborSampler")
TopkNeighborSampler = _fixed_strategy(NeighborSampler, "topk", "TopkNeighborSampler")
InDegreeNeighborSampler = _fixed_strategy(NeighborSampler, "in_degree", "InDegreeNeighborSampler")
