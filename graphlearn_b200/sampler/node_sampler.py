"""Imperative seed samplers (graphlearn/python/sampler/node_sampler.py:85-117,
edge_sampler.py): ``get()`` returns the next batch of Nodes / Edges of the LOCAL
shard and raises ``OutOfRangeError`` at the end of an epoch."""
from __future__ import annotations

import torch

from .. import config as _config
from ..data import values as V_
from ..gsl.iterators import SeedIterator

NODE, EDGE_SRC, EDGE_DST = 0, 1, 2


class NodeSampler(object):
    def __init__(self, graph, t, batch_size, strategy="by_order", node_from=NODE):
        assert strategy in ("by_order", "random", "shuffle")
        self._g, self._type, self._bs, self._strategy, self._from = graph, t, batch_size, strategy, node_from
        store, rt = graph.store, graph.runtime
        self._rt = rt
        if node_from == NODE:
            tab = store.nodes[t]
            self._rows = tab.present.nonzero().flatten() if tab.present is not None else torch.arange(tab.n_local, device=rt.device)
            n = int(self._rows.numel())
            self._node_type = t
        else:
            csr = store.edges[t]
            n = csr.n_edges
            self._csr = csr
            self._node_type = csr.src_type if node_from == EDGE_SRC else csr.dst_type
        self._it = SeedIterator(n, batch_size, strategy, rt.device, seed=_config.get().seed + 31 * rt.rank)

    def get(self):
        idx = self._it.next_index()
        if self._from != NODE:
            idx = self._csr.insertion_pos()[idx]     # edge traversal follows insertion (edge id) order
        W, r = self._rt.world, self._rt.rank
        if self._from == NODE:
            vids = self._rows[idx] * W + r
        elif self._from == EDGE_SRC:
            vids = self._csr._row_of_edge[idx] * W + r
        else:
            vids = self._csr.indices.local[idx]
        ids = self._g.to_ids(self._node_type, vids)
        return V_.Nodes(ids, self._node_type, graph=self._g, vids=vids if self._from != NODE or self._node_type == self._type else None)

    @property
    def epoch(self):
        return self._it.epoch

    def state_dict(self):
        return self._it.state_dict()

    def load_state_dict(self, sd):
        self._it.load_state_dict(sd)


class EdgeSampler(object):
    def __init__(self, graph, edge_type, batch_size, strategy="by_order"):
        assert strategy in ("by_order", "random", "shuffle")
        self._g, self._type = graph, edge_type
        self._csr = graph.store.edges[edge_type]
        self._rt = graph.runtime
        self._it = SeedIterator(self._csr.n_edges, batch_size, strategy, self._rt.device,
                                seed=_config.get().seed + 37 * self._rt.rank)

    def get(self):
        csr = self._csr
        idx = csr.insertion_pos()[self._it.next_index()]     # insertion (edge id) order, like the reference
        W, r = self._rt.world, self._rt.rank
        src_v = csr._row_of_edge[idx] * W + r
        dst_v = csr.indices.local[idx]
        src = self._g.to_ids(csr.src_type, src_v)
        dst = self._g.to_ids(csr.dst_type, dst_v)
        return V_.Edges(src, csr.src_type, dst, csr.dst_type, self._type, idx, graph=self._g, src_vids=src_v)

    @property
    def epoch(self):
        return self._it.epoch


from .neighbor_sampler import _fixed_strategy  # noqa: E402

RandomNodeSampler = _fixed_strategy(NodeSampler, "random", "RandomNodeSampler")
ByOrderNodeSampler = _fixed_strategy(NodeSampler, "by_order", "ByOrderNodeSampler")
ShuffleNodeSampler = _fixed_strategy(NodeSampler, "shuffle", "ShuffleNodeSampler")
RandomEdgeSampler = _fixed_strategy(EdgeSampler, "random", "RandomEdgeSampler")
ByOrderEdgeSampler = _fixed_strategy(EdgeSampler, "by_order", "ByOrderEdgeSampler")
ShuffleEdgeSampler = _fixed_strategy(EdgeSampler, "shuffle", "ShuffleEdgeSampler")
