"""Imperative node traversal (graphlearn/python/sampler/node_sampler.py:85-117): ``get()`` returns the next batch of Nodes of
the LOCAL shard - node tables, or the source / destination ends of an edge table - and raises ``OutOfRangeError`` at the end
of an epoch."""
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


from .neighbor_sampler import _fixed_strategy  # noqa: E402

RandomNodeSampler = _fixed_strategy(NodeSampler, "random", "RandomNodeSampler")
ByOrderNodeSampler = _fixed_strategy(NodeSampler, "by_order", "ByOrderNodeSampler")
ShuffleNodeSampler = _fixed_strategy(NodeSampler, "shuffle", "ShuffleNodeSampler")


def __getattr__(name):          # the edge samplers lived here once; keep the old import path working
    if name in ("EdgeSampler", "RandomEdgeSampler", "ByOrderEdgeSampler", "ShuffleEdgeSampler"):
        from . import edge_sampler
        return getattr(edge_sampler, name)
    raise AttributeError(name)
