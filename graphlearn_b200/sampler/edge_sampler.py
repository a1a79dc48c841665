"""Imperative edge traversal (graphlearn/python/sampler/edge_sampler.py:29-98): ``get()`` returns the next batch of Edges of the
LOCAL shard in insertion (edge id) order, at random, or as a shuffled epoch, and raises ``OutOfRangeError`` at the end of an
epoch."""
from __future__ import annotations

from .. import config as _config
from ..data import values as V_
from ..gsl.iterators import SeedIterator
from .neighbor_sampler import _fixed_strategy


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

    def state_dict(self):
        return self._it.state_dict()

    def load_state_dict(self, sd):
        self._it.load_state_dict(sd)


RandomEdgeSampler = _fixed_strategy(EdgeSampler, "random", "RandomEdgeSampler")
ByOrderEdgeSampler = _fixed_strategy(EdgeSampler, "by_order", "ByOrderEdgeSampler")
ShuffleEdgeSampler = _fixed_strategy(EdgeSampler, "shuffle", "ShuffleEdgeSampler")
