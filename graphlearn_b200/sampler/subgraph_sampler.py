"""SubGraph sampler (graphlearn/python/sampler/subgraph_sampler.py).  Unlike the
reference - whose ``get`` reads a ``_seed_type`` attribute that ``__init__`` never sets
 - this one works."""
from __future__ import annotations

import torch

from .. import config as _config
from ..data import values as V_
from ..gsl.iterators import SeedIterator
from ..ops import rng as rng_ops
from ..ops import subgraph as SUB


class SubGraphSampler(object):
    def __init__(self, graph, seed_type, nbr_type, batch_size=64, strategy="random_node", num_nbrs=None,
                 need_dist=False):
        self._g, self._seed_type, self._nbr_type = graph, seed_type, nbr_type
        self._num_nbrs = list(num_nbrs or [])
        self._need_dist = need_dist
        rt = graph.runtime
        tab = graph.store.nodes[seed_type]
        self._it = SeedIterator(tab.n_local, batch_size, "shuffle" if "random" in strategy else "by_order", rt.device,
                                seed=_config.get().seed + 41 * rt.rank)
        self._rng = rng_ops.DeviceRng(rt, 307)

    def get(self, ids=None, dst_ids=None):
        """``ids`` alone: the sub-graph induced around those seeds (None = the next batch of local seeds); ``ids`` + ``dst_ids``
        (one pair, SEAL): the enclosing sub-graph of the pair, nodes 0 and 1 are src and dst, with ``dist_to_src`` /
        ``dist_to_dst`` when ``need_dist`` (subgraph_sampler.py:56-100 of the reference)."""
        g, rt = self._g, self._g.runtime
        src = dst = None
        if ids is None:
            idx = self._it.next_index()
            seeds = idx * rt.world + rt.rank
        else:
            seeds = g.to_vids(self._seed_type, ids).reshape(-1)
            if dst_ids is not None:
                src = seeds
                dst = g.to_vids(g.store.edges[self._nbr_type].dst_type, dst_ids).reshape(-1)
                seeds = torch.cat([src, dst])
        sg = SUB.induce_subgraph(g.store, self._nbr_type, seeds, self._num_nbrs, need_dist=self._need_dist and src is not None,
                                 src=src, dst=dst, rng=self._rng)
        csr = g.store.edges[self._nbr_type]
        nid = g.to_ids(csr.src_type, sg["nodes"])
        nodes = V_.Nodes(nid, csr.src_type, graph=g, vids=sg["nodes"])
        edges = V_.Edges(nid[sg["row"]], csr.src_type, nid[sg["col"]], csr.dst_type, self._nbr_type, sg["eids"], graph=g,
                         src_vids=sg["nodes"][sg["row"]])
        self._rng.advance(1)
        extra = {}
        if sg.get("dist_to_src") is not None:
            extra = {"dist_to_src": sg["dist_to_src"].cpu().numpy(), "dist_to_dst": sg["dist_to_dst"].cpu().numpy()}
        return V_.SubGraph(torch.stack([sg["row"], sg["col"]]), nodes, edges, **extra)
