"""EgoGraph data loaders of the reference's TF examples (examples/tf/ego_data_loader.py + ego_sage/ego_sage_data_loader.py):
a loader owns the GSL query of one data split and hands out EgoGraphs by alias.

    train = EgoSAGESupervisedDataLoader(g, gl.Mask.TRAIN, 'random', batch_size=128, node_type='i', edge_type='e', nbrs_num=[10, 5])
    for ego in train:                       # one epoch, ends on OutOfRange
        logits = model([h.floats for h in ego.hops()], ego.nbr_nums); labels = ego.src.labels

The reference builds TF iterators once and re-initialises them per epoch; here every ``__iter__`` is one epoch of the query.
"""
from common import sys  # noqa: F401  (repo root on sys.path)

import graphlearn_b200 as gl
from graphlearn_b200 import nn as glnn

_PREFIX = {gl.Mask.TRAIN: "train", gl.Mask.TEST: "test", gl.Mask.VAL: "val", gl.Mask.NONE: "all"}


class EgoDataLoader(object):
    def __init__(self, graph, mask=None, sampler="random", batch_size=128, window=10):
        self._graph = graph
        self._mask = gl.Mask[mask.upper()] if isinstance(mask, str) else (mask or gl.Mask.NONE)
        self._sampler, self._batch_size = sampler, batch_size
        self._q = self._query(graph)
        self._dataset = glnn.Dataset(self._q, window=window)
        self._res = None

    # ---- one batch at a time (the reference's iterator / data_dict)
    def next(self):
        """advance to the next batch; raises gl.OutOfRangeError at the end of the epoch"""
        self._res = self._dataset.next()
        return self

    @property
    def data_dict(self):
        return {k: glnn.Data.from_values(v) for k, v in self._res.items() if hasattr(v, "_t")}

    def data(self, key):
        return glnn.Data.from_values(self._res[key])

    def __getitem__(self, key):
        return self.data(key)

    def get_egograph(self, key, neighbors=None):
        return self._dataset.get_egograph(key, neighbors, res=self._res)

    def __iter__(self):
        while True:
            try:
                self.next()
            except gl.OutOfRangeError:
                return
            yield self.src_ego

    @property
    def src_ego(self):
        raise NotImplementedError

    @property
    def dst_ego(self):
        raise NotImplementedError

    @property
    def neg_dst_ego(self):
        raise NotImplementedError

    def _query(self, graph):
        raise NotImplementedError


class EgoSAGESupervisedDataLoader(EgoDataLoader):
    """V(node_type, mask).batch(B) -> K sampled hops over edge_type (ego_sage_data_loader.py:38-62)."""

    def __init__(self, graph, mask=gl.Mask.TRAIN, sampler="random", batch_size=128, window=10, node_type="i", edge_type="e",
                 nbrs_num=None, hops_num=None):
        self._node_type, self._edge_type, self._nbrs_num = node_type, edge_type, list(nbrs_num or [])
        assert hops_num is None or hops_num == len(self._nbrs_num)
        super().__init__(graph, mask, sampler, batch_size, window)

    @property
    def prefix(self):
        return _PREFIX[self._mask]

    @property
    def src_ego(self):
        return self.get_egograph(self.prefix)

    def _query(self, graph):
        q = graph.V(self._node_type, mask=self._mask).batch(self._batch_size).shuffle(traverse=True).alias(self.prefix)
        for idx, hop in enumerate(self._nbrs_num):
            q = q.outV(self._edge_type).sample(hop).by(self._sampler).alias("%s_hop%d" % (self.prefix, idx))
        return q.values()


class EgoSAGEUnsupervisedDataLoader(EgoDataLoader):
    """E(edge_type).batch(B): src / dst / sampled negative dst, each with K sampled hops (ego_sage_data_loader.py:65-130)."""

    def __init__(self, graph, mask=gl.Mask.TRAIN, sampler="random", neg_sampler="random", batch_size=128, window=10, node_type="i",
                 edge_type="e", nbrs_num=None, neg_num=5):
        self._neg_sampler, self._node_type, self._edge_type = neg_sampler, node_type, edge_type
        self._nbrs_num, self._neg_num = list(nbrs_num or []), neg_num
        super().__init__(graph, mask, sampler, batch_size, window)

    src_ego = property(lambda self: self.get_egograph("src"))
    dst_ego = property(lambda self: self.get_egograph("dst"))
    neg_dst_ego = property(lambda self: self.get_egograph("neg_dst"))

    def _hops(self, q, prefix):
        for idx, hop in enumerate(self._nbrs_num):
            q = q.outV(self._edge_type).sample(hop).by(self._sampler).alias("%s_hop%d" % (prefix, idx))

    def _query(self, graph):
        seed = graph.E(self._edge_type).batch(self._batch_size).shuffle(traverse=True)
        src, dst = seed.outV().alias("src"), seed.inV().alias("dst")
        neg = src.outNeg(self._edge_type).sample(self._neg_num).by(self._neg_sampler).alias("neg_dst")
        for node, prefix in ((src, "src"), (dst, "dst"), (neg, "neg_dst")):
            self._hops(node, prefix)
        return seed.values()
