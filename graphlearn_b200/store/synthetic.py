"""Synthetic graphs of a named shape, generated directly in HBM, shard by shard.

There is no network for datasets, so the benchmark graphs (BASELINE.json
configs: ogbn-products-shaped 2.4M nodes / 123M edges / 100-d, papers100M-
shaped, ...) are random graphs with the same node / edge / feature counts and
a skewed (log-normal) degree distribution.  Every rank generates only the
edges whose source it owns (src % world == rank), so generation scales with
the number of GPUs and never touches the host.
"""
from __future__ import annotations

import torch

from ..parallel.runtime import Runtime
from .shards import CsrShard, IdMap, NodeTable


def make_sharded_graph(rt: Runtime, num_nodes: int, num_edges: int, feat_dim: int, num_classes: int,
                       feature_dtype: torch.dtype = torch.float32, seed: int = 0, weighted: bool = False,
                       ntype: str = "n", etype: str = "e", skew: float = 1.0):
    """Returns (NodeTable, CsrShard) for a dense-id homogeneous graph."""
    W, r, dev = rt.world, rt.rank, rt.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 9973 + r)
    n_local = (num_nodes - r + W - 1) // W
    local_ids = torch.arange(n_local, device=dev, dtype=torch.int64) * W + r
    idmap = IdMap(rt, local_ids, dense=True)
    nodes = NodeTable(rt, ntype, idmap)
    # features: class-dependent mean + noise so that the task is learnable
    labels = torch.randint(0, num_classes, (n_local,), device=dev, generator=g)
    centers = torch.randn(num_classes, feat_dim, device=dev, generator=torch.Generator(device=dev).manual_seed(seed + 17))
    x = centers[labels] * 0.5 + torch.randn(n_local, feat_dim, device=dev, generator=g)
    nodes.set_float(x, feature_dtype)
    del x
    nodes.set_labels(labels)
    # edges: skewed out-degree, uniformly random destinations
    e_local = num_edges // W + (1 if r < num_edges % W else 0)
    wdeg = torch.exp(torch.randn(n_local, device=dev, generator=g, dtype=torch.float64) * skew)
    deg = torch.floor(wdeg / wdeg.sum() * e_local).to(torch.int64)
    rem = int(e_local - int(deg.sum()))
    if rem > 0:      # hand the rounding remainder to random rows
        extra = torch.randint(0, max(n_local, 1), (rem,), device=dev, generator=g)
        deg += torch.bincount(extra, minlength=n_local)
    src_rows = torch.repeat_interleave(torch.arange(n_local, device=dev, dtype=torch.int64), deg)
    del wdeg, deg
    dst = torch.randint(0, num_nodes, (e_local,), device=dev, generator=g)
    weights = torch.rand(e_local, device=dev, generator=g) + 0.05 if weighted else None
    csr = CsrShard.from_coo(rt, etype, ntype, ntype, src_rows, dst, n_local, weights=weights)
    return nodes, csr


def make_partitioned_sources(rt: Runtime, num_nodes: int, num_edges: int, feat_dim: int, num_classes: int, seed: int = 0,
                             skew: float = 1.0):
    """This rank's share of the same synthetic graph as in-memory ``gl.Graph`` sources:
    ``(node_dict, edge_dict)`` for ``g.node(node_dict, ...)`` / ``g.edge(edge_dict, ...)`` (``partitioned=True``: the rows are
    already the ones this rank owns, so ``init()`` neither filters nor shuffles them).  Lets benchmarks and tests go
    through the public Graph / GSL API without writing gigabytes of TSV first."""
    W, r, dev = rt.world, rt.rank, rt.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 9973 + r)
    n_local = (num_nodes - r + W - 1) // W
    ids = torch.arange(n_local, device=dev, dtype=torch.int64) * W + r
    labels = torch.randint(0, num_classes, (n_local,), device=dev, generator=g)
    centers = torch.randn(num_classes, feat_dim, device=dev, generator=torch.Generator(device=dev).manual_seed(seed + 17))
    x = centers[labels] * 0.5 + torch.randn(n_local, feat_dim, device=dev, generator=g)
    e_local = num_edges // W + (1 if r < num_edges % W else 0)
    wdeg = torch.exp(torch.randn(n_local, device=dev, generator=g, dtype=torch.float64) * skew)
    deg = torch.floor(wdeg / wdeg.sum() * e_local).to(torch.int64)
    rem = int(e_local - int(deg.sum()))
    if rem > 0:
        deg += torch.bincount(torch.randint(0, max(n_local, 1), (rem,), device=dev, generator=g), minlength=n_local)
    src = torch.repeat_interleave(ids, deg)
    dst = torch.randint(0, num_nodes, (e_local,), device=dev, generator=g)
    node = {"ids": ids, "labels": labels, "float_attrs": x, "partitioned": True}
    edge = {"src_ids": src, "dst_ids": dst, "partitioned": True}
    return node, edge
