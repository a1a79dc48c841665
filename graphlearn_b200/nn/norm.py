"""GraphSAINT-style sampling normalisation (graphlearn/python/nn/tf/utils/compute_norm.py:23-73):
node / edge appearance frequencies over many sampled subgraphs turned into loss (`node_norm`) and aggregator
(`edge_norm`) weights."""
from __future__ import annotations

import torch

from .. import errors


def _node_edge_ids(sg):
    """(node ids, edge ids | None, row node id of every edge | None) of a SubGraph result or a (nodes, edges) pair."""
    if isinstance(sg, (tuple, list)):
        nodes, edges = sg
        return torch.as_tensor(nodes).reshape(-1).cpu(), (None if edges is None else torch.as_tensor(edges).reshape(-1).cpu()), None
    nodes = sg.nodes.ids_t.reshape(-1).cpu()
    if sg.edges is None:
        return nodes, None, None
    eids = sg.edges.tensor("edge_ids").reshape(-1).cpu()
    rows = nodes[sg.edge_index_t[0].cpu()]
    return nodes, eids, rows


def compute_saint_norm(subgraphs, num_nodes: int, num_edges: int = 0):
    """subgraphs: iterable of ``SubGraph`` results or (node_ids, edge_ids | None) pairs.
    Returns (node_norm [num_nodes], edge_norm [num_edges] | None):
      node_norm[v] = N_subgraphs / C_v / num_nodes          (loss weight)
      edge_norm[e] = C_row(e) / C_e, clipped to [0, 1e4], 0.1 where the edge never appeared   (aggregator weight)"""
    node_cnt = torch.zeros(num_nodes)
    edge_cnt = torch.zeros(max(num_edges, 1))
    edge_row = torch.zeros(max(num_edges, 1), dtype=torch.long)
    n_sub = 0
    for sg in subgraphs:
        nodes, eids, rows = _node_edge_ids(sg)
        node_cnt[torch.unique(nodes[nodes >= 0])] += 1
        if eids is not None and num_edges:
            ok = eids >= 0
            u = torch.unique(eids[ok])
            edge_cnt[u] += 1
            if rows is not None:
                edge_row[eids[ok]] = rows[ok]
        n_sub += 1
    node_norm = n_sub / node_cnt.clamp(min=1) / max(num_nodes, 1)
    edge_norm = None
    if num_edges:
        ratio = node_cnt[edge_row.clamp(0, num_nodes - 1)] / edge_cnt
        edge_norm = torch.where(edge_cnt > 0, ratio.clamp(0, 1e4), torch.full_like(ratio, 0.1))
    return node_norm, edge_norm


def compute_norm(total_nodes: int, total_edges: int, subgraph_sampler, sample_coverage: int = 10):
    """Reference entry point: keep drawing from ``subgraph_sampler`` (``g.subgraph_sampler(...)``) until
    ``total_nodes * sample_coverage`` nodes have been sampled, then normalise."""
    def stream():
        seen = 0
        while seen < total_nodes * sample_coverage:
            try:
                sg = subgraph_sampler.get()
            except errors.OutOfRangeError:
                continue
            seen += int(sg.nodes.ids_t.numel())
            yield sg
    return compute_saint_norm(stream(), total_nodes, total_edges)
