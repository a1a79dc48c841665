"""GraphSAINT-style sampling normalisation (graphlearn/python/nn/tf/utils/compute_norm.py:23-73):
node / edge appearance frequencies over a number of sampled subgraphs turned into loss
(`node_norm`) and aggregator (`edge_norm`) weights."""
from __future__ import annotations

import torch


def compute_saint_norm(subgraphs, num_nodes: int, num_edges: int = 0):
    """subgraphs: iterable of (node_ids [n], edge_ids [m] | None).  Returns (node_norm, edge_norm)."""
    node_cnt = torch.zeros(num_nodes)
    edge_cnt = torch.zeros(max(num_edges, 1))
    n_sub = 0
    for nodes, edges in subgraphs:
        node_cnt[torch.unique(nodes.cpu())] += 1
        if edges is not None and num_edges:
            edge_cnt[torch.unique(edges.cpu())] += 1
        n_sub += 1
    node_norm = n_sub / node_cnt.clamp(min=1) / max(num_nodes, 1)
    edge_norm = (edge_cnt / node_cnt.mean().clamp(min=1)).clamp(min=1e-6) if num_edges else None
    return node_norm, edge_norm
