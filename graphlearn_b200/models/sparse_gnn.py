"""SubGraph-based models (GCN / SAGE / GAT over edge_index batches; SEAL link prediction with
DRNL node labels) - graphlearn/examples/tf/{sage,seal,seal_v2} and examples/pytorch/gcn."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..nn.sparse_conv import GATConv, GCNConv, SAGEConv

_CONV = {"gcn": GCNConv, "sage": SAGEConv, "gat": GATConv}


class SparseGNN(nn.Module):
    def __init__(self, kind, in_dim, hidden, out_dim, num_layers=2, dropout=0.0, **kw):
        super().__init__()
        dims = [in_dim] + [hidden] * (num_layers - 1) + [out_dim]
        self.convs = nn.ModuleList([_CONV[kind](dims[i], dims[i + 1], **kw) for i in range(num_layers)])
        self.dropout = dropout

    def forward(self, x, edge_index):
        for i, conv in enumerate(self.convs):
            x = conv(x, edge_index)
            if i < len(self.convs) - 1:
                x = F.relu(x)
                if self.training and self.dropout:
                    x = F.dropout(x, self.dropout)
        return x


def drnl_node_labeling(dist_to_src: torch.Tensor, dist_to_dst: torch.Tensor, max_label: int = 1000) -> torch.Tensor:
    """Double-radius node labelling of SEAL from the two BFS distance vectors."""
    ds, dd = dist_to_src.clone(), dist_to_dst.clone()
    INF = 2 ** 31 - 1
    unreachable = (ds >= INF) | (dd >= INF)
    d = ds + dd
    half, mod = torch.div(d, 2, rounding_mode="floor"), d % 2
    z = 1 + torch.minimum(ds, dd) + half * (half + mod - 1)
    z[0], z[1] = 1, 1
    z[unreachable] = 0
    return z.clamp(max=max_label)


class SEAL(nn.Module):
    def __init__(self, feat_dim, hidden, num_layers=3, max_label=1000, kind="gcn"):
        super().__init__()
        self.label_emb = nn.Embedding(max_label + 1, hidden)
        self.gnn = SparseGNN(kind, feat_dim + hidden, hidden, hidden, num_layers)
        self.mlp = nn.Sequential(nn.Linear(2 * hidden, hidden), nn.ReLU(), nn.Linear(hidden, 1))

    def forward(self, x, edge_index, z):
        h = torch.cat([x.float(), self.label_emb(z)], 1) if x is not None and x.numel() else self.label_emb(z)
        h = self.gnn(h, edge_index)
        return self.mlp(torch.cat([h[0], h[1]], -1)).squeeze(-1)     # src is node 0, dst is node 1
