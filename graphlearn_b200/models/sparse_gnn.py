"""SubGraph-based models (GCN / SAGE / GAT over edge_index batches; SEAL link prediction with
DRNL node labels) - graphlearn/examples/tf/{sage,seal,seal_v2} and examples/pytorch/gcn."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..nn.sparse_conv import GATConv, GCNConv, SAGEConv

_CONV = {"gcn": GCNConv, "sage": SAGEConv, "gat": GATConv}


def _nn_conf():
    from ..nn.utils import conf          # the reference's tfg.conf.training switch
    return conf


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
                if self.training and self.dropout and _nn_conf().training:
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


class _BatchGraphModel(nn.Module):
    """SubGraph-based link model of the reference (nn/tf/model/{gcn,sage,gat}.py): stacks sparse convs over a ``BatchGraph`` and
    returns the embeddings of the first two nodes of every subgraph (the (src, dst) pair the sub-graph was induced around)."""

    kind = "gcn"

    def __init__(self, batch_size, input_dim, hidden_dim, output_dim, depth=2, drop_rate=0.0, encoder=None, **kw):
        super().__init__()
        self.batch_size, self.depth, self.encoder = batch_size, depth, encoder
        self.gnn = SparseGNN(self.kind, input_dim, hidden_dim, output_dim, num_layers=depth, dropout=drop_rate, **kw)

    def forward(self, batchgraph):
        bg = batchgraph.transform(self.encoder)
        h = self.gnn(bg.nodes, bg.edge_index)
        off = bg.graph_node_offsets[:-1] if bg.graph_node_offsets.numel() > bg.num_graphs else bg.graph_node_offsets
        return h[off], h[off + 1]


class GCN(_BatchGraphModel):
    kind = "gcn"


class GraphSAGE(_BatchGraphModel):
    """``agg_type`` in mean | sum (sage.py:29-60)"""
    kind = "sage"

    def __init__(self, batch_size, input_dim, hidden_dim, output_dim, depth=2, drop_rate=0.0, agg_type="mean", **kw):
        super().__init__(batch_size, input_dim, hidden_dim, output_dim, depth, drop_rate, agg_type=agg_type, **kw)


class GAT(_BatchGraphModel):
    """``attn_heads`` on the hidden layers, one head on the output layer (gat.py:29-66)."""
    kind = "gat"

    def __init__(self, batch_size, input_dim, hidden_dim, output_dim, depth=2, drop_rate=0.0, attn_heads=1, attn_drop=0.0, **kw):
        nn.Module.__init__(self)
        self.batch_size, self.depth, self.encoder = batch_size, depth, kw.pop("encoder", None)
        dims = [input_dim] + [hidden_dim] * (depth - 1) + [output_dim]
        convs = []
        for i in range(depth):
            last = i == depth - 1 and depth != 1
            convs.append(GATConv(dims[i], dims[i + 1], num_heads=1 if last else attn_heads, concat=False, attn_drop=attn_drop))
        self.gnn = SparseGNN.__new__(SparseGNN)
        nn.Module.__init__(self.gnn)
        self.gnn.convs, self.gnn.dropout = nn.ModuleList(convs), drop_rate
