"""Sparse (edge_index) GNN layers for SubGraph / BatchGraph inputs - PyTorch ports of
graphlearn/python/nn/tf/layers/{gcn,sage,gat}_conv.py and utils/softmax.py."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops.sparse import spmm


def segment_softmax(score: torch.Tensor, index: torch.Tensor, num_segments: int) -> torch.Tensor:
    """softmax of `score` [E, ...] within groups given by `index` [E] (utils/softmax.py:24-50)."""
    shape = (num_segments,) + tuple(score.shape[1:])
    idx = index.view(-1, *([1] * (score.dim() - 1))).expand_as(score)
    mx = torch.full(shape, float("-inf"), device=score.device, dtype=score.dtype).scatter_reduce(0, idx, score, "amax")
    e = torch.exp(score - mx.gather(0, idx))
    den = torch.zeros(shape, device=score.device, dtype=score.dtype).scatter_add(0, idx, e)
    return e / den.gather(0, idx).clamp(min=1e-16)


class GCNConv(nn.Module):
    """add self loops, D^-1/2 A D^-1/2 X W (gcn_conv.py:48-77).  edge_index[0] = target row,
    edge_index[1] = source (messages flow col -> row, like the reference's segment_sum by row)."""

    def __init__(self, in_dim, out_dim, bias=True, normalize=True):
        super().__init__()
        self.lin = nn.Linear(in_dim, out_dim, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_dim)) if bias else None
        self.normalize = normalize

    def forward(self, x, edge_index, num_nodes=None):
        n = x.size(0) if num_nodes is None else num_nodes
        loop = torch.arange(n, device=x.device)
        row = torch.cat([edge_index[0], loop])
        col = torch.cat([edge_index[1], loop])
        h = self.lin(x.float())
        w = torch.ones(row.numel(), device=x.device)
        if self.normalize:
            deg = torch.zeros(n, device=x.device).scatter_add(0, row, w)
            dinv = deg.clamp(min=1).pow(-0.5)
            w = dinv[row] * dinv[col]
        out = spmm(h, row, col, w, n)                  # fused gather x norm -> scatter-add (K6)
        return out + self.bias if self.bias is not None else out


class SAGEConv(nn.Module):
    """W_self x + W_nbr agg_{j in N(i)} x_j, agg in sum|mean, optional L2 norm (sage_conv.py:61-92)."""

    def __init__(self, in_dim, out_dim, agg_type="mean", bias=True, normalize=False):
        super().__init__()
        self.lin_self = nn.Linear(in_dim, out_dim, bias=bias)
        self.lin_nbr = nn.Linear(in_dim, out_dim, bias=False)
        self.agg_type, self.normalize = agg_type, normalize

    def forward(self, x, edge_index, num_nodes=None):
        n = x.size(0) if num_nodes is None else num_nodes
        row, col = edge_index[0], edge_index[1]
        xf = x.float()
        agg = spmm(xf, row, col, None, n)
        if self.agg_type == "mean":
            deg = torch.zeros(n, device=x.device).scatter_add(0, row, torch.ones(row.numel(), device=x.device))
            agg = agg / deg.clamp(min=1)[:, None]
        out = self.lin_self(xf[:n]) + self.lin_nbr(agg)
        return F.normalize(out, dim=-1) if self.normalize else out


class GATConv(nn.Module):
    """h = W x; e_ij = leaky_relu(a_src.h_i + a_dst.h_j); segment softmax over incoming edges;
    heads concatenated or averaged (gat_conv.py:69-119)."""

    def __init__(self, in_dim, out_dim, num_heads=1, concat=True, bias=True, attn_drop=0.0, negative_slope=0.2):
        super().__init__()
        self.H, self.D, self.concat = num_heads, out_dim, concat
        self.lin = nn.Linear(in_dim, num_heads * out_dim, bias=False)
        self.att_i = nn.Parameter(torch.empty(num_heads, out_dim))
        self.att_j = nn.Parameter(torch.empty(num_heads, out_dim))
        self.bias = nn.Parameter(torch.zeros(num_heads * out_dim if concat else out_dim)) if bias else None
        self.attn_drop, self.slope = attn_drop, negative_slope
        nn.init.xavier_uniform_(self.att_i)
        nn.init.xavier_uniform_(self.att_j)

    def forward(self, x, edge_index, num_nodes=None):
        n = x.size(0) if num_nodes is None else num_nodes
        loop = torch.arange(n, device=x.device)
        row = torch.cat([edge_index[0], loop])
        col = torch.cat([edge_index[1], loop])
        h = self.lin(x.float()).view(-1, self.H, self.D)
        e = F.leaky_relu((h[row] * self.att_i).sum(-1) + (h[col] * self.att_j).sum(-1), self.slope)   # [E, H]
        a = segment_softmax(e, row, n)
        if self.training and self.attn_drop > 0:
            a = F.dropout(a, self.attn_drop)
        out = spmm(h.reshape(-1, self.H * self.D), row, col, a, n, heads=self.H).view(n, self.H, self.D)
        out = out.reshape(n, self.H * self.D) if self.concat else out.mean(1)
        return out + self.bias if self.bias is not None else out
