"""EgoGraphSAGE: K layers over a K-hop fixed fan-out ego graph.

Layer i applies the same conv to every adjacent hop pair; K layers consume K
hops (graphlearn/python/nn/tf/model/ego_gnn.py:58-110, ego_layer.py:54-91).
The first layer reads raw features from the sharded store inside the fused
kernel; deeper layers run on dense bf16 activations.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from ..nn.conv import EgoSAGEConv


class EgoGraphSAGE(nn.Module):
    def __init__(self, in_dim: int, hidden_dim: int, out_dim: int, num_layers: int = 2, agg_type: str = "mean",
                 dropout: float = 0.0, bf16_activations: bool = True):
        super().__init__()
        dims = [in_dim] + [hidden_dim] * (num_layers - 1) + [out_dim]
        self.convs = nn.ModuleList([EgoSAGEConv(dims[i], dims[i + 1], agg_type) for i in range(num_layers)])
        self.num_layers = num_layers
        self.dropout = dropout
        self.bf16 = bf16_activations

    def forward_store(self, table, hops: Sequence[torch.Tensor], fanouts: Sequence[int]) -> torch.Tensor:
        """hops[0] = seed vids [B]; hops[i] = vids [B*k1*..*k_i]; returns logits [B, out]."""
        L = self.num_layers
        assert len(hops) == L + 1 and len(fanouts) == L
        h: List[torch.Tensor] = []
        last = L == 1
        for i in range(L):      # layer 0 straight from the store
            h.append(self.convs[0].forward_store(table, hops[i], hops[i + 1], fanouts[i], relu=not last,
                                                 out_bf16=self.bf16 and not last))
        for l in range(1, L):
            last = l == L - 1
            nh = []
            for i in range(L - l):
                x = h[i]
                if self.training and self.dropout > 0:
                    x = torch.nn.functional.dropout(x, self.dropout)
                nh.append(self.convs[l](x, h[i + 1], fanouts[i], relu=not last, out_bf16=self.bf16 and not last))
            h = nh
        return h[0].float()

    def forward(self, xs: Sequence[torch.Tensor], fanouts: Sequence[int]) -> torch.Tensor:
        """Dense variant: xs[i] = features of hop i ([B*prod(k), d])."""
        L = self.num_layers
        h = list(xs)
        for l in range(L):
            last = l == L - 1
            h = [self.convs[l](h[i], h[i + 1], fanouts[i], relu=not last, out_bf16=self.bf16 and not last)
                 for i in range(L - l)]
        return h[0].float()
