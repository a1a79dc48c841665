"""Generic EgoGNN (graphlearn/python/nn/tf/model/ego_gnn.py:58-110): K EgoLayers over a
K-hop ego graph; layer i applies its convs to every adjacent hop pair, optional BN /
activation / dropout between layers.  Works with EgoSAGEConv, EgoGATConv, EgoGINConv."""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..nn.conv import EgoGATConv, EgoGINConv, EgoLayer, EgoSAGEConv


class EgoGNN(nn.Module):
    def __init__(self, layers: Sequence[EgoLayer], bn_func: Callable = None, act_func: Callable = F.relu,
                 dropout: float = 0.0):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        self.bn = nn.ModuleList()
        self.act, self.dropout = act_func, dropout
        self._bn_func = bn_func

    def forward(self, x_list: List[torch.Tensor], expands: Sequence[int]) -> torch.Tensor:
        """x_list[i]: features of hop i ([B * prod(expands[:i]), d_i])."""
        h = list(x_list)
        L = len(self.layers)
        for l, layer in enumerate(self.layers):
            h = layer(h, expands[:len(h) - 1])
            if l < L - 1:
                h = [self.act(t.float()) if self.act else t for t in h]
                if self.dropout and self.training:
                    h = [F.dropout(t, self.dropout) for t in h]
        return h[0].float()


def make_ego_gnn(kind: str, in_dim: int, hidden_dim: int, out_dim: int, num_layers: int, **kw) -> EgoGNN:
    dims = [in_dim] + [hidden_dim] * (num_layers - 1) + [out_dim]
    layers = []
    for l in range(num_layers):
        convs = []
        for _ in range(num_layers - l):
            if kind == "sage":
                convs.append(EgoSAGEConv(dims[l], dims[l + 1], kw.get("agg_type", "mean")))
            elif kind == "gat":
                convs.append(EgoGATConv(dims[l], dims[l + 1], kw.get("num_head", 4), attn_drop=kw.get("attn_drop", 0.0)))
            elif kind == "gin":
                convs.append(EgoGINConv(dims[l], dims[l + 1], eps=kw.get("eps", 0.0)))
            else:
                raise ValueError(kind)
        # the reference shares ONE conv per layer across hop pairs
        shared = convs[0]
        layers.append(EgoLayer([shared] * (num_layers - l)))
    return EgoGNN(layers, dropout=kw.get("dropout", 0.0))
