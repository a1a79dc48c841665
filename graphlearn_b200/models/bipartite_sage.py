"""Bipartite GraphSAGE for u2i recommendation
(graphlearn/examples/tf/ego_bipartite_sage, bipartite_sage): separate user / item towers over
heterogeneous ego graphs (u -> i -> u ..., i -> u -> i ...), dot-product score, sampled
softmax / sigmoid loss with in-batch or sampled negatives."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..nn.conv import EgoGATConv, EgoSAGEConv
from ..nn.loss import sigmoid_cross_entropy_loss, unsupervised_softmax_cross_entropy_loss


class _Tower(nn.Module):
    def __init__(self, dims_by_hop: Sequence[int], hidden: int, out: int, agg="mean", conv="sage", num_head=4):
        super().__init__()
        self.conv_kind = conv
        L = len(dims_by_hop) - 1
        self.L = L
        convs = []
        for l in range(L):
            row = nn.ModuleList()
            for i in range(L - l):
                d_self = dims_by_hop[i] if l == 0 else hidden
                d_nbr = dims_by_hop[i + 1] if l == 0 else hidden
                o = out if l == L - 1 else hidden
                row.append(EgoSAGEConv((d_self, d_nbr), o, agg) if conv == "sage" else
                           EgoGATConv((d_self, d_nbr), o, num_head))
            convs.append(row)
        self.convs = nn.ModuleList(convs)

    def forward(self, xs, expands):
        h = list(xs)
        for l in range(self.L):
            last = l == self.L - 1
            if self.conv_kind == "sage":
                h = [self.convs[l][i](h[i], h[i + 1], expands[i], relu=not last) for i in range(self.L - l)]
            else:
                h = [self.convs[l][i](h[i], h[i + 1], expands[i]) for i in range(self.L - l)]
                if not last:
                    h = [F.elu(t) for t in h]
        return h[0].float()

    def forward_store(self, tables, vids, expands):
        """Same recursion, but layer 0 reads the raw rows of hop i / i+1 from ``tables[i]`` / ``tables[i + 1]`` by vid
        INSIDE the fused kernels (``vids[i]``: device int64 ids of hop i) - no [M * k, d] feature tensors."""
        L = self.L
        last = L == 1
        if self.conv_kind == "sage":
            h = [self.convs[0][i].forward_store(tables[i], vids[i].reshape(-1), vids[i + 1].reshape(-1), expands[i], relu=not last,
                                                nbr_table=tables[i + 1]) for i in range(L)]
        else:
            h = [self.convs[0][i].forward_store(tables[i], vids[i].reshape(-1), vids[i + 1].reshape(-1), expands[i],
                                                nbr_table=tables[i + 1]) for i in range(L)]
            if not last:
                h = [F.elu(t) for t in h]
        for l in range(1, L):
            last = l == L - 1
            if self.conv_kind == "sage":
                h = [self.convs[l][i](h[i], h[i + 1], expands[i], relu=not last) for i in range(L - l)]
            else:
                h = [self.convs[l][i](h[i], h[i + 1], expands[i]) for i in range(L - l)]
                if not last:
                    h = [F.elu(t) for t in h]
        return h[0].float()


class EgoBipartiteSAGE(nn.Module):
    def __init__(self, user_dim: int, item_dim: int, hidden: int, out: int, hops: int = 2, agg="mean", conv="sage",
                 num_head=4):
        """conv = "sage" (EgoSAGEConv) or "gat" (EgoGATConv with `num_head` averaged heads - the 2-layer 4-head
        bipartite GAT of the Taobao-shaped benchmark config)."""
        super().__init__()
        u_dims = [user_dim if i % 2 == 0 else item_dim for i in range(hops + 1)]
        i_dims = [item_dim if i % 2 == 0 else user_dim for i in range(hops + 1)]
        self.user_tower = _Tower(u_dims, hidden, out, agg, conv, num_head)
        self.item_tower = _Tower(i_dims, hidden, out, agg, conv, num_head)

    def forward(self, user_ego, item_ego, expands_u, expands_i):
        return self.user_tower(user_ego, expands_u), self.item_tower(item_ego, expands_i)

    def forward_store(self, user_tables, user_vids, item_tables, item_vids, expands_u, expands_i):
        """Towers fed by (table, vid) pairs: the first layer gathers inside the fused kernels."""
        return (self.user_tower.forward_store(user_tables, user_vids, expands_u),
                self.item_tower.forward_store(item_tables, item_vids, expands_i))

    @staticmethod
    def loss(u_emb, pos_emb, neg_emb, kind="softmax", temperature=1.0):
        if kind == "softmax":
            return unsupervised_softmax_cross_entropy_loss(u_emb, pos_emb, neg_emb, temperature)
        pos = (u_emb * pos_emb).sum(-1)
        neg = torch.einsum("bd,bkd->bk", u_emb, neg_emb.view(u_emb.size(0), -1, u_emb.size(1))).reshape(-1)
        return sigmoid_cross_entropy_loss(pos, neg)

    @staticmethod
    def in_batch_negative_loss(u_emb, i_emb, temperature=1.0):
        """in-batch negatives: every other item of the batch is a negative."""
        logits = u_emb @ i_emb.t() / temperature
        return F.cross_entropy(logits, torch.arange(u_emb.size(0), device=u_emb.device))
