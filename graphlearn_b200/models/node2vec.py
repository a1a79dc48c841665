"""DeepWalk / node2vec skip-gram model (graphlearn/examples/tf/node2vec/node2vec.py:53-112):
(center, context) pairs from walks with a (left, right) window, two embedding tables, sigmoid
cross entropy against k negatives.  The walks come from the resident-walker kernel (K3)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..nn.loss import sigmoid_cross_entropy_loss


def gen_pair(path: torch.Tensor, left_win: int, right_win: int):
    """path [B, L] -> (src [P], dst [P]) for every in-window pair (same order as the reference's
    utils.gen_pair: for each position, the left then right context)."""
    B, L = path.shape
    src, dst = [], []
    for c in range(L):
        for j in range(max(0, c - left_win), min(L, c + right_win + 1)):
            if j != c:
                src.append(path[:, c])
                dst.append(path[:, j])
    return torch.stack(src, 1).reshape(-1), torch.stack(dst, 1).reshape(-1)


class Node2Vec(nn.Module):
    def __init__(self, num_nodes: int, dim: int, sparse: bool = True):
        super().__init__()
        self.src_emb = nn.Embedding(num_nodes, dim, sparse=sparse)
        self.ctx_emb = nn.Embedding(num_nodes, dim, sparse=sparse)
        nn.init.uniform_(self.src_emb.weight, -0.5 / dim, 0.5 / dim)
        nn.init.zeros_(self.ctx_emb.weight)

    def forward(self, src, pos, neg):
        """src [P], pos [P], neg [P, k] -> loss."""
        s = self.src_emb(src)
        pos_logit = (s * self.ctx_emb(pos)).sum(-1)
        neg_logit = torch.einsum("pd,pkd->pk", s, self.ctx_emb(neg))
        return sigmoid_cross_entropy_loss(pos_logit, neg_logit.reshape(-1))

    def embeddings(self):
        return self.src_emb.weight.detach()
