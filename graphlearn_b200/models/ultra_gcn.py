"""UltraGCN recommender (graphlearn/examples/tf/ultra_gcn/ultra_gcn.py:30-118): user/item embedding
tables trained with a degree-weighted u2i sigmoid cross entropy, an i2i constraint loss over each
positive item's top-k similar items, and L2 regularisation.  No message passing at all - the
"infinite-layer" GCN is approximated by the loss weights 1 + 1/sqrt(d_u d_i)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class UltraGCN(nn.Module):
    def __init__(self, user_num: int, item_num: int, emb_dim: int, neg_weight: float = 1.0, i2i_weight: float = 1.0,
                 l2_weight: float = 1e-4, need_hash: bool = False):
        super().__init__()
        self.user_num, self.item_num, self.need_hash = int(user_num), int(item_num), need_hash
        self.neg_weight, self.i2i_weight, self.l2_weight = neg_weight, i2i_weight, l2_weight
        self.user_emb = nn.Embedding(user_num, emb_dim)
        self.item_emb = nn.Embedding(item_num, emb_dim)
        nn.init.normal_(self.user_emb.weight, std=0.05)
        nn.init.normal_(self.item_emb.weight, std=0.05)

    def _u(self, ids):
        return self.user_emb(ids % self.user_num if self.need_hash else ids)

    def _i(self, ids):
        return self.item_emb(ids % self.item_num if self.need_hash else ids)

    def forward(self, user_ids, user_deg, item_ids, item_deg, nbr_ids, nbr_weights, neg_ids):
        """user/item ids [B], degrees [B] (float), nbr_ids/weights [B, k] (top-k i2i neighbours of the
        positive item), neg_ids [B, n] -> scalar loss."""
        u, i = self._u(user_ids), self._i(item_ids)
        nbr, neg = self._i(nbr_ids.clamp(min=0)), self._i(neg_ids.clamp(min=0))
        pos_logit = (u * i).sum(-1)
        true_x = F.binary_cross_entropy_with_logits(pos_logit, torch.ones_like(pos_logit), reduction="none")
        neg_logit = (u.unsqueeze(1) * neg).sum(-1)
        neg_x = F.binary_cross_entropy_with_logits(neg_logit, torch.zeros_like(neg_logit), reduction="none")
        beta = 1.0 + 1.0 / torch.sqrt((user_deg.float() * item_deg.float()).clamp(min=1.0))
        loss_u2i = (true_x * beta).sum() + self.neg_weight * neg_x.mean(-1).sum()
        nbr_logit = (u.unsqueeze(1) * nbr).sum(-1)
        nbr_x = F.binary_cross_entropy_with_logits(nbr_logit, torch.ones_like(nbr_logit), reduction="none")
        valid = (nbr_ids >= 0).float()
        loss_i2i = (nbr_x * (1.0 + nbr_weights.float()) * valid).sum()
        l2 = 0.5 * (u.pow(2).sum() + i.pow(2).sum() + nbr.pow(2).sum() + neg.pow(2).sum())
        return loss_u2i + self.i2i_weight * loss_i2i + self.l2_weight * l2

    def user_embeddings(self, ids):
        return self._u(ids).detach()

    def item_embeddings(self, ids):
        return self._i(ids).detach()
