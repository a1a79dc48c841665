"""Losses of the reference (graphlearn/python/nn/tf/loss.py:28-109) in PyTorch."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def sigmoid_cross_entropy_loss(pos_logit, neg_logit):
    """mean BCE of positive logits vs 1 and negative logits vs 0 (loss.py:28-40)."""
    pos = F.binary_cross_entropy_with_logits(pos_logit, torch.ones_like(pos_logit))
    neg = F.binary_cross_entropy_with_logits(neg_logit, torch.zeros_like(neg_logit))
    return pos + neg


def unsupervised_softmax_cross_entropy_loss(src_emb, pos_emb, neg_emb, temperature=1.0):
    """sampled softmax: positive logit against `k` sampled negative logits (loss.py:43-66).
    src [B, d], pos [B, d], neg [B, k, d] or [B*k, d]."""
    B, d = src_emb.shape
    neg = neg_emb.view(B, -1, d)
    pos_logit = (src_emb * pos_emb).sum(-1, keepdim=True)
    neg_logit = torch.einsum("bd,bkd->bk", src_emb, neg)
    logits = torch.cat([pos_logit, neg_logit], 1) / temperature
    return F.cross_entropy(logits, torch.zeros(B, dtype=torch.long, device=logits.device))


def triplet_margin_loss(pos_src_emb, pos_edge_emb, pos_dst_emb, neg_src_emb, neg_edge_emb, neg_dst_emb, margin=1.0,
                        neg_num=1, L=2):
    """TransE (loss.py:69-92): max(0, margin + d(h + r, t) - d(h' + r', t'))."""
    pos_d = torch.norm(pos_src_emb + pos_edge_emb - pos_dst_emb, p=L, dim=-1)
    neg_d = torch.norm(neg_src_emb + neg_edge_emb - neg_dst_emb, p=L, dim=-1)
    pos_d = pos_d.repeat_interleave(neg_num) if neg_d.numel() == pos_d.numel() * neg_num else pos_d
    return F.relu(margin + pos_d - neg_d).mean()


def triplet_softplus_loss(pos_src_emb, pos_edge_emb, pos_dst_emb, neg_src_emb, neg_edge_emb, neg_dst_emb):
    """DistMult (loss.py:95-109): softplus(-score(pos)) + softplus(score(neg))."""
    pos = (pos_src_emb * pos_edge_emb * pos_dst_emb).sum(-1)
    neg = (neg_src_emb * neg_edge_emb * neg_dst_emb).sum(-1)
    return F.softplus(-pos).mean() + F.softplus(neg).mean()
