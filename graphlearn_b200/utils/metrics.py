"""Evaluation metrics of the reference's example suite (graphlearn/examples/eval/):

* ``eval_metrics(gt_ids, recall_ids)``  Recall / NDCG / HitRate of a recall list against ground-truth items, summed over the
  batch like eval_rec_metric.py:23-55 (divide by the number of triggers at the end);
* ``recall_metrics_at_k``               the same three metrics for a whole evaluation set in one vectorised call (device tensors);
* ``evaluate_recall``                   embeddings -> KNN recall (``Graph.search`` / ``ops.knn``) -> metrics, the test_rec.py flow;
* ``hits_at_k``                         OGB link-prediction Hits@K (link_trainer.py ``eval_hits``);
* ``multilabel_f1``                     micro / macro F1 of a one-vs-rest logistic regression on frozen embeddings with the
  "predict as many labels as the node has" protocol (blogcatelog_eval.py:41-104; needs scikit-learn).
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np
import torch


def eval_metrics(gt_ids: Sequence[Sequence[int]], recall_ids) -> Tuple[float, float, float]:
    """-> (total_recall, total_ndcg, total_hits) summed over the rows (eval_rec_metric.py)."""
    total_hits = total_recall = total_ndcg = 0.0
    for gt, rec in zip(gt_ids, recall_ids):
        gt = set(int(x) for x in gt)
        hit_pos = [i for i, r in enumerate(np.asarray(rec).reshape(-1).tolist()) if int(r) in gt]
        if hit_pos:
            dcg = sum(1.0 / np.log2(i + 2) for i in hit_pos)
            idcg = sum(1.0 / np.log2(i + 2) for i in range(len(hit_pos)))
            total_ndcg += dcg / idcg
            total_hits += 1
        total_recall += len(hit_pos) / max(len(gt), 1)
    return total_recall, total_ndcg, total_hits


def recall_metrics_at_k(recall_ids: torch.Tensor, gt_ids: torch.Tensor, gt_mask: torch.Tensor = None) -> Dict[str, float]:
    """recall_ids [B, k] (best first, -1 = none), gt_ids [B, G] padded with ``gt_mask`` False (or -1) -> averaged
    {"recall", "ndcg", "hit_rate"} with the reference's definitions."""
    rec = torch.as_tensor(recall_ids)
    gt = torch.as_tensor(gt_ids).to(rec.device)
    mask = (gt >= 0) if gt_mask is None else torch.as_tensor(gt_mask).to(rec.device)
    hit = ((rec[:, :, None] == gt[:, None, :]) & mask[:, None, :] & (rec[:, :, None] >= 0)).any(2)        # [B, k]
    n_hit = hit.sum(1)
    disc = 1.0 / torch.log2(torch.arange(rec.size(1), device=rec.device, dtype=torch.float64) + 2)
    dcg = (hit.to(torch.float64) * disc).sum(1)
    idcg = torch.cumsum(disc, 0)[(n_hit - 1).clamp(min=0)]
    ndcg = torch.where(n_hit > 0, dcg / idcg, torch.zeros_like(dcg))
    recall = n_hit.to(torch.float64) / mask.sum(1).clamp(min=1)
    return {"recall": float(recall.mean()), "ndcg": float(ndcg.mean()), "hit_rate": float((n_hit > 0).double().mean())}


def evaluate_recall(graph, item_type: str, query_vectors, gt_items: Sequence[Sequence[int]], top_k: int = 20) -> Dict[str, float]:
    """KNN recall of ``query_vectors`` over the float attributes (= embeddings) of ``item_type`` (examples/eval/test_rec.py:
    u2i - queries are user embeddings, i2i - item embeddings; the metric comes from ``gl.set_knn_metric``)."""
    from ..ops.knn import KnnOption
    ids, _ = graph.search(item_type, query_vectors, KnnOption(k=top_k))
    r, n, h = eval_metrics(gt_items, ids)
    m = max(len(gt_items), 1)
    return {"recall": r / m, "ndcg": n / m, "hit_rate": h / m, "k": top_k}


def hits_at_k(y_pred_pos, y_pred_neg, k: int) -> float:
    """OGB-style Hits@K (examples/tf/link_trainer.py:162-173): the share of positive scores that beat the k-th best negative
    score; 1.0 when there are fewer than k negatives or no positives."""
    pos = np.asarray(y_pred_pos, dtype=np.float64).reshape(-1)
    neg = np.asarray(y_pred_neg, dtype=np.float64).reshape(-1)
    if neg.size < k or pos.size == 0:
        return 1.0
    kth = np.sort(neg)[-k]
    return float((pos > kth).sum()) / pos.size


def multilabel_f1(embeddings: np.ndarray, labels: np.ndarray, train_ratios: Sequence[float] = (0.5, 0.9), shuffles: int = 2,
                  seed: int = 0, max_iter: int = 200) -> Dict[float, Dict[str, float]]:
    """embeddings [N, d], labels [N, C] multi-hot -> {train_ratio: {"micro": f1, "macro": f1}} averaged over ``shuffles``
    random splits; a test node with m labels is assigned its m most probable classes."""
    from sklearn.linear_model import LogisticRegression
    from sklearn.metrics import f1_score
    from sklearn.multiclass import OneVsRestClassifier
    emb, lab = np.asarray(embeddings, dtype=np.float64), np.asarray(labels) > 0
    rs = np.random.RandomState(seed)
    out: Dict[float, Dict[str, float]] = {}
    for ratio in train_ratios:
        acc = {"micro": 0.0, "macro": 0.0}
        for _ in range(shuffles):
            perm = rs.permutation(emb.shape[0])
            n_tr = int(ratio * emb.shape[0])
            tr, te = perm[:n_tr], perm[n_tr:]
            keep = lab[tr].any(0) & ~lab[tr].all(0)                  # classes a classifier can be fitted for
            clf = OneVsRestClassifier(LogisticRegression(max_iter=max_iter))
            clf.fit(emb[tr], lab[tr][:, keep])
            prob = np.asarray(clf.predict_proba(emb[te]))
            full = np.zeros((te.size, lab.shape[1]))
            full[:, keep] = prob
            pred = np.zeros_like(lab[te])
            for i, m in enumerate(lab[te].sum(1)):
                if m > 0:
                    pred[i, np.argsort(full[i])[-int(m):]] = True
            for avg in acc:
                acc[avg] += f1_score(lab[te], pred, average=avg, zero_division=0)
        out[float(ratio)] = {a: v / shuffles for a, v in acc.items()}
    return out
