"""K10: brute-force KNN over a node type's float attributes.

The reference wraps faiss indexes (flat / ivfflat / ivfpq, CPU or GPU) behind
``KnnOperator``: the query is broadcast to every server and the per-server
top-k lists are merged with a k-heap (graphlearn/src/contrib/knn/knn_request.cc:96-111,
167-202).  On B200 a flat index is one bf16/fp32 GEMM (queries x shard^T, tensor
cores via cuBLAS - a plain library GEMM) + ``topk`` per shard, followed by an
all-gather and a final top-k merge.  Metric 0 = L2, 1 = inner product
(``gl.set_knn_metric``).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class KnnOption(object):
    def __init__(self, k=1):
        self.k = int(k)


def _shard_topk(x: torch.Tensor, q: torch.Tensor, k: int, metric: int, chunk: int = 1 << 18):
    n = x.size(0)
    best_s = torch.full((q.size(0), k), float("-inf"), device=q.device)
    best_i = torch.full((q.size(0), k), -1, dtype=torch.int64, device=q.device)
    qn = (q * q).sum(1, keepdim=True)
    for s in range(0, n, chunk):
        xs = x[s:s + chunk].to(q.dtype)
        ip = q @ xs.t()
        score = ip if metric == 1 else -(qn - 2 * ip + (xs * xs).sum(1)[None, :])
        kk = min(k, score.size(1))
        v, i = torch.topk(score, kk, dim=1)
        cat_s = torch.cat([best_s, v], 1)
        cat_i = torch.cat([best_i, i + s], 1)
        v2, j = torch.topk(cat_s, k, dim=1)
        best_s, best_i = v2, torch.gather(cat_i, 1, j)
    return best_s, best_i


def search(rt, table, queries: torch.Tensor, k: int, metric: int = 0):
    """-> (ids [B, k], distances [B, k]); distances are squared L2 (metric 0) or inner products (1)."""
    W = rt.world
    q = queries.to(rt.device).float()
    x = table.feats.local[:, :table.float_dim]
    s, rows = _shard_topk(x, q, k, metric)
    vids = torch.where(rows >= 0, rows * W + rt.rank, rows)
    if W > 1:
        all_s = [torch.empty_like(s) for _ in range(W)]
        all_v = [torch.empty_like(vids) for _ in range(W)]
        dist.all_gather(all_s, s.contiguous())
        dist.all_gather(all_v, vids.contiguous())
        s, vids = torch.cat(all_s, 1), torch.cat(all_v, 1)
        s, j = torch.topk(s, k, dim=1)
        vids = torch.gather(vids, 1, j)
    ids = table.idmap.to_id(vids)
    d = s if metric == 1 else -s
    return ids, d


class KnnOperator(object):
    """Imperative form of ``Graph.search`` (graphlearn/python/operator/knn_operator.py): bound to a node type,
    ``search(inputs, k)`` returns (ids [B, k], distances [B, k]) as numpy arrays."""

    def __init__(self, graph, node_type):
        self._g, self._type = graph, node_type

    def search(self, inputs, k=1):
        return self._g.search(self._type, inputs, KnnOption(k))
