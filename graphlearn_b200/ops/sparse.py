"""Edge-list sparse ops for the SubGraph / BatchGraph models (K6, sparse flavour) and the hash-table
relabel primitive (K4).

``spmm(x, row, col, w, n_out, heads)``: out[row[e], h] += w[e, h] * x[col[e], h]  - the aggregation of
GCNConv (w = D^-1/2 A D^-1/2 entries), SAGEConv (w = None) and GATConv (w = attention) in ONE kernel
(csrc/graph_ops.cu edge_scatter_kernel: gather x weight -> float4 ``red.global.add`` atomics) instead of
gather -> multiply -> index_add.  Differentiable: d/dx is the same kernel with row/col swapped, d/dw is
the per-edge dot product kernel.  CPU tensors use the equivalent torch ops (oracle for the tests).
Reference ops replaced: graphlearn/python/nn/tf/layers/{gcn,sage,gat}_conv.py (gather + segment_sum / softmax) and
the host-side id->index maps of nn/pytorch/data/pyg_dataloader.py:56-64."""
from __future__ import annotations

from typing import Optional

import torch


def _native_for(t: torch.Tensor):
    if not t.is_cuda:
        return None
    from ..parallel.runtime import native
    return native()


def _spmm_torch(x, row, col, w, n_out, H):
    F_ = x.size(1)
    D = F_ // H
    msg = x[col].view(-1, H, D)
    if w is not None:
        msg = msg * w.view(-1, H, 1)
    return torch.zeros(n_out, H, D, device=x.device, dtype=x.dtype).index_add_(0, row, msg).reshape(n_out, F_)


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, row, col, w, n_out, H):
        C = _native_for(x)
        ctx.meta = (int(n_out), int(H), int(x.size(0)))
        ctx.save_for_backward(x, row, col, w if w is not None else torch.zeros(0, device=x.device))
        ctx.has_w = w is not None
        return C.edge_scatter(x, row, col, w, int(H), int(n_out))

    @staticmethod
    def backward(ctx, g):
        x, row, col, w = ctx.saved_tensors
        n_out, H, n_src = ctx.meta
        C = _native_for(g)
        g = g.contiguous()
        w_ = w if ctx.has_w else None
        gx = C.edge_scatter(g, col, row, w_, H, n_src) if ctx.needs_input_grad[0] else None
        gw = C.edge_dot(g, x, row, col, H).view_as(w) if (ctx.has_w and ctx.needs_input_grad[3]) else None
        return gx, None, None, gw, None, None


def spmm(x: torch.Tensor, row: torch.Tensor, col: torch.Tensor, w: Optional[torch.Tensor] = None,
         n_out: Optional[int] = None, heads: int = 1) -> torch.Tensor:
    """x [n_src, heads*D] fp32, row/col [E] int64, w [E] / [E, heads] / None -> [n_out, heads*D]."""
    n_out = int(x.size(0) if n_out is None else n_out)
    x = x.float()
    if w is not None:
        w = w.float().reshape(row.numel(), heads)
    if x.is_cuda and _native_for(x) is not None:
        return _SpMM.apply(x.contiguous(), row.contiguous(), col.contiguous(), w, n_out, int(heads))
    return _spmm_torch(x, row, col, w, n_out, int(heads))


class Relabel(object):
    """Unique ids in FIRST-OCCURRENCE order + id -> compact index map (so seeds placed first in the input
    get the indices 0..B-1, the convention of PyG/DGL mini-batch loaders).  ``lookup(q)`` returns the
    compact index of arbitrary ids (-1 when absent) - the sorted-set intersection of the subgraph sampler
    without sorting."""

    def __init__(self, ids: torch.Tensor, sync_free: bool = False):
        """``sync_free`` (CUDA): ``uniq`` keeps the upper-bound length n (-1 padded) and ``n_unique`` stays on the device - no
        host sync, CUDA-graph capturable."""
        flat = ids.reshape(-1).to(torch.int64)
        self._shape = tuple(ids.shape)
        C = _native_for(flat)
        if C is not None:
            self.uniq, inv, self._keys, self._rank, self.n_unique = C.relabel(flat, bool(sync_free))
            self._C = C
        else:
            self._C = None
            valid = flat >= 0
            v = flat[valid]
            su, sinv = torch.unique(v, return_inverse=True)
            first = torch.full((su.numel(),), v.numel(), dtype=torch.int64, device=flat.device)
            first.scatter_reduce_(0, sinv, torch.arange(v.numel(), device=flat.device), reduce="amin")
            order = torch.argsort(first)
            rank_of_sorted = torch.empty_like(order)
            rank_of_sorted[order] = torch.arange(order.numel(), device=flat.device)
            self.uniq = su[order]
            inv = torch.full_like(flat, -1)
            inv[valid] = rank_of_sorted[sinv]
            self._sorted, self._rank_of_sorted = su, rank_of_sorted
        self.inverse = inv.reshape(self._shape)

    def lookup(self, q: torch.Tensor) -> torch.Tensor:
        flat = q.reshape(-1).to(torch.int64)
        if self._C is not None:
            return self._C.relabel_lookup(self._keys, self._rank, flat).reshape(q.shape)
        n = self._sorted.numel()
        if n == 0:
            return torch.full_like(flat, -1).reshape(q.shape)
        pos = torch.searchsorted(self._sorted, flat).clamp_(max=n - 1)
        hit = self._sorted[pos] == flat
        return torch.where(hit, self._rank_of_sorted[pos], torch.full_like(pos, -1)).reshape(q.shape)
