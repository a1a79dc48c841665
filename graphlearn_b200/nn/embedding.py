"""K9: id-embedding tables sharded over the GPUs of the box.

Tables for id features (node2vec, bipartite SAGE with hashed ids) can exceed one GPU
(10^8 x 128); the reference keeps them as partitioned TF variables on parameter servers updated
asynchronously (``AdamAsyncOptimizer``, graphlearn/examples/tf/trainer.py:111-116,366-369;
feature_column.py:128-157).  Here the table is hash-partitioned by ``id % world`` in symmetric
memory: the forward is the peer-memory row gather (K5) and the backward applies the sparse SGD
update straight to the owning GPU's rows with atomics over NVLink (``scatter_add_rows_kernel``) -
the same asynchronous-PS semantics, no dense gradient and no all-reduce."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ..ops import gather as G
from ..parallel.runtime import Runtime, make_table_desc, native


class _ShardedLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, ids, emb):
        ctx.emb = emb
        ctx.save_for_backward(ids)
        return emb._gather(ids)

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        ctx.emb._apply_sparse_sgd(ids, g)
        return torch.zeros_like(ctx.emb.anchor), None, None


class ShardedEmbedding(nn.Module):
    def __init__(self, rt: Runtime, num_embeddings: int, dim: int, lr: float = 0.05, init_scale: float = None):
        super().__init__()
        self.rt, self.num, self.dim, self.lr = rt, int(num_embeddings), int(dim), float(lr)
        W, r = rt.world, rt.rank
        n_local = (self.num - r + W - 1) // W
        stride = (dim + 3) // 4 * 4
        self.table = rt.symm_empty((n_local, stride), torch.float32)
        s = init_scale if init_scale is not None else 0.5 / dim
        g = torch.Generator(device=rt.device).manual_seed(1234 + r)
        self.table.local[:, :dim] = (torch.rand(n_local, dim, device=rt.device, generator=g) * 2 - 1) * s
        rt.barrier()
        self.desc = make_table_desc(W, dim, stride, torch.float32, self.table.nrows, self.table.ptrs)
        # autograd anchor: makes the lookup part of the graph although the table itself is updated in place
        self.anchor = nn.Parameter(torch.zeros(1, device=rt.device))

    def _gather(self, ids: torch.Tensor) -> torch.Tensor:
        out = G.gather_rows(self.rt, self.table, self.desc, ids.reshape(-1), self.dim)
        return out.reshape(tuple(ids.shape) + (self.dim,))

    def _apply_sparse_sgd(self, ids: torch.Tensor, grad: torch.Tensor):
        flat = ids.reshape(-1)
        g = grad.reshape(-1, self.dim).float().contiguous()
        if self.rt.is_cuda:
            native().scatter_add_rows(self.desc, flat, g, -self.lr)
        elif self.rt.world == 1:
            self.table.local[:, :self.dim].index_add_(0, flat, g, alpha=-self.lr)
        else:
            from ..parallel import partition as part
            W = self.rt.world

            def upd(v, gg):
                self.table.local[:, :self.dim].index_add_(0, torch.div(v, W, rounding_mode="floor"), gg, alpha=-self.lr)
                return (torch.zeros(v.numel(), 1, device=v.device),)

            part.remote_apply(flat, upd, W, extra=(g,))

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        return _ShardedLookup.apply(self.anchor, ids, self)

    def local_weight(self) -> torch.Tensor:
        return self.table.local[:, :self.dim]
