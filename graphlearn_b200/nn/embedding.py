"""K9: id-embedding tables sharded over the GPUs of the box.

Tables for id features (node2vec, bipartite SAGE with hashed ids) can exceed one GPU
(10^8 x 128); the reference keeps them as partitioned TF variables on parameter servers updated
asynchronously (``AdamAsyncOptimizer``, graphlearn/examples/tf/trainer.py:111-116,366-369;
feature_column.py:128-157).  Here the table is hash-partitioned by ``id % world`` in symmetric
memory: the forward is the peer-memory row gather (K5) and the backward applies the sparse update
straight to the owning GPU's rows over NVLink - SGD with atomics (``scatter_add_rows_kernel``) or
Adam on the touched rows with peer-resident moment tables (``sparse_adam_rows_kernel``) - the same
asynchronous-PS semantics, no dense gradient and no all-reduce."""
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
        ctx.emb._apply_sparse_update(ids, g)
        return torch.zeros_like(ctx.emb.anchor), None, None


class ShardedEmbedding(nn.Module):
    def __init__(self, rt: Runtime, num_embeddings: int, dim: int, lr: float = 0.05, init_scale: float = None,
                 optimizer: str = "sgd", betas=(0.9, 0.999), eps: float = 1e-8):
        super().__init__()
        assert optimizer in ("sgd", "adam")
        self.rt, self.num, self.dim, self.lr = rt, int(num_embeddings), int(dim), float(lr)
        self.optimizer, self.betas, self.eps, self.opt_step = optimizer, (float(betas[0]), float(betas[1])), float(eps), 0
        W, r = rt.world, rt.rank
        n_local = (self.num - r + W - 1) // W
        stride = (dim + 3) // 4 * 4
        self.table = rt.symm_empty((n_local, stride), torch.float32)
        s = init_scale if init_scale is not None else 0.5 / dim
        g = torch.Generator(device=rt.device).manual_seed(1234 + r)
        self.table.local[:, :dim] = (torch.rand(n_local, dim, device=rt.device, generator=g) * 2 - 1) * s
        rt.barrier()
        self.desc = make_table_desc(W, dim, stride, torch.float32, self.table.nrows, self.table.ptrs)
        if optimizer == "adam":                      # first / second moments, sharded and peer mapped like the weights
            self.m_tab = rt.symm_empty((n_local, stride), torch.float32)
            self.v_tab = rt.symm_empty((n_local, stride), torch.float32)
            rt.barrier()
            self.m_desc = make_table_desc(W, dim, stride, torch.float32, self.m_tab.nrows, self.m_tab.ptrs)
            self.v_desc = make_table_desc(W, dim, stride, torch.float32, self.v_tab.nrows, self.v_tab.ptrs)
        # autograd anchor: makes the lookup part of the graph although the table itself is updated in place
        self.anchor = nn.Parameter(torch.zeros(1, device=rt.device))

    def _gather(self, ids: torch.Tensor) -> torch.Tensor:
        out = G.gather_rows(self.rt, self.table, self.desc, ids.reshape(-1), self.dim)
        return out.reshape(tuple(ids.shape) + (self.dim,))

    def _apply_sparse_update(self, ids: torch.Tensor, grad: torch.Tensor):
        if self.optimizer == "adam":
            return self._apply_sparse_adam(ids, grad)
        return self._apply_sparse_sgd(ids, grad)

    def _apply_sparse_adam(self, ids: torch.Tensor, grad: torch.Tensor):
        """Adam on the rows touched by this batch (duplicates are summed first); global step for the bias correction"""
        self.opt_step += 1
        flat = ids.reshape(-1)
        g = grad.reshape(-1, self.dim).float()
        uniq, inv = torch.unique(flat, return_inverse=True)
        gs = torch.zeros(uniq.numel(), self.dim, device=g.device).index_add_(0, inv, g)
        b1, b2 = self.betas
        if self.rt.is_cuda:
            native().sparse_adam_rows(self.desc, self.m_desc, self.v_desc, uniq, gs, self.lr, b1, b2, self.eps, self.opt_step)
            return
        bc1, bc2 = 1.0 - b1 ** self.opt_step, 1.0 - b2 ** self.opt_step
        W = self.rt.world

        def upd(v, gg):
            rows = torch.div(v, W, rounding_mode="floor")
            m, vv, w = self.m_tab.local[:, :self.dim], self.v_tab.local[:, :self.dim], self.table.local[:, :self.dim]
            m[rows] = b1 * m[rows] + (1 - b1) * gg
            vv[rows] = b2 * vv[rows] + (1 - b2) * gg * gg
            w[rows] -= (self.lr / bc1) * m[rows] / (vv[rows].sqrt() / math.sqrt(bc2) + self.eps)
            return (torch.zeros(v.numel(), 1, device=v.device),)
        if W == 1:
            upd(uniq, gs)
        else:
            from ..parallel import partition as part
            part.remote_apply(uniq, upd, W, extra=(gs,))

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
