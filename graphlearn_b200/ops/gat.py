"""Fused multi-head attention aggregation (K6, GAT flavour) - autograd wrapper + portable oracle.

EgoGATConv of the reference (graphlearn/python/nn/tf/layers/ego_gat_conv.py:89-117), per head h::

    e_j  = LeakyReLU( a_h . [W_x x + b_x || W_n n_j + b_n] + b_a )  =  LeakyReLU( u_x,h . x + u_n,h . n_j + c_h )
    coef = softmax_j(e)                       (over the k neighbours only)
    out  = mean_h ( W_n,h sum_j coef_j n_j + b_n,h )

``gat_aggregate`` returns ``A[m] = [ sum_j coef_1j n_j || ... || sum_j coef_Hj n_j ]`` (each block zero padded to ``kp``
columns, bf16 on the CUDA path): on CUDA one kernel (csrc/gat.cu) pulls the rows straight from the local / peer-mapped
feature shards (or dense activations) and runs the online softmax; the projection with the concatenated ``W_n`` is then
ONE tensor-core GEMM (``ops.linear.tc_linear``).  The backward kernel re-gathers the rows and produces the gradients of
``u_n``, of the self logits (-> ``u_x``, ``c``) and, for dense inputs, of the rows themselves.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from .. import config as _config
from ..parallel.runtime import local_table_desc, native
from . import gather as G


def pad_k(d: int) -> int:
    kp = 64
    while kp < d:
        kp *= 2
    return kp


def kernel_supported(d_self: int, d_nbr: int, dtype: torch.dtype, heads: int) -> bool:
    vec = 4 if dtype == torch.float32 else 8
    return heads <= 4 and dtype in (torch.float32, torch.bfloat16) and pad_k(max(d_self, d_nbr)) // vec <= 32


def gat_aggregate_reference(u_x, u_n, c, xs, xn, k: int, slope: float = 0.2):
    """torch oracle on gathered rows: xs [M, ds], xn [M*k, dn] -> A [M, H*kp] fp32 (zero padded blocks)."""
    M, H = xs.size(0), u_n.size(0)
    xs, xn = xs.float(), xn.float().view(M, k, -1)
    e = F.leaky_relu((xs @ u_x.t())[:, None, :] + xn @ u_n.t() + c, slope)              # [M, k, H]
    coef = torch.softmax(e, dim=1)
    agg = torch.einsum("mkh,mkd->mhd", coef, xn)                                        # [M, H, dn]
    kp = pad_k(max(xs.size(1), xn.size(2)))
    out = torch.zeros(M, H, kp, device=xs.device)
    out[:, :, :agg.size(2)] = agg
    return out.reshape(M, H * kp)


class _GatAggFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u_x, u_n, c, x_self, x_nbr, tself_desc, self_vids, tnbr_desc, nbr_vids, M, k, slope, self_table):
        C = native()
        if x_self is not None:
            x_self = _rows16(x_self)
            tself_desc, self_vids = local_table_desc(x_self), None
        if x_nbr is not None:
            x_nbr = _rows16(x_nbr)
            tnbr_desc, nbr_vids = local_table_desc(x_nbr), None
        a, e, stat = C.gat_agg_forward(tself_desc, self_vids, 0, tnbr_desc, nbr_vids, 0, int(M), int(k), u_x.detach().float().contiguous(),
                                       u_n.detach().float().contiguous(), c.detach().float().contiguous(), float(slope))
        ctx.save_for_backward(u_x.detach(), u_n.detach(), c.detach(), e, stat, x_self, x_nbr, self_vids, nbr_vids)
        ctx.descs = (tself_desc, tnbr_desc)
        ctx.meta = (int(M), int(k), float(slope), self_table)
        return a

    @staticmethod
    def backward(ctx, dA):
        C = native()
        u_x, u_n, c, e, stat, x_self, x_nbr, self_vids, nbr_vids = ctx.saved_tensors
        tself_desc, tnbr_desc = ctx.descs
        M, k, slope, self_table = ctx.meta
        want_dx = x_nbr is not None and ctx.needs_input_grad[4]
        du_n, dsx, dx_nbr = C.gat_agg_backward(tself_desc, tnbr_desc, nbr_vids, 0, M, k, u_x.float().contiguous(), u_n.float().contiguous(),
                                               c.float().contiguous(), slope, dA.to(torch.bfloat16).contiguous(), e, stat, bool(want_dx))
        # self logit s_h(m) = u_x,h . x_m + c_h  ->  du_x = dsx^T X,  dc = sum_m dsx,  dX_self = dsx u_x
        if x_self is not None:
            xs = x_self.float()
        else:
            xs = G.gather_rows(self_table.rt, self_table.feats, self_table.feat_desc, self_vids, self_table.float_dim)
        du_x = dsx.t() @ xs
        dc = dsx.sum(0)
        dxs = (dsx @ u_x.float()).to(x_self.dtype) if (x_self is not None and ctx.needs_input_grad[3]) else None
        dxn = dx_nbr.to(x_nbr.dtype) if want_dx else None
        return (du_x, du_n, dc, dxs, dxn) + (None,) * 8


def _rows16(x: torch.Tensor) -> torch.Tensor:
    """[n, d] view whose rows start on 16-byte boundaries."""
    esz = x.element_size()
    if x.stride(1) == 1 and (x.stride(0) * esz) % 16 == 0 and x.data_ptr() % 16 == 0:
        return x
    per = 16 // esz
    dp = (x.size(1) + per - 1) // per * per
    buf = torch.zeros(x.size(0), dp, dtype=x.dtype, device=x.device)
    buf[:, :x.size(1)] = x
    return buf[:, :x.size(1)]


def gat_aggregate(u_x, u_n, c, *, k: int, slope: float = 0.2, x_self: Optional[torch.Tensor] = None,
                  x_nbr: Optional[torch.Tensor] = None, self_table=None, self_vids=None, nbr_table=None, nbr_vids=None) -> torch.Tensor:
    """A [M, H * kp].  Each side is either a dense local matrix (gradients flow) or (NodeTable, vids)."""
    M = x_self.size(0) if x_self is not None else int(self_vids.numel())
    d_self = x_self.size(1) if x_self is not None else self_table.float_dim
    d_nbr = x_nbr.size(1) if x_nbr is not None else nbr_table.float_dim
    dt_s = x_self.dtype if x_self is not None else self_table.feats.local.dtype
    dt_n = x_nbr.dtype if x_nbr is not None else nbr_table.feats.local.dtype
    dev = u_n.device
    if dev.type == "cuda" and _config.get().use_peer_kernels and dt_s == dt_n and kernel_supported(d_self, d_nbr, dt_n, u_n.size(0)):
        return _GatAggFn.apply(u_x, u_n, c, x_self, x_nbr, None if self_table is None else self_table.feat_desc,
                               None if self_vids is None else self_vids.reshape(-1),
                               None if nbr_table is None else nbr_table.feat_desc,
                               None if nbr_vids is None else nbr_vids.reshape(-1), M, k, slope, self_table)
    xs = x_self if x_self is not None else G.gather_rows(self_table.rt, self_table.feats, self_table.feat_desc, self_vids.reshape(-1), d_self)
    xn = x_nbr if x_nbr is not None else G.gather_rows(nbr_table.rt, nbr_table.feats, nbr_table.feat_desc, nbr_vids.reshape(-1), d_nbr)
    return gat_aggregate_reference(u_x.float(), u_n.float(), c.float(), xs, xn, k, slope)
