"""Fused GraphSAGE layer op (K6+K7): gather -> aggregate -> tcgen05 GEMM -> bias/ReLU.

CUDA path = ``csrc/sage_fused.cu`` wrapped in an autograd Function; the
backward uses the bf16 A tile saved by the forward kernel (dW = dY^T A,
dA = dY W) - plain library GEMMs.  The portable path composes the same math
from torch ops in fp32 and doubles as the numerics oracle.

Math: EgoSAGEConv, graphlearn/python/nn/tf/layers/ego_sage_conv.py:71-106.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from .. import config as _config
from ..parallel.runtime import local_table_desc, native

MODE = {"mean": 0, "sum": 1, "gcn": 2}


def _pad_n(n_out: int) -> int:
    return max(64, (n_out + 63) // 64 * 64)


def fused_supported(d_self: int, d_nbr: int, n_out: int, mode: str) -> bool:
    if max(d_self, d_nbr) > 512 or n_out > 256:
        return False
    C = native()
    kt = (0 if mode == "gcn" else C.sage_pad_k(d_self)) + C.sage_pad_k(d_nbr)
    return C.sage_smem_bytes(kt, _pad_n(n_out)) <= 232448


class _SageFusedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, bias, x_self, x_nbr, tself_desc, self_vids, tnbr_desc, nbr_vids, M, k, mode,
                relu, out_bf16):
        C = native()
        n_out = weight.size(0)
        if x_self is not None:
            x_self = x_self.contiguous()
            tself_desc, self_vids = local_table_desc(x_self), None
        if x_nbr is not None:
            x_nbr = x_nbr.contiguous()
            tnbr_desc, nbr_vids = local_table_desc(x_nbr), None
        d_self, d_nbr = int(tself_desc[1]), int(tnbr_desc[1])
        kp_self = 0 if mode == "gcn" else C.sage_pad_k(d_self)
        kp_nbr = C.sage_pad_k(d_nbr)
        kt = kp_self + kp_nbr
        N = _pad_n(n_out)
        wp = torch.zeros(n_out, kt, dtype=torch.bfloat16, device=weight.device)
        if mode == "gcn":
            wp[:, :d_nbr] = weight
        else:
            wp[:, :d_self] = weight[:, :d_self]
            wp[:, kp_self:kp_self + d_nbr] = weight[:, d_self:]
        img = C.pack_weight_sw128(wp, N)
        bp = None
        if bias is not None:
            bp = torch.zeros(N, dtype=torch.float32, device=weight.device)
            bp[:n_out] = bias
        need_grad = any(ctx.needs_input_grad[:4])
        out, a_save = C.sage_fused_forward(tself_desc, self_vids, tnbr_desc, nbr_vids, int(M), int(k), MODE[mode], img,
                                           bp, N, n_out, bool(relu), bool(out_bf16), bool(need_grad))
        ctx.save_for_backward(a_save if need_grad else None, wp, out if relu else None)
        ctx.meta = (d_self, d_nbr, kp_self, kp_nbr, int(M), int(k), mode, relu, bias is not None,
                    x_self is not None, x_nbr is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        a_save, wp, out = ctx.saved_tensors
        d_self, d_nbr, kp_self, kp_nbr, M, k, mode, relu, has_bias, dense_self, dense_nbr = ctx.meta
        g = grad_out
        if relu:
            g = g * (out > 0).to(g.dtype)
        g16 = g.to(torch.bfloat16)
        dw = db = dxs = dxn = None
        if ctx.needs_input_grad[0]:
            dwp = (g16.t() @ a_save).float()                  # [n_out, K_total]
            if mode == "gcn":
                dw = dwp[:, :d_nbr].contiguous()
            else:
                dw = torch.cat([dwp[:, :d_self], dwp[:, kp_self:kp_self + d_nbr]], 1)
        if has_bias and ctx.needs_input_grad[1]:
            db = g.float().sum(0)
        need_xs = dense_self and ctx.needs_input_grad[2]
        need_xn = dense_nbr and ctx.needs_input_grad[3]
        if need_xs or need_xn:
            da = g16 @ wp                                      # [M, K_total]
            if mode == "gcn":
                scale = 1.0 / (k + 1)
                d_agg = da[:, :d_nbr] * scale
                if need_xs:
                    dxs = d_agg
                if need_xn:
                    dxn = d_agg[:, None, :].expand(M, k, d_nbr).reshape(M * k, d_nbr)
            else:
                if need_xs:
                    dxs = da[:, :d_self]
                if need_xn:
                    scale = (1.0 / k) if mode == "mean" else 1.0
                    dn = da[:, kp_self:kp_self + d_nbr] * scale
                    dxn = dn[:, None, :].expand(M, k, d_nbr).reshape(M * k, d_nbr)
        return (dw, db, dxs, dxn) + (None,) * 9


def sage_layer_reference(weight, bias, xs, xn, k, mode="mean", relu=False):
    """Pure torch math on already-gathered rows: xs [M, ds], xn [M*k, dn]."""
    M = xs.size(0)
    xn3 = xn.view(M, k, -1).float()
    xs = xs.float()
    if mode == "gcn":
        a = (xs + xn3.sum(1)) / (k + 1)
    else:
        agg = xn3.mean(1) if mode == "mean" else xn3.sum(1)
        a = torch.cat([xs, agg], 1)
    y = F.linear(a, weight.float(), None if bias is None else bias.float())
    return F.relu(y) if relu else y


def sage_layer(weight: torch.Tensor, bias: Optional[torch.Tensor], *, k: int, mode: str = "mean",
               relu: bool = False, out_bf16: bool = False, x_self: Optional[torch.Tensor] = None,
               x_nbr: Optional[torch.Tensor] = None, self_table=None, self_vids: Optional[torch.Tensor] = None,
               nbr_table=None, nbr_vids: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One GraphSAGE layer.  Each of {self, nbr} is either a dense local matrix
    (``x_*``; gradients flow) or (node table, vids) read from the sharded store.

    ``*_table`` are :class:`graphlearn_b200.store.shards.NodeTable`.
    """
    from . import gather as G
    M = x_self.size(0) if x_self is not None else int(self_vids.numel())
    dev = weight.device
    d_self = x_self.size(1) if x_self is not None else self_table.float_dim
    d_nbr = x_nbr.size(1) if x_nbr is not None else nbr_table.float_dim
    use_cuda = dev.type == "cuda" and _config.get().use_peer_kernels
    if use_cuda and fused_supported(d_self, d_nbr, weight.size(0), mode):
        return _SageFusedFn.apply(weight, bias, x_self, x_nbr,
                                  None if self_table is None else self_table.feat_desc, self_vids,
                                  None if nbr_table is None else nbr_table.feat_desc, nbr_vids,
                                  M, k, mode, relu, out_bf16)
    # portable / unfused path
    xs = x_self if x_self is not None else G.gather_rows(self_table.rt, self_table.feats, self_table.feat_desc,
                                                         self_vids, d_self)
    xn = x_nbr if x_nbr is not None else G.gather_rows(nbr_table.rt, nbr_table.feats, nbr_table.feat_desc,
                                                       nbr_vids, d_nbr)
    y = sage_layer_reference(weight, bias, xs, xn, k, mode, relu)
    return y.to(torch.bfloat16) if out_bf16 else y
