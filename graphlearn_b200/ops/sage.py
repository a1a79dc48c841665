"""Fused GraphSAGE layer op (K6+K7): gather -> aggregate -> tcgen05 GEMM -> bias/ReLU.

Weights are kept in the kernel's padded K layout ``[n_out, kp_self + kp_nbr]``
(each half padded to 64/128/256/512 columns; pad columns stay zero because
their gradient is exactly zero), so one tiny kernel per step turns the fp32
master weight into the bf16 SWIZZLE_128B image the TMA engine streams into
shared memory.  CUDA path = ``csrc/sage_fused.cu`` behind an autograd Function;
the backward uses the bf16 A tile saved by the forward kernel (dW = dY^T A,
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
_SMEM_LIMIT = 232448


def pad_k(d: int) -> int:
    if d <= 0:
        return 0
    for p in (64, 128, 256, 512):
        if d <= p:
            return p
    raise ValueError("fused SAGE layer supports feature dims up to 512")


def pad_n(n_out: int) -> int:
    return max(64, (n_out + 63) // 64 * 64)


def padded_dims(d_self: int, d_nbr: int, mode: str):
    kp_self = 0 if mode == "gcn" else pad_k(d_self)
    return kp_self, pad_k(d_nbr)


def fused_supported(d_self: int, d_nbr: int, n_out: int, mode: str, k: int = 0, dtype=torch.bfloat16) -> bool:
    if max(d_self, d_nbr) > 512 or n_out > 256:
        return False
    kp_self, kp_nbr = padded_dims(d_self, d_nbr, mode)
    # a gather warp keeps the row pointers of one work item (1-4 destination rows) in a 64-entry scratch
    lanes_row = max(kp_self, kp_nbr) // (4 if dtype == torch.float32 else 8)
    rows_per_item = 32 // lanes_row if lanes_row < 32 else 1
    if rows_per_item * (k + 1) > 64:
        return False
    return smem_fits(kp_self + kp_nbr, pad_n(n_out))


def smem_fits(k_total: int, n_pad: int) -> bool:
    """A tile + W image + pointer staging of the persistent kernel (csrc/sage_fused.cu sage_smem_bytes)."""
    return 1024 + (k_total // 64) * (128 * 128 + n_pad * 128) + 23 * 64 * 8 + 1024 + 72 <= _SMEM_LIMIT


def logical_weight(weight_p: torch.Tensor, d_self: int, d_nbr: int, mode: str) -> torch.Tensor:
    """[n_out, d_self + d_nbr] (or [n_out, d] for gcn) view of a padded weight."""
    kp_self, _ = padded_dims(d_self, d_nbr, mode)
    if mode == "gcn":
        return weight_p[:, :d_nbr]
    return torch.cat([weight_p[:, :d_self], weight_p[:, kp_self:kp_self + d_nbr]], 1)


def _aligned_rows(x: torch.Tensor) -> torch.Tensor:
    """[n, d] view whose rows start on 16-byte boundaries (the kernel moves 16-byte chunks)."""
    esz = x.element_size()
    if x.stride(1) == 1 and (x.stride(0) * esz) % 16 == 0 and x.data_ptr() % 16 == 0:
        return x
    d = x.size(1)
    per = 16 // esz
    dp = (d + per - 1) // per * per
    buf = torch.zeros(x.size(0), dp, dtype=x.dtype, device=x.device)
    buf[:, :d] = x
    return buf[:, :d]


class _SageFusedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight_p, bias, x_self, x_nbr, tself_desc, self_vids, tnbr_desc, nbr_vids, M, k, mode,
                relu, out_bf16, rows_per_cta):
        C = native()
        n_out = weight_p.size(0)
        if x_self is not None:
            x_self = _aligned_rows(x_self)
            tself_desc, self_vids = local_table_desc(x_self), None
        if x_nbr is not None:
            x_nbr = _aligned_rows(x_nbr)
            tnbr_desc, nbr_vids = local_table_desc(x_nbr), None
        d_self, d_nbr = int(tself_desc[1]), int(tnbr_desc[1])
        kp_self, kp_nbr = padded_dims(d_self, d_nbr, mode)
        assert weight_p.size(1) == kp_self + kp_nbr, "weight must be in the padded K layout"
        N = pad_n(n_out)
        need_w = ctx.needs_input_grad[0]
        need_x = (x_self is not None and ctx.needs_input_grad[2]) or (x_nbr is not None and ctx.needs_input_grad[3])
        img, w16 = C.pack_weight_f32(weight_p.detach().contiguous(), N, bool(need_x))
        out, a_save = C.sage_fused_forward(tself_desc, self_vids, tnbr_desc, nbr_vids, int(M), int(k), MODE[mode], img,
                                           bias, N, n_out, bool(relu), bool(out_bf16), bool(need_w),
                                           int(rows_per_cta), None, None)
        ctx.save_for_backward(a_save if need_w else None, w16 if need_x else None, out if relu else None)
        ctx.meta = (d_self, d_nbr, kp_self, kp_nbr, int(M), int(k), mode, relu, bias is not None,
                    x_self is not None, x_nbr is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        a_save, w16, out = ctx.saved_tensors
        d_self, d_nbr, kp_self, kp_nbr, M, k, mode, relu, has_bias, dense_self, dense_nbr = ctx.meta
        g = grad_out
        if relu:
            g = g * (out > 0).to(g.dtype)
        g16 = g.to(torch.bfloat16)
        dw = db = dxs = dxn = None
        if ctx.needs_input_grad[0]:
            dw = torch.mm(g16.t(), a_save, out_dtype=torch.float32) if _HAS_OUT_DTYPE else (g16.t() @ a_save).float()
        if has_bias and ctx.needs_input_grad[1]:
            db = g.float().sum(0)
        need_xs = dense_self and ctx.needs_input_grad[2]
        need_xn = dense_nbr and ctx.needs_input_grad[3]
        if need_xs or need_xn:
            da = g16 @ w16                                     # [M, K_total]
            if mode == "gcn":
                d_agg = da[:, :d_nbr] * (1.0 / (k + 1))
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
        return (dw, db, dxs, dxn) + (None,) * 10


# torch.mm(bf16, bf16, out_dtype=fp32) (CUDA only) writes the fp32 gradient without a cast kernel
_HAS_OUT_DTYPE = "out_dtype" in (torch.mm.__doc__ or "")


def pad_weight(weight: torch.Tensor, d_self: int, d_nbr: int, mode: str) -> torch.Tensor:
    """logical [n_out, d_self + d_nbr] (gcn: [n_out, d]) -> padded K layout."""
    kp_self, kp_nbr = padded_dims(d_self, d_nbr, mode)
    wp = torch.zeros(weight.size(0), kp_self + kp_nbr, dtype=weight.dtype, device=weight.device)
    if mode == "gcn":
        wp[:, :d_nbr] = weight
    else:
        wp[:, :d_self] = weight[:, :d_self]
        wp[:, kp_self:kp_self + d_nbr] = weight[:, d_self:]
    return wp


def sage_layer_reference(weight, bias, xs, xn, k, mode="mean", relu=False):
    """Pure torch math on already-gathered rows: xs [M, ds], xn [M*k, dn]; `weight` logical."""
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


def sage_layer(weight_p: torch.Tensor, bias: Optional[torch.Tensor], *, k: int, mode: str = "mean",
               relu: bool = False, out_bf16: bool = False, x_self: Optional[torch.Tensor] = None,
               x_nbr: Optional[torch.Tensor] = None, self_table=None, self_vids: Optional[torch.Tensor] = None,
               nbr_table=None, nbr_vids: Optional[torch.Tensor] = None, rows_per_cta: int = 0) -> torch.Tensor:
    """One GraphSAGE layer.  ``weight_p`` is in the padded K layout (see module doc).
    Each of {self, nbr} is either a dense local matrix (``x_*``; gradients flow) or
    (:class:`NodeTable`, vids) read from the sharded store inside the kernel."""
    from . import gather as G
    M = x_self.size(0) if x_self is not None else int(self_vids.numel())
    dev = weight_p.device
    d_self = x_self.size(1) if x_self is not None else self_table.float_dim
    d_nbr = x_nbr.size(1) if x_nbr is not None else nbr_table.float_dim
    use_cuda = dev.type == "cuda" and _config.get().use_peer_kernels
    dt_self = x_self.dtype if x_self is not None else self_table.feats.local.dtype
    dt_nbr = x_nbr.dtype if x_nbr is not None else nbr_table.feats.local.dtype
    kp_s, kp_n = padded_dims(d_self, d_nbr, mode)
    # torch.uint8 = fp8 block-scaled table storage (store/shards.py): only as (table, vids) operands
    ok_dt = (torch.float32, torch.bfloat16) + ((torch.uint8,) if x_self is None and x_nbr is None else ())
    compatible = dt_self == dt_nbr and dt_self in ok_dt and (kp_s == 0 or kp_s == kp_n)
    if use_cuda and compatible and fused_supported(d_self, d_nbr, weight_p.size(0), mode, k, dt_self):
        return _SageFusedFn.apply(weight_p, bias, x_self, x_nbr,
                                  None if self_table is None else self_table.feat_desc, self_vids,
                                  None if nbr_table is None else nbr_table.feat_desc, nbr_vids,
                                  M, k, mode, relu, out_bf16, rows_per_cta)
    # portable / unfused path
    xs = x_self if x_self is not None else G.gather_rows(self_table.rt, self_table.feats, self_table.feat_desc,
                                                         self_vids, d_self)
    xn = x_nbr if x_nbr is not None else G.gather_rows(nbr_table.rt, nbr_table.feats, nbr_table.feat_desc,
                                                       nbr_vids, d_nbr)
    y = sage_layer_reference(logical_weight(weight_p, d_self, d_nbr, mode), bias, xs, xn, k, mode, relu)
    return y.to(torch.bfloat16) if out_bf16 else y
