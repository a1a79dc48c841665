"""GNN layers on EgoGraphs (dense, fixed fan-out) - PyTorch modules.

Re-implements the math of the reference's TF1 layers
(graphlearn/python/nn/tf/layers/ego_{sage,gat,gin,rgcn}_conv.py)
with the SAGE layer running on the fused sm_100a kernel.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops import sage as sage_ops


class EgoSAGEConv(nn.Module):
    """out = W . [x || agg(nbrs)]  (agg in mean|sum|max) or 'gcn' = W . mean({x} U nbrs).

    ego_sage_conv.py:71-106.  ``forward`` takes already-gathered dense tensors;
    ``forward_store`` reads rows straight from the sharded feature store inside
    the fused kernel (no [B*k, D] intermediate)."""

    def __init__(self, in_dim, out_dim, agg_type="mean", bias=True, in_nbr_dim=None):
        super().__init__()
        self.in_self = int(in_dim if not isinstance(in_dim, (tuple, list)) else in_dim[0])
        self.in_nbr = int(in_nbr_dim if in_nbr_dim is not None else
                          (in_dim[1] if isinstance(in_dim, (tuple, list)) else in_dim))
        self.out_dim = int(out_dim)
        self.agg_type = agg_type
        assert agg_type in ("mean", "sum", "max", "gcn")
        k_in = self.in_nbr if agg_type == "gcn" else self.in_self + self.in_nbr
        self._mode = "mean" if agg_type == "max" else agg_type
        w = torch.empty(out_dim, k_in)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        # stored in the fused kernel's padded K layout (pad columns are zero and stay zero)
        self.weight_p = nn.Parameter(sage_ops.pad_weight(w, self.in_self, self.in_nbr, self._mode))
        self.bias = nn.Parameter(torch.zeros(out_dim)) if bias else None

    @property
    def weight(self):
        """Logical [out, in_self + in_nbr] weight (gcn: [out, in])."""
        return sage_ops.logical_weight(self.weight_p, self.in_self, self.in_nbr, self._mode)

    def forward(self, x, neighbor, expand, relu=False, out_bf16=False):
        """x [M, d], neighbor [M*k, d_n], expand = k."""
        if self.agg_type == "max":
            agg = neighbor.view(-1, expand, neighbor.size(-1)).float().max(1).values
            y = F.linear(torch.cat([x.float(), agg], 1), self.weight, self.bias)
            y = F.relu(y) if relu else y
            return y.to(torch.bfloat16) if out_bf16 else y
        return sage_ops.sage_layer(self.weight_p, self.bias, k=expand, mode=self.agg_type, relu=relu,
                                   out_bf16=out_bf16, x_self=x, x_nbr=neighbor)

    def forward_store(self, table, self_vids, nbr_vids, expand, relu=False, out_bf16=False, nbr_table=None):
        if self.agg_type == "max":
            from ..ops import gather as G
            nt = nbr_table or table
            xs = G.gather_rows(table.rt, table.feats, table.feat_desc, self_vids, table.float_dim)
            xn = G.gather_rows(nt.rt, nt.feats, nt.feat_desc, nbr_vids, nt.float_dim)
            return self.forward(xs, xn, expand, relu, out_bf16)
        return sage_ops.sage_layer(self.weight_p, self.bias, k=expand, mode=self.agg_type, relu=relu,
                                   out_bf16=out_bf16, self_table=table, self_vids=self_vids.reshape(-1),
                                   nbr_table=nbr_table or table, nbr_vids=nbr_vids.reshape(-1))
