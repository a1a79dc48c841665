"""Feature gather (K5) and gather-aggregate (K6) over sharded tables.

CUDA: one kernel dereferences local or peer rows directly (csrc/gather.cu).
Portable path: partition -> all-to-all -> local index_select -> all-to-all
(the reference's LookupNodes / Aggregating requests,
graphlearn/src/service/request/graph_lookup_request.cc:330-334,
aggregating_request.cc:172-213).
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import config as _config
from ..parallel import partition as part
from ..parallel.runtime import Runtime, SymmTensor, native

AGG = {"sum": 0, "mean": 1, "max": 2, "min": 3, "prod": 4}


def _local_rows(st: SymmTensor, vids: torch.Tensor, world: int, fill):
    """rows of a local shard for vids OWNED by this rank (missing -> fill)."""
    n = st.local.size(0)
    rows = torch.div(vids, world, rounding_mode="floor")
    ok = (vids >= 0) & (rows < n)
    rc = torch.where(ok, rows, torch.zeros_like(rows))
    if n == 0:
        shape = (vids.numel(),) + tuple(st.local.shape[1:])
        return torch.full(shape, fill, dtype=st.local.dtype, device=vids.device)
    out = st.local[rc]
    if out.dim() > 1:
        okb = ok.view(-1, *([1] * (out.dim() - 1)))
    else:
        okb = ok
    return torch.where(okb, out, torch.full_like(out, fill))


def gather_rows(rt: Runtime, st: SymmTensor, desc: Optional[torch.Tensor], vids: torch.Tensor, dim: int,
                out_dtype: torch.dtype = torch.float32, fill: float = 0.0) -> torch.Tensor:
    """out[i] = table[vid_i][:dim]  (float tables; fp32 or bf16 storage)."""
    v = vids.reshape(-1).to(torch.int64)
    if rt.is_cuda and _config.get().use_peer_kernels and desc is not None:
        return native().gather_rows(desc, v, out_dtype == torch.bfloat16, float(fill))
    if st.local.dtype == torch.uint8:
        # fp8 block-scaled table on the portable path: fetch the raw byte rows, dequantise here
        from ..store.shards import NodeTable
        (raw,) = part.remote_apply(v, lambda x: (_local_rows(st, x, rt.world, 0),), rt.world)
        return NodeTable.dequantize_fp8_rows(raw, dim).to(out_dtype)
    cmap = getattr(st, "cache_map", None)
    if cmap is not None and rt.world > 1:
        # replica cache (N17): hits are served from local HBM, only misses travel
        inb = (v >= 0) & (v < cmap.numel())
        slot = torch.where(inb, cmap[v.clamp(min=0, max=cmap.numel() - 1)].long(), torch.full_like(v, -1))
        hit = slot >= 0
        (rows,) = part.remote_apply(torch.where(hit, torch.full_like(v, -1), v),
                                    lambda x: (_local_rows(st, x, rt.world, fill)[:, :dim],), rt.world)
        rows = torch.where(hit[:, None], st.cache_rows[slot.clamp(min=0)][:, :dim].to(rows.dtype), rows)
        return rows.to(out_dtype)
    (rows,) = part.remote_apply(v, lambda x: (_local_rows(st, x, rt.world, fill)[:, :dim],), rt.world)
    return rows.to(out_dtype)


def gather_rows_dedup(rt: Runtime, st: SymmTensor, desc: Optional[torch.Tensor], vids: torch.Tensor, dim: int,
                      out_dtype: torch.dtype = torch.float32, fill: float = 0.0) -> torch.Tensor:
    """Dedup before the pull: relabel the requested ids (K4 hash table, first-occurrence order), fetch
    every DISTINCT row once - over NVLink when it is remote - and expand by index.  Pays off when a frontier repeats ids
    (power-law graphs: a hub is sampled by many seeds; multi-hop ego networks of one batch); on the uniform synthetic
    benchmark graph only ~6 % of the 282 K ids of a batch repeat, so the fused engine does not use it."""
    from .sparse import Relabel
    v = vids.reshape(-1).to(torch.int64)
    if v.numel() == 0:
        return gather_rows(rt, st, desc, v, dim, out_dtype, fill)
    rl = Relabel(v)
    uniq_rows = gather_rows(rt, st, desc, rl.uniq, dim, out_dtype, fill)
    inv = rl.inverse.reshape(-1)
    out = uniq_rows[inv.clamp(min=0)]
    if bool((inv < 0).any()):                       # negative (padding) ids are not part of the table
        out = torch.where((inv >= 0)[:, None], out, torch.full_like(out, fill))
    return out


def gather_any(rt: Runtime, st: SymmTensor, vids: torch.Tensor, fill) -> torch.Tensor:
    """Generic (int64 / float / 1-D or 2-D) sharded lookup.  On CUDA, 8-byte rows are
    moved bit-exactly by the float gather kernel (pure copies, no arithmetic)."""
    v = vids.reshape(-1).to(torch.int64)
    loc = st.local
    if rt.is_cuda and _config.get().use_peer_kernels and loc.dtype in (torch.int64, torch.float32):
        from ..parallel.runtime import make_table_desc
        width = 1 if loc.dim() == 1 else int(loc.size(1))
        f32_per = 2 if loc.dtype == torch.int64 else 1
        dimf = width * f32_per
        # stride (in fp32 elements) must keep rows 16-byte aligned for the vector path
        if (dimf % 4) == 0:
            desc = make_table_desc(rt.world, dimf, dimf, torch.float32, st.nrows, st.ptrs)
            out = native().gather_rows(desc, v, False, 0.0)
            res = out.view(loc.dtype).reshape((v.numel(),) + tuple(loc.shape[1:]))
            if fill != 0:
                W = rt.world
                nrows = torch.tensor(st.nrows, device=v.device)
                ok = (v >= 0) & (torch.div(v, W, rounding_mode="floor") < nrows[v.clamp(min=0) % W])
                okb = ok.view(-1, *([1] * (res.dim() - 1)))
                res = torch.where(okb, res, torch.full_like(res, fill))
            return res
    (rows,) = part.remote_apply(v, lambda x: (_local_rows(st, x, rt.world, fill),), rt.world)
    return rows


def gather_agg(rt: Runtime, st: SymmTensor, desc: Optional[torch.Tensor], vids: torch.Tensor, dim: int,
               mode: str = "mean", offsets: Optional[torch.Tensor] = None, k: int = 0) -> torch.Tensor:
    """Segment reduce of gathered rows: dense [S, k] ids or ragged (offsets[S+1])."""
    v = vids.reshape(-1).to(torch.int64)
    if rt.is_cuda and _config.get().use_peer_kernels and desc is not None:
        return native().gather_agg(desc, v, offsets, int(k), AGG[mode])
    rows = gather_rows(rt, st, desc, v, dim)
    return segment_reduce(rows, mode, offsets=offsets, k=k)


def segment_reduce(rows: torch.Tensor, mode: str, offsets: Optional[torch.Tensor] = None, k: int = 0) -> torch.Tensor:
    """Portable segment reduce used as the oracle for the kernels."""
    d = rows.size(1)
    if offsets is None:
        x = rows.view(-1, k, d).float()
        if mode == "sum":
            return x.sum(1)
        if mode == "mean":
            return x.mean(1)
        if mode == "max":
            return x.max(1).values
        if mode == "min":
            return x.min(1).values
        if mode == "prod":
            return x.prod(1)
        raise ValueError(mode)
    S = offsets.numel() - 1
    lens = offsets[1:] - offsets[:-1]
    seg = torch.repeat_interleave(torch.arange(S, device=rows.device), lens)
    out = torch.zeros(S, d, dtype=torch.float32, device=rows.device)
    x = rows.float()
    if mode in ("sum", "mean"):
        out.index_add_(0, seg, x)
        if mode == "mean":
            out = out / lens.clamp(min=1)[:, None].float()
        return out
    red = {"max": "amax", "min": "amin", "prod": "prod"}[mode]
    init = {"max": float("-inf"), "min": float("inf"), "prod": 1.0}[mode]
    out.fill_(init)
    out.scatter_reduce_(0, seg[:, None].expand(-1, d), x, reduce=red, include_self=True)
    out[lens == 0] = 0.0
    return out
