"""Hash partition -> remote process -> stitch over ``torch.distributed``.

This is the *baseline / portable* implementation of the reference's
Partition -> Process -> Stitch operator pattern
(graphlearn/src/core/runner/op_runner.h:60-83, hash_partitioner.h:33-92,
stitcher.h:48-108): ids are bucketed by owner = |id| % world, exchanged with one
all-to-all, processed on the owner and sent back with a second all-to-all.  It
runs on gloo (CPU tests, world_size 2) and on NCCL (the A/B baseline for the
fused peer-memory kernels, which need neither collective).
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def owner_of(ids: torch.Tensor, world: int) -> torch.Tensor:
    """owner = llabs(id) % world  (graphlearn/src/core/partition/hash_partitioner.h:90-92)."""
    return ids.abs() % world


def partition_by_owner(ids: torch.Tensor, world: int):
    """Returns (sorted_ids, order, counts): `order` is the permutation that groups
    ids by owner (stable), counts[r] the bucket sizes (R1/R2: Shards + Sticker)."""
    owner = owner_of(ids, world)
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world)
    return ids[order], order, counts


def _all_to_all_v(send: torch.Tensor, send_counts: List[int], recv_counts: List[int]) -> torch.Tensor:
    out_shape = (int(sum(recv_counts)),) + tuple(send.shape[1:])
    recv = torch.empty(out_shape, dtype=send.dtype, device=send.device)
    if dist.get_backend() == "gloo" and send.dim() > 0:
        # gloo all_to_all_single with splits is supported on CPU tensors
        dist.all_to_all_single(recv, send.contiguous(), recv_counts, send_counts)
    else:
        dist.all_to_all_single(recv, send.contiguous(), recv_counts, send_counts)
    return recv


def exchange_counts(counts: torch.Tensor) -> List[int]:
    world = dist.get_world_size()
    c = counts.to(torch.int64)
    out = torch.empty_like(c)
    dist.all_to_all_single(out, c, [1] * world, [1] * world)
    return [int(x) for x in out.tolist()]


def remote_apply(ids: torch.Tensor, fn: Callable[[torch.Tensor], Sequence[torch.Tensor]], world: int,
                 extra: Sequence[torch.Tensor] = ()) -> List[torch.Tensor]:
    """Run ``fn(ids_owned_here, *extra_owned_here)`` on each id's owner and return
    the outputs in the caller's original order.  Every output of ``fn`` must have
    one leading row per input id (dense, fixed fan-out responses)."""
    if world == 1:
        return list(fn(ids, *extra))
    sorted_ids, order, counts = partition_by_owner(ids, world)
    send_counts = [int(x) for x in counts.tolist()]
    recv_counts = exchange_counts(counts)
    got_ids = _all_to_all_v(sorted_ids, send_counts, recv_counts)
    got_extra = [_all_to_all_v(e[order], send_counts, recv_counts) for e in extra]
    outs = fn(got_ids, *got_extra)
    results = []
    inv = torch.empty_like(order)
    inv[order] = torch.arange(order.numel(), device=order.device)
    for o in outs:
        back = _all_to_all_v(o, recv_counts, send_counts)
        results.append(back[inv])
    return results


def shuffle_rows(cols: dict, owner: torch.Tensor, world: int, device) -> dict:
    """Load-time shuffle (C4 / K11): row i of every column goes to rank ``owner[i]``.

    Reference: loader threads batch parsed rows into UpdateNodes/UpdateEdges requests that the
    distributed op runner routes to the owning server (graphlearn/src/core/graph/graph_store.cc:
    60-165,210-250; shard keys graph_update_request.cc:151-156,234-237).  Here: one all-to-all-v
    per column over NCCL/gloo; object (string) columns travel pickled."""
    if world == 1:
        return dict(cols)
    import numpy as np
    owner = owner.to(device)
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world)
    send_counts = [int(x) for x in counts.tolist()]
    recv_counts = exchange_counts(counts)
    out = {}
    for k, v in cols.items():
        if v is None:
            out[k] = None
        elif isinstance(v, torch.Tensor):
            out[k] = _all_to_all_v(v.to(device)[order], send_counts, recv_counts)
        else:
            vs = np.asarray(v, dtype=object)[order.cpu().numpy()]
            pieces = np.split(vs, np.cumsum(send_counts)[:-1])
            gathered = [None] * world
            dist.all_gather_object(gathered, pieces)
            me = dist.get_rank()
            out[k] = np.concatenate([g[me] for g in gathered]) if gathered else vs[:0]
    return out
