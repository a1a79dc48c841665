"""Random walks (K3): DeepWalk (p = q = 1) and node2vec second-order walks.

CUDA: one resident walker thread per seed chasing local / remote adjacency rows
(csrc/walk.cu).  Portable path: one sampling round per step over
``torch.distributed`` - the reference's algorithm
(graphlearn/src/core/operator/random_walk/random_walk.cc:53-135).
Output ``[B, walk_len]`` excludes the seed, like ``RandomWalkResponse``.
"""
from __future__ import annotations

import torch

from .. import config as _config
from ..parallel.runtime import native
from . import rng as _rng
from . import sampling as S


def random_walk(csr, src_vids: torch.Tensor, walk_len: int, p: float = 1.0, q: float = 1.0, rng=None, salt: int = 0):
    cfg = _config.get()
    src = src_vids.reshape(-1).to(torch.int64)
    rng = rng or _rng.default_rng(csr.rt)
    if csr.rt.is_cuda and cfg.use_peer_kernels:
        if not (p == 1.0 and q == 1.0):
            csr.ensure_sorted_rows()            # node2vec: binary-search membership in the parent's (id-sorted) row
        return native().random_walk(csr.desc, src, int(walk_len), float(p), float(q), int(cfg.default_neighbor_id),
                                    int(cfg.default_full_nbr_num), rng.state, int(salt))
    B = int(src.numel())
    out = torch.empty(B, walk_len, dtype=torch.int64, device=src.device)
    cur, prev = src, None
    strategy = "edge_weight" if csr.cumw is not None else "random"
    second = not (p == 1.0 and q == 1.0)
    gen = rng.torch_generator(salt)
    for step in range(walk_len):
        if not second or prev is None:
            nxt, _ = S.sample_neighbors(csr, cur, 1, strategy, want_eids=False, rng=rng, salt=salt * 131 + step)
            nxt = nxt.reshape(-1)
        else:
            nxt = _node2vec_step(csr, cur, prev, p, q, rng, salt * 131 + step, gen, cfg)
        out[:, step] = nxt
        prev, cur = cur, nxt
    return out


def _node2vec_step(csr, cur, prev, p, q, rng, salt, gen, cfg):
    """Rejection sampling with uniform proposals; membership against the parent's (capped) row."""
    B = int(cur.numel())
    dev = cur.device
    wmax = max(1.0, 1.0 / p, 1.0 / q)
    pvals, _, poffs = S.sample_full(csr, prev, cap=cfg.default_full_nbr_num, want_eids=False)
    pcount = poffs[1:] - poffs[:-1]
    seg = torch.repeat_interleave(torch.arange(B, device=dev), pcount)
    base = int(max(int(pvals.max().item()) if pvals.numel() else 0, 1) + 2)
    keys = torch.sort(seg * base + (pvals + 1))[0] if pvals.numel() else torch.zeros(0, dtype=torch.int64, device=dev)
    result = torch.full((B,), cfg.default_neighbor_id, dtype=torch.int64, device=dev)
    pending = torch.ones(B, dtype=torch.bool, device=dev)
    for t in range(64):
        cand, _ = S.sample_neighbors(csr, cur, 1, "random", want_eids=False, rng=rng, salt=salt * 71 + t)
        cand = cand.reshape(-1)
        is_prev = cand == prev
        if keys.numel():
            # a candidate larger than every parent neighbour cannot be a member: mark it instead of clamping it onto
            # the largest neighbour id (which would hand it weight 1 instead of 1/q when that id is in the row)
            in_range = (cand >= 0) & (cand <= base - 2)
            qk = torch.arange(B, device=dev) * base + (cand.clamp(min=0, max=base - 2) + 1)
            pos = torch.searchsorted(keys, qk).clamp_(max=keys.numel() - 1)
            nb = (keys[pos] == qk) & in_range
        else:
            nb = torch.zeros(B, dtype=torch.bool, device=dev)
        w = torch.where(is_prev, torch.full((B,), 1.0 / p, device=dev),
                        torch.where(nb, torch.ones(B, device=dev), torch.full((B,), 1.0 / q, device=dev)))
        accept = torch.rand(B, device=dev, generator=gen) * wmax <= w
        take = pending & accept
        result = torch.where(take | (pending & (t == 63)), cand, result)
        pending = pending & ~accept
        # every rank must run the same number of (collective) proposal rounds
        more = pending.any().to(torch.int32).reshape(1)
        if csr.rt.world > 1:
            import torch.distributed as dist
            dist.all_reduce(more, op=dist.ReduceOp.MAX)
        if int(more.item()) == 0:
            break
    return result
