"""Seed traversal over the LOCAL shard (GetNodes / GetEdges operators, O17).

Modes (graphlearn/src/core/operator/graph/node_generator.h:111-264):
  by_order  cursor over the local ids; OutOfRange when a pass ends
  shuffle   a fresh permutation per epoch, traversed once (``shuffle(traverse=True)``)
  random    uniform draws with replacement, never ends
Traversal is unsharded in the reference (each server iterates what it stores);
here each rank iterates the nodes / edges it owns.  The state (epoch, cursor,
permutation seed) is tiny and checkpointable - the reference's Save()/Load()
stubs (node_generator.h:60-66) are real here.
"""
from __future__ import annotations

import torch

from .. import errors


class SeedIterator(object):
    def __init__(self, n: int, batch_size: int, strategy: str, device, seed: int = 0, drop_last: bool = False):
        self.n, self.bs, self.strategy = int(n), int(batch_size), strategy
        self.device = device
        self.seed = int(seed)
        self.drop_last = drop_last
        self.epoch = 0
        self.cursor = 0
        self._perm = None

    def _gen(self):
        g = torch.Generator(device="cpu")
        g.manual_seed(self.seed * 1000003 + self.epoch)
        return g

    def next_index(self) -> torch.Tensor:
        """local row indices of the next batch (device tensor); raises OutOfRangeError at epoch end."""
        if self.n == 0:
            self.epoch += 1
            raise errors.OutOfRangeError("no data in this shard")
        if self.strategy == "random":
            g = torch.Generator(device="cpu")
            g.manual_seed(self.seed * 7919 + self.cursor)
            self.cursor += 1
            return torch.randint(0, self.n, (self.bs,), generator=g).to(self.device)
        if self.cursor >= self.n or (self.drop_last and self.cursor + self.bs > self.n):
            self.cursor = 0
            self.epoch += 1
            self._perm = None
            raise errors.OutOfRangeError("end of epoch")
        end = min(self.cursor + self.bs, self.n)
        if self.strategy == "shuffle":
            if self._perm is None:
                self._perm = torch.randperm(self.n, generator=self._gen()).to(self.device)
            idx = self._perm[self.cursor:end]
        else:
            idx = torch.arange(self.cursor, end, device=self.device)
        self.cursor = end
        return idx

    def prime(self):
        """Build the current epoch's permutation now (``randperm`` of millions of rows costs tens of ms - callers
        that time their first batch, or pre-gather per-epoch values, do it up front)."""
        if self.strategy == "shuffle" and self._perm is None and self.n > 0:
            self._perm = torch.randperm(self.n, generator=self._gen()).to(self.device)
        return self._perm

    def has_next(self) -> bool:
        """would ``next_index()`` return a batch (True) or raise OutOfRangeError (False)?"""
        if self.n == 0:
            return False
        if self.strategy == "random":
            return True
        return not (self.cursor >= self.n or (self.drop_last and self.cursor + self.bs > self.n))

    def end_epoch(self):
        """close the current pass early (remaining indices are dropped) - used for multi-rank lock step"""
        if self.strategy != "random":
            self.cursor = 0
            self._perm = None
        self.epoch += 1

    def state_dict(self):
        return {"epoch": self.epoch, "cursor": self.cursor, "seed": self.seed, "strategy": self.strategy}

    def load_state_dict(self, sd):
        self.epoch, self.cursor, self.seed = int(sd["epoch"]), int(sd["cursor"]), int(sd["seed"])
        self._perm = None
