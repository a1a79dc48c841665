"""Device-resident RNG state shared by all sampling kernels.

``state`` is an int64[2] device tensor {seed, step offset}; kernels derive a
Philox4x32-10 stream from (seed, offset, op salt, element index)
(csrc/common.cuh).  Keeping the offset in device memory lets a captured CUDA
graph draw fresh samples on every replay: ``advance()`` is a 1-thread kernel
that is part of the graph.  The pair (seed, offset) is what a checkpoint stores
to resume sampling deterministically.
"""
from __future__ import annotations

import torch

from .. import config as _config
from ..parallel.runtime import Runtime, native


class DeviceRng:
    def __init__(self, rt: Runtime, seed: int = 0):
        self.rt = rt
        # decorrelate ranks: every rank samples its own seeds
        self.seed = int(seed) * 1000003 + rt.rank * 7919 + 12345
        self.state = torch.tensor([self.seed, 0], dtype=torch.int64, device=rt.device)
        self._host_offset = 0

    def advance(self, inc: int = 1):
        self._host_offset += inc
        if self.rt.is_cuda:
            native().rng_advance(self.state, int(inc))
        else:
            self.state[1] += inc

    def torch_generator(self, salt: int = 0) -> torch.Generator:
        """Generator for the portable torch path (advances on every call)."""
        g = torch.Generator(device=self.rt.device)
        off = int(self.state[1].item()) if not self.rt.is_cuda else self._host_offset
        g.manual_seed((self.seed * 2654435761 + off * 40503 + salt * 97 + self._bump()) % (2 ** 63 - 1))
        return g

    def _bump(self):
        self._calls = getattr(self, "_calls", 0) + 1
        return self._calls * 1315423911

    def state_dict(self):
        return {"seed": self.seed, "offset": int(self.state[1].item()), "calls": getattr(self, "_calls", 0)}

    def load_state_dict(self, sd):
        self.seed = int(sd["seed"])
        self.state[0] = self.seed
        self.state[1] = int(sd["offset"])
        self._host_offset = int(sd["offset"])
        self._calls = int(sd.get("calls", 0))


_DEFAULT = {}


def default_rng(rt: Runtime) -> DeviceRng:
    key = (id(rt), str(rt.device))
    if key not in _DEFAULT:
        _DEFAULT[key] = DeviceRng(rt, _config.get().seed)
    return _DEFAULT[key]


def reset_default():
    _DEFAULT.clear()
