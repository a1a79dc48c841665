"""Tracing / profiling / metrics: NVTX ranges, CUDA-event device timers and a
steps/sec progress logger.  The reference has compile-time PROFILING timers with no call sites
(graphlearn/src/common/base/profiling.h:24-70) and prints LocalStep/sec in its TF trainers
(graphlearn/examples/tf/trainer.py:143-160)."""
from __future__ import annotations

import contextlib
import logging
import time
from collections import defaultdict

import torch

log = logging.getLogger("graphlearn_b200")


@contextlib.contextmanager
def nvtx_range(name: str):
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class DeviceTimer(object):
    """Accumulates device time per key with CUDA events (host wall clock on CPU)."""

    def __init__(self):
        self._pending = defaultdict(list)
        self.total_ms = defaultdict(float)
        self.count = defaultdict(int)

    @contextlib.contextmanager
    def section(self, key: str):
        if torch.cuda.is_available():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            yield
            e1.record()
            self._pending[key].append((e0, e1))
        else:
            t0 = time.perf_counter()
            yield
            self.total_ms[key] += (time.perf_counter() - t0) * 1e3
            self.count[key] += 1

    def flush(self):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        for key, evs in self._pending.items():
            for e0, e1 in evs:
                self.total_ms[key] += e0.elapsed_time(e1)
                self.count[key] += 1
        self._pending.clear()

    def summary(self) -> dict:
        self.flush()
        return {k: {"calls": self.count[k], "total_ms": round(v, 3), "mean_ms": round(v / max(self.count[k], 1), 4)}
                for k, v in self.total_ms.items()}


class ProgressLogger(object):
    """`LocalStep/sec`-style progress lines every `every` steps."""

    def __init__(self, every: int = 100, name: str = "train"):
        self.every, self.name = every, name
        self._t0 = time.perf_counter()
        self._step0 = 0
        self.step = 0

    def update(self, loss=None, n: int = 1):
        self.step += n
        if self.step % self.every == 0:
            dt = time.perf_counter() - self._t0
            rate = (self.step - self._step0) / max(dt, 1e-9)
            log.info("%s step %d  %.1f steps/s%s", self.name, self.step, rate,
                     "" if loss is None else "  loss %.4f" % float(loss))
            self._t0, self._step0 = time.perf_counter(), self.step
            return rate
        return None
