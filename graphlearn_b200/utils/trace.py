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


class StepProfiler(object):
    """Chrome-trace dumps of selected training steps - the reference's ``profiling=True`` option writes
    ``timeline_<step>.json`` for steps 500-1000 every 100 (examples/tf/trainer.py:309-322,387-402).

        prof = StepProfiler("traces", start=500, stop=1000, every=100)
        for step in range(n):
            with prof.step(step):
                trainer.step(...)

    Selected steps run under ``torch.profiler`` (CPU + CUDA activities when a GPU is present) and are exported
    to ``<dir>/timeline_<step>.json`` (open in chrome://tracing or Perfetto); all other steps cost nothing."""

    def __init__(self, out_dir: str, start: int = 0, stop: int = 1 << 62, every: int = 1):
        import os
        self.out_dir, self.start, self.stop, self.every = out_dir, int(start), int(stop), max(1, int(every))
        os.makedirs(out_dir, exist_ok=True)
        self.written = []

    def selected(self, step: int) -> bool:
        return self.start <= step < self.stop and (step - self.start) % self.every == 0

    @contextlib.contextmanager
    def step(self, step: int):
        if not self.selected(step):
            yield
            return
        import os
        from torch.profiler import ProfilerActivity, profile
        acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if torch.cuda.is_available() else [])
        with profile(activities=acts) as prof:
            yield
            if torch.cuda.is_available():
                torch.cuda.synchronize()
        path = os.path.join(self.out_dir, "timeline_%d.json" % step)
        prof.export_chrome_trace(path)
        self.written.append(path)
